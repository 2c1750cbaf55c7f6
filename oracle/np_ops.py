"""CPU oracle (numpy, fp32) for the MRFA motion-refinement hot path primitives.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mrfa_b200/`` may import this module; it is
used by ``tests/``, by ``__graft_entry__.smoke()`` and by ``bench.py``'s CPU-baseline leg
as the *checker*, never as the thing being measured or shipped.

Every function restates one reference function (file:line given relative to the MRFA
reference checkout) in plain numpy, with the arithmetic replayed in float32 in the same
order the reference's PyTorch ops evaluate it.  The restatement is pinned against outputs of
the unmodified reference executed in the build container: see ``tests/golden/make_golden.py``
(generator) and ``tests/test_oracle_golden.py`` (check).  The reference ships no tests or
golden vectors of its own, so that live execution is the only pin.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------------------
# grids
# --------------------------------------------------------------------------------------
def make_coordinate_grid(h: int, w: int) -> np.ndarray:
    """modules/util.py:90-108 -- normalised [-1,1] mesh, last dim = (x, y), fp32.

    x_j = 2*(j/(w-1)) - 1 : true division first, then *2, then -1 (bit-exact contract).
    """
    with np.errstate(divide="ignore", invalid="ignore"):
        x = np.arange(w, dtype=F32) / F32(w - 1)
        y = np.arange(h, dtype=F32) / F32(h - 1)
    x = F32(2) * x - F32(1)
    y = F32(2) * y - F32(1)
    out = np.empty((h, w, 2), dtype=F32)
    out[..., 0] = x[None, :]
    out[..., 1] = y[:, None]
    return out


def coords_grid(batch: int, ht: int, wd: int) -> np.ndarray:
    """modules/util.py:53-56 -- integer pixel grid (B,2,ht,wd); channel 0 = x, 1 = y."""
    ys, xs = np.meshgrid(np.arange(ht), np.arange(wd), indexing="ij")
    g = np.stack([xs, ys], axis=0).astype(F32)
    return np.broadcast_to(g[None], (batch, 2, ht, wd)).copy()


def kp2gaussian(kp: np.ndarray, h: int, w: int, variance: float) -> np.ndarray:
    """modules/util.py:59-87 -- exp((-0.5 * |grid - kp|^2) / var); kp (...,2) -> (...,h,w)."""
    grid = make_coordinate_grid(h, w)                       # (h,w,2)
    kp = kp.astype(F32)
    d = grid.reshape((1,) * (kp.ndim - 1) + (h, w, 2)) - kp.reshape(kp.shape[:-1] + (1, 1, 2))
    sq = d * d
    s = sq[..., 0] + sq[..., 1]
    return np.exp((F32(-0.5) * s) / F32(variance)).astype(F32)


# --------------------------------------------------------------------------------------
# bilinear sampling (ATen grid_sampler_2d, bilinear, zeros / reflection padding)
# --------------------------------------------------------------------------------------
def _unnormalize(g: np.ndarray, size: int, align_corners: bool) -> np.ndarray:
    g = g.astype(F32)
    if align_corners:
        return ((g + F32(1)) / F32(2)) * F32(size - 1)
    # ATen rounds (g + 1) * size - 1 ONCE (an FMA in both its CUDA kernel and its vectorised CPU kernel, where it reads
    # (g + 1) * (size / 2) - 0.5): the fp64 product of an fp32 number and an integer is exact, so one cast reproduces it
    u = (g + F32(1)).astype(np.float64)
    return ((u * size - 1.0).astype(F32)) / F32(2)


def _reflect(x: np.ndarray, twice_low: int, twice_high: int) -> np.ndarray:
    if twice_low == twice_high:
        return np.zeros_like(x)
    mn = F32(twice_low) / F32(2)
    span = F32(twice_high - twice_low) / F32(2)
    x = np.abs(x - mn)
    extra = np.fmod(x, span)
    flips = np.floor(x / span).astype(np.int64)
    return np.where(flips % 2 == 0, extra + mn, span - extra + mn).astype(F32)


def grid_sample(img: np.ndarray, grid: np.ndarray, align_corners: bool = False,
                padding_mode: str = "zeros") -> np.ndarray:
    """torch.nn.functional.grid_sample(mode='bilinear') as called at raft.py:166,168,271,
    dense_motion.py:83 (align_corners=False), dense_motion.py:241 and util.py:34 (True),
    model.py:48 (reflection).  img (N,C,H,W), grid (N,Ho,Wo,2) normalised -> (N,C,Ho,Wo).
    """
    img = img.astype(F32)
    N, C, H, W = img.shape
    ix = _unnormalize(grid[..., 0], W, align_corners)
    iy = _unnormalize(grid[..., 1], H, align_corners)
    if padding_mode == "reflection":
        if align_corners:
            ix = _reflect(ix, 0, 2 * (W - 1))
            iy = _reflect(iy, 0, 2 * (H - 1))
        else:
            ix = _reflect(ix, -1, 2 * W - 1)
            iy = _reflect(iy, -1, 2 * H - 1)
        ix = np.clip(ix, 0, W - 1).astype(F32)
        iy = np.clip(iy, 0, H - 1).astype(F32)
    x0f = np.floor(ix)
    y0f = np.floor(iy)
    x1f = x0f + F32(1)
    y1f = y0f + F32(1)
    w_nw = (x1f - ix) * (y1f - iy)
    w_ne = (ix - x0f) * (y1f - iy)
    w_sw = (x1f - ix) * (iy - y0f)
    w_se = (ix - x0f) * (iy - y0f)
    # clamp before the integer cast so wild coordinates cannot overflow
    big = F32(1 << 30)
    x0 = np.clip(x0f, -big, big).astype(np.int64)
    y0 = np.clip(y0f, -big, big).astype(np.int64)
    x1 = x0 + 1
    y1 = y0 + 1
    out = np.zeros((N, C) + ix.shape[1:], dtype=F32)
    n_idx = np.arange(N).reshape(N, 1, 1)
    for (yy, xx, ww) in ((y0, x0, w_nw), (y0, x1, w_ne), (y1, x0, w_sw), (y1, x1, w_se)):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H) & np.isfinite(ix) & np.isfinite(iy)
        xc = np.clip(xx, 0, W - 1)
        yc = np.clip(yy, 0, H - 1)
        vals = img[n_idx, :, yc, xc]                         # (N,Ho,Wo,C)
        vals = np.where(ok[..., None], vals, F32(0))
        out += np.moveaxis(vals * ww[..., None].astype(F32), -1, 1)
    return out


def pixel_to_normalised(coords: np.ndarray, H: int, W: int) -> np.ndarray:
    """modules/util.py:29-33 -- 2*x/(W-1) - 1 in fp32."""
    c = coords.astype(F32)
    g = np.empty_like(c)
    g[..., 0] = (F32(2) * c[..., 0]) / F32(W - 1) - F32(1)
    g[..., 1] = (F32(2) * c[..., 1]) / F32(H - 1) - F32(1)
    return g


def bilinear_sampler(img: np.ndarray, coords: np.ndarray, mask: bool = False):
    """modules/util.py:26-38 -- pixel coordinates, align_corners=True, zeros padding."""
    H, W = img.shape[-2:]
    g = pixel_to_normalised(coords, H, W)
    out = grid_sample(img, g, align_corners=True)
    if mask:
        m = (g[..., 0:1] > -1) & (g[..., 1:2] > -1) & (g[..., 0:1] < 1) & (g[..., 1:2] < 1)
        return out, m.astype(F32)
    return out


def batch_bilinear_sampler(img, coords, h=256, w=256, mini_batch=4):
    """modules/util.py:40-51 -- chunked a6; trailing remainder chunks are dropped."""
    batch = img.shape[0] // (h * w)
    outs = []
    step = mini_batch * h * w
    for i in range(batch // mini_batch):
        outs.append(bilinear_sampler(img[i * step:(i + 1) * step], coords[i * step:(i + 1) * step]))
    return np.concatenate(outs, axis=0)


def deform_input(inp: np.ndarray, deformation: np.ndarray) -> np.ndarray:
    """FOMM-named warp; MRFA call sites raft.py:160-166 and :265-271.  The (B,h,w,2) grid is
    bilinearly resized (align_corners=True) to the feature size, then sampled with
    align_corners=False."""
    B, C, H, W = inp.shape
    if deformation.shape[1] != H or deformation.shape[2] != W:
        d = interpolate_bilinear_ac(np.moveaxis(deformation, -1, 1), H, W)
        deformation = np.moveaxis(d, 1, -1)
    return grid_sample(inp, deformation, align_corners=False)


# --------------------------------------------------------------------------------------
# resampling helpers
# --------------------------------------------------------------------------------------
def avg_pool2d(x: np.ndarray, k: int) -> np.ndarray:
    """F.avg_pool2d(x, k, stride=k) on the trailing two dims (raft.py:20, :219)."""
    *lead, H, W = x.shape
    x = x.astype(F32)[..., : (H // k) * k, : (W // k) * k]
    x = x.reshape(*lead, H // k, k, W // k, k)
    s = np.zeros(tuple(lead) + (H // k, W // k), dtype=F32)
    for a in range(k):
        for b in range(k):
            s = s + x[..., :, a, :, b]
    return (s / F32(k * k)).astype(F32)


def interpolate_bilinear_ac(x: np.ndarray, Ho: int, Wo: int) -> np.ndarray:
    """F.interpolate(mode='bilinear', align_corners=True) (raft.py:205,228,243,266,...)."""
    x = x.astype(F32)
    *lead, H, W = x.shape

    def axis(o, i):
        scale = F32(i - 1) / F32(o - 1) if o > 1 else F32(0)
        src = scale * np.arange(o, dtype=F32)
        i0 = np.minimum(np.floor(src).astype(np.int64), i - 1)
        i1 = np.minimum(i0 + 1, i - 1)
        l1 = (src - i0.astype(F32)).astype(F32)
        return i0, i1, F32(1) - l1, l1

    y0, y1, hy0, hy1 = axis(Ho, H)
    x0, x1, hx0, hx1 = axis(Wo, W)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    hy0 = hy0[:, None]
    hy1 = hy1[:, None]
    out = hy0 * (hx0 * top[..., x0] + hx1 * top[..., x1]) + hy1 * (hx0 * bot[..., x0] + hx1 * bot[..., x1])
    return out.astype(F32)


# --------------------------------------------------------------------------------------
# correlation volume, pyramid, lookup
# --------------------------------------------------------------------------------------
def round_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round to nearest even) -> fp32."""
    u = np.ascontiguousarray(x, dtype=F32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(F32)


def corr_volume(q_d: np.ndarray, k_s: np.ndarray, scale: float, bf16_inputs: bool = False) -> np.ndarray:
    """raft.py:183-185 -- corr[b,i,j] = scale * sum_c q_d[b,c,i] * k_s[b,c,j].

    q_d, k_s: (B,C,h,w).  Returns (B, h*w, h*w) fp32; i = driving pixel, j = source pixel.
    ``bf16_inputs`` rounds the operands to bf16 first (what the tensor-core path consumes).
    """
    B, C = q_d.shape[:2]
    fd = q_d.reshape(B, C, -1).astype(F32)
    fs = k_s.reshape(B, C, -1).astype(F32)
    if bf16_inputs:
        fd, fs = round_bf16(fd), round_bf16(fs)
    out = np.einsum("bci,bcj->bij", fd.astype(np.float64), fs.astype(np.float64))
    return (out * float(scale)).astype(F32)


def corr_pyramid_rows(volume: np.ndarray, h: int, w: int, k: int) -> np.ndarray:
    """raft.py:208,219,235-236 -- pool the *driving* (row) dims of the (B,N,N) volume by k.

    Returns (B, (h/k)*(w/k), N): the map each coarse-resolution query looks up.
    """
    B, N, M = volume.shape
    v = volume.reshape(B, h, w, M)
    v = np.moveaxis(v, -1, 1)                      # (B, M, h, w)  == '(b n) h w'
    v = avg_pool2d(v, k) if k > 1 else v
    return np.moveaxis(v.reshape(B, M, -1), 1, 2).copy()


def corr_lookup(level_maps, coords: np.ndarray, radius: int = 3) -> np.ndarray:
    """CorrBlock.__call__ raft.py:23-48.

    level_maps: list of (B*h1*w1, 1, H_l, W_l) maps (level 0 first).
    coords: (B,2,h1,w1) pixel coordinates on level 0 (channel 0 = x).
    Output (B, L*(2r+1)^2, h1, w1); channel k = lvl*(2r+1)^2 + a*(2r+1) + b samples at
    (x/2^lvl + (a-r), y/2^lvl + (b-r)) -- the first window index moves x (meshgrid quirk).
    """
    B, _, h1, w1 = coords.shape
    r = radius
    n = 2 * r + 1
    c = np.moveaxis(coords.astype(F32), 1, -1).reshape(B * h1 * w1, 1, 1, 2)
    d = np.linspace(-r, r, n, dtype=F32)
    delta = np.stack(np.meshgrid(d, d, indexing="ij"), axis=-1).reshape(1, n, n, 2)
    outs = []
    for lvl, m in enumerate(level_maps):
        cl = c / F32(2 ** lvl) + delta
        s = bilinear_sampler(m, cl)                # (P,1,n,n)
        outs.append(s.reshape(B, h1, w1, n * n))
    out = np.concatenate(outs, axis=-1)
    return np.ascontiguousarray(np.moveaxis(out, -1, 1)).astype(F32)


# --------------------------------------------------------------------------------------
# prior dense motion: sparse motions, deformed source
# --------------------------------------------------------------------------------------
def inverse_2x2(m: np.ndarray) -> np.ndarray:
    a, b, c, d = m[..., 0, 0], m[..., 0, 1], m[..., 1, 0], m[..., 1, 1]
    det = a * d - b * c
    out = np.empty_like(m)
    out[..., 0, 0] = d / det
    out[..., 0, 1] = -b / det
    out[..., 1, 0] = -c / det
    out[..., 1, 1] = a / det
    return out


def sparse_motions(kp_d, kp_s, h, w, jac_d=None, jac_s=None, bg_param=None) -> np.ndarray:
    """DenseMotionNetwork.create_sparse_motions dense_motion.py:48-76 -> (B,K+1,h,w,2)."""
    kp_d = kp_d.astype(F32)
    kp_s = kp_s.astype(F32)
    B, K, _ = kp_d.shape
    ident = make_coordinate_grid(h, w).reshape(1, 1, h, w, 2)
    cg = ident - kp_d.reshape(B, K, 1, 1, 2)
    if jac_d is not None:
        J = np.matmul(jac_s.astype(F32), inverse_2x2(jac_d.astype(F32))).astype(F32)  # (B,K,2,2)
        J = J.reshape(B, K, 1, 1, 2, 2)
        cg = (J[..., 0] * cg[..., None, 0] + J[..., 1] * cg[..., None, 1]).astype(F32)
    d2s = cg + kp_s.reshape(B, K, 1, 1, 2)
    bg = np.broadcast_to(ident, (B, 1, h, w, 2)).astype(F32)
    if bg_param is not None:
        P = bg_param.astype(F32).reshape(B, 1, 1, 1, 3, 3)
        hom = np.concatenate([bg, np.ones((B, 1, h, w, 1), F32)], axis=-1)
        t = (P[..., 0] * hom[..., None, 0] + P[..., 1] * hom[..., None, 1] + P[..., 2] * hom[..., None, 2])
        bg = (t[..., :2] / t[..., 2:3]).astype(F32)
    return np.concatenate([bg, d2s], axis=1).astype(F32)


def heatmap_representation(kp_d, kp_s, h, w, variance=0.01) -> np.ndarray:
    """dense_motion.py:36-46 / :200-210 -> (B,K+1,h,w): zero bg channel, then drv - src."""
    g = kp2gaussian(kp_d, h, w, variance) - kp2gaussian(kp_s, h, w, variance)
    return np.concatenate([np.zeros_like(g[:, :1]), g], axis=1)


def deformed_source(source: np.ndarray, motions: np.ndarray, align_corners: bool) -> np.ndarray:
    """dense_motion.py:78-85 (align_corners=False) / :235-243 (True) -> (B,K+1,C,h,w)."""
    B, K1, h, w, _ = motions.shape
    rep = np.repeat(source[:, None], K1, axis=1).reshape(B * K1, -1, h, w)
    out = grid_sample(rep, motions.reshape(B * K1, h, w, 2), align_corners=align_corners)
    return out.reshape(B, K1, -1, h, w)


# --------------------------------------------------------------------------------------
# thin-plate splines (kp mode)
# --------------------------------------------------------------------------------------
def tps_params(kp_1: np.ndarray, kp_2: np.ndarray):
    """TPS.__init__(mode='kp') util.py:355-383.  kp_1/kp_2: (B,G,n,2).

    Returns theta (B,G,2,3), control_points (B,G,n,2), control_params (B,G,n,2).
    The 8x8 solve is done in float64 and cast back (the reference uses an fp32 LU inverse;
    the systems are mildly conditioned so both agree far inside the 1e-5 warp tolerance
    only for well-separated key-points -- tests use the reference's own output as golden).
    """
    B, G, n, _ = kp_1.shape
    k1 = kp_1.astype(F32)
    diff = k1[:, :, :, None, :] - k1[:, :, None, :, :]
    dist = np.sqrt((diff * diff).sum(-1, dtype=F32)).astype(F32)
    Kmat = (dist * dist).astype(F32)
    Kmat = (Kmat * np.log(Kmat + F32(1e-9))).astype(F32)
    ones = np.ones((B, G, n, 1), F32)
    kp1p = np.concatenate([k1, ones], axis=3)                        # (B,G,n,3)
    P = np.concatenate([kp1p, np.zeros((B, G, 3, 3), F32)], axis=2)  # (B,G,n+3,3)
    L = np.concatenate([Kmat, np.swapaxes(kp1p, 2, 3)], axis=2)      # (B,G,n+3,n)
    L = np.concatenate([L, P], axis=3)                               # (B,G,n+3,n+3)
    L = L + (np.eye(n + 3, dtype=F32) * F32(0.01))
    Y = np.concatenate([kp_2.astype(F32), np.zeros((B, G, 3, 2), F32)], axis=2)
    param = np.linalg.solve(L.astype(np.float64), Y.astype(np.float64)).astype(F32)
    theta = np.swapaxes(param[:, :, n:, :], 2, 3)
    return theta, k1, param[:, :, :n, :]


def tps_warp_grid(theta, control_points, control_params, h, w) -> np.ndarray:
    """TPS.transform_frame + warp_coordinates (kp mode) util.py:387-410 -> (B,G,h,w,2)."""
    B, G, n, _ = control_points.shape
    grid = make_coordinate_grid(h, w).reshape(1, 1, h * w, 2)
    aff = np.einsum("bgij,xyni->bgnj", np.swapaxes(theta[..., :2], 2, 3).astype(F32), grid) \
        + theta[:, :, None, :, 2]
    d = grid.reshape(1, 1, 1, h * w, 2) - control_points.reshape(B, G, n, 1, 2)
    r2 = (d * d).sum(-1, dtype=F32)
    U = (r2 * np.log(r2 + F32(1e-9))).astype(F32)                    # (B,G,n,hw)
    res = np.einsum("bgnp,bgnc->bgpc", U, control_params.astype(F32))
    return (aff + res).astype(F32).reshape(B, G, h, w, 2)


def tps_transformations(kp_d, kp_s, h, w, bg_param=None) -> np.ndarray:
    """TPSDenseMotionNetwork.create_transformations dense_motion.py:212-233 -> (B,G+1,h,w,2)."""
    B = kp_d.shape[0]
    k1 = kp_d.reshape(B, -1, 5, 2)
    k2 = kp_s.reshape(B, -1, 5, 2)
    theta, cp, cw = tps_params(k1, k2)
    d2s = tps_warp_grid(theta, cp, cw, h, w)
    ident = np.broadcast_to(make_coordinate_grid(h, w).reshape(1, 1, h, w, 2), (B, 1, h, w, 2)).astype(F32)
    if bg_param is not None:
        P = bg_param.astype(F32).reshape(B, 1, 1, 1, 3, 3)
        hom = np.concatenate([ident, np.ones((B, 1, h, w, 1), F32)], axis=-1)
        t = (P[..., 0] * hom[..., None, 0] + P[..., 1] * hom[..., None, 1] + P[..., 2] * hom[..., None, 2])
        ident = (t[..., :2] / t[..., 2:3]).astype(F32)
    return np.concatenate([ident, d2s], axis=1).astype(F32)


# --------------------------------------------------------------------------------------
# training-only random warps (equivariance branch)
# --------------------------------------------------------------------------------------
def random_warp_coordinates(theta, control_points, control_params, coords, metric: str = "l1") -> np.ndarray:
    """Transform.warp_coordinates model.py:50-70 (metric 'l1': d = |dx|+|dy|, U = d^2 log(d+1e-6)) and
    TPS.warp_coordinates mode 'random' util.py:412-423 ('l2sq': r2 = dx^2+dy^2, U = r2 log(r2+1e-9)).
    theta (B,2,3), control_points (P,2), control_params (B,P) or None, coords (B|1,M,2) -> (B,M,2)."""
    theta = theta.astype(F32)
    c = np.broadcast_to(coords.astype(F32), (theta.shape[0],) + coords.shape[1:])
    out = (theta[:, None, :, 0] * c[..., 0:1] + theta[:, None, :, 1] * c[..., 1:2]).astype(F32) + theta[:, None, :, 2]
    if control_params is not None:
        d = c[:, :, None, :] - control_points.astype(F32).reshape(1, 1, -1, 2)
        if metric == "l1":
            r = np.abs(d).sum(-1, dtype=F32)
            U = (r * r).astype(F32) * np.log(r + F32(1e-6)).astype(F32)
        else:
            r = (d * d).sum(-1, dtype=F32)
            U = r * np.log(r + F32(1e-9)).astype(F32)
        out = out + (U.astype(F32) * control_params.astype(F32).reshape(theta.shape[0], 1, -1)).sum(-1, dtype=F32)[..., None]
    return out.astype(F32)


def random_warp_grid(theta, control_points, control_params, h: int, w: int, metric: str = "l1") -> np.ndarray:
    """Transform.transform_frame's grid (model.py:44-47) / TPS.transform_frame mode 'random' (util.py:387-395)."""
    grid = make_coordinate_grid(h, w).reshape(1, h * w, 2)
    return random_warp_coordinates(theta, control_points, control_params, grid, metric).reshape(-1, h, w, 2)


def transform_frame(frame, theta, control_points, control_params) -> np.ndarray:
    """Transform.transform_frame model.py:44-48: reflection-padded grid_sample (align_corners=False)."""
    g = random_warp_grid(theta, control_points, control_params, frame.shape[2], frame.shape[3], "l1")
    return grid_sample(frame, g, align_corners=False, padding_mode="reflection")


# --------------------------------------------------------------------------------------
# prior -> flow conversion
# --------------------------------------------------------------------------------------
def init_flow_from_prior(deformation: np.ndarray, h: int) -> np.ndarray:
    """raft.py:189-190 -- (h-1)*(deformation+1)/2 - id_grid ; uses self.h for both axes.

    deformation (B,h,w,2) normalised -> flow (B,2,h,w) in basic-resolution pixels.
    """
    B, H, W, _ = deformation.shape
    d = np.moveaxis(deformation.astype(F32), -1, 1)
    return ((F32(h - 1) * (d + F32(1))) / F32(2.0) - coords_grid(B, H, W)).astype(F32)
