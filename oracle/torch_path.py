"""CPU oracle of the full refinement path (prior dense motion -> RaftFlow), written with the
*stock PyTorch ops the reference itself executes* (F.grid_sample, F.avg_pool2d, einsum,
F.interpolate, torch.inverse), so its arithmetic is the reference's arithmetic.

TEST INFRASTRUCTURE ONLY -- never imported by ``mrfa_b200``.  Used by ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline / ``--impl reference`` legs of ``bench.py``.

It is a restatement, not a copy: the reference control flow (modules/raft.py:141-311,
modules/dense_motion.py:104-146 and :262-312) is re-derived as an explicit per-level schedule.
Pinned against the unmodified reference run in the build container with identical weights
and inputs (tests/golden/make_golden.py -> tests/golden/*.npz; tests/test_oracle_golden.py).
The dense-convolution blocks are the oracle's own plain nn.Conv2d / BatchNorm2d restatement
(oracle/conv_blocks.py): importing the oracle never loads the product package or its CUDA library.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from .conv_blocks import AntiAliasInterpolation2d, Hourglass, OcclusionAwareGenerator


# ------------------------------------------------------------------ primitives (util.py)
def make_coordinate_grid(spatial_size, dtype=torch.float32, device="cpu"):
    """util.py:90-108."""
    h, w = spatial_size
    x = 2 * (torch.arange(w, device=device).to(dtype) / (w - 1)) - 1
    y = 2 * (torch.arange(h, device=device).to(dtype) / (h - 1)) - 1
    return torch.stack([x[None, :].expand(h, w), y[:, None].expand(h, w)], dim=2).contiguous()


def coords_grid(batch, ht, wd, device="cpu"):
    """util.py:53-56."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing="ij")
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def kp2gaussian(kp, spatial_size, kp_variance):
    """util.py:59-87."""
    grid = make_coordinate_grid(spatial_size, kp.dtype, kp.device)
    lead = kp.shape[:-1]
    d = grid.view((1,) * len(lead) + grid.shape) - kp.view(lead + (1, 1, 2))
    return torch.exp(-0.5 * (d ** 2).sum(-1) / kp_variance)


def bilinear_sampler(img, coords):
    """util.py:26-38 (mask=False path)."""
    H, W = img.shape[-2:]
    gx = 2 * coords[..., 0:1] / (W - 1) - 1
    gy = 2 * coords[..., 1:2] / (H - 1) - 1
    return F.grid_sample(img, torch.cat([gx, gy], dim=-1), align_corners=True)


def _up(x, size):
    return F.interpolate(x, size=size, mode="bilinear", align_corners=True)


def corr_lookup(level0, coords, radius=3, num_levels=2):
    """CorrBlock raft.py:12-48: pyramid by 2x2 average pooling of the source dims, then a
    (2r+1)^2 bilinear window per level around coords / 2^level."""
    B, _, h1, w1 = coords.shape
    n = 2 * radius + 1
    d = torch.linspace(-radius, radius, n, device=coords.device)
    delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, n, n, 2)
    centre = coords.permute(0, 2, 3, 1).reshape(B * h1 * w1, 1, 1, 2)
    level, feats = level0, []
    for lvl in range(num_levels):
        if lvl:
            level = F.avg_pool2d(level, 2, stride=2)
        pts = centre / 2 ** lvl + delta
        if B > 1 and h1 >= 128:                                    # raft.py:39-40, util.py:40-51
            per = h1 * w1
            s = torch.cat([bilinear_sampler(level[b * per:(b + 1) * per], pts[b * per:(b + 1) * per])
                           for b in range(B)], dim=0)
        else:
            s = bilinear_sampler(level, pts)
        feats.append(s.view(B, h1, w1, n * n))
    return torch.cat(feats, dim=-1).permute(0, 3, 1, 2).contiguous().float()


# ------------------------------------------------------------------ prior dense motion
class DenseMotionOracle(nn.Module):
    """DenseMotionNetwork (dense_motion.py:8-146), FOMM / MTIA prior."""

    def __init__(self, block_expansion, num_blocks, max_features, num_kp, num_channels,
                 estimate_occlusion_map=True, scale_factor=1, kp_variance=0.01):
        super().__init__()
        self.hourglass = Hourglass(block_expansion, (num_kp + 1) * (num_channels + 1), num_blocks, max_features)
        self.mask = nn.Conv2d(self.hourglass.out_filters, num_kp + 1, kernel_size=7, padding=3)
        self.occlusion = nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=7, padding=3) \
            if estimate_occlusion_map else None
        self.num_kp, self.scale_factor, self.kp_variance = num_kp, scale_factor, kp_variance
        if scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, scale_factor)

    def sparse_motions(self, h, w, kp_driving, kp_source, bg_param):
        B, K = kp_source["kp"].shape[:2]
        ident = make_coordinate_grid((h, w)).view(1, 1, h, w, 2)
        local = ident - kp_driving["kp"].view(B, K, 1, 1, 2)
        if "jacobian" in kp_driving:
            J = torch.matmul(kp_source["jacobian"], torch.inverse(kp_driving["jacobian"]))
            local = torch.matmul(J.view(B, K, 1, 1, 2, 2), local.unsqueeze(-1)).squeeze(-1)
        moved = local + kp_source["kp"].view(B, K, 1, 1, 2)
        bg = ident.repeat(B, 1, 1, 1, 1)
        if bg_param is not None:
            hom = torch.cat([bg, torch.ones_like(bg[..., :1])], dim=-1)
            hom = torch.matmul(bg_param.view(B, 1, 1, 1, 3, 3), hom.unsqueeze(-1)).squeeze(-1)
            bg = hom[..., :2] / hom[..., 2:3]
        return torch.cat([bg, moved], dim=1)

    def forward(self, source_image, kp_driving, kp_source, bg_param=None):
        if self.scale_factor != 1:
            source_image = self.down(source_image)
        B, _, h, w = source_image.shape
        K1 = self.num_kp + 1
        heat = kp2gaussian(kp_driving["kp"], (h, w), self.kp_variance) \
            - kp2gaussian(kp_source["kp"], (h, w), self.kp_variance)
        heat = torch.cat([torch.zeros_like(heat[:, :1]), heat], dim=1).unsqueeze(2)
        motions = self.sparse_motions(h, w, kp_driving, kp_source, bg_param)
        rep = source_image[:, None].expand(B, K1, -1, h, w).reshape(B * K1, -1, h, w)
        deformed = F.grid_sample(rep, motions.view(B * K1, h, w, 2)).view(B, K1, -1, h, w)
        pred = self.hourglass(torch.cat([heat, deformed], dim=2).view(B, -1, h, w))
        logits = self.mask(pred)
        mask = F.softmax(logits, dim=1)
        out = {"sparse_deformed": deformed, "logit_mask": logits, "mask": mask,
               "deformation": (motions.permute(0, 1, 4, 2, 3) * mask.unsqueeze(2)).sum(1).permute(0, 2, 3, 1)}
        if self.occlusion is not None:
            out["occlusion"] = self.occlusion(pred)
        out["sparse_motion"] = motions
        return out


class TPSDenseMotionOracle(nn.Module):
    """TPSDenseMotionNetwork (dense_motion.py:150-312), multi_mask=False."""

    def __init__(self, block_expansion, num_blocks, max_features, num_tps, num_channels,
                 scale_factor=0.25, bg=False, multi_mask=False, kp_variance=0.01):
        super().__init__()
        assert not multi_mask, "multi_mask=True is broken in the reference (dense_motion.py:174)"
        if scale_factor != 1:
            self.down = AntiAliasInterpolation2d(num_channels, scale_factor)
        self.hourglass = Hourglass(block_expansion, num_channels * (num_tps + 1) + num_tps * 5 + 1,
                                   num_blocks, max_features)
        self.maps = nn.Conv2d(self.hourglass.out_filters, num_tps + 1, kernel_size=7, padding=3)
        self.occlusion = nn.ModuleList([nn.Conv2d(self.hourglass.out_filters, 1, kernel_size=7, padding=3)])
        self.scale_factor, self.num_tps, self.kp_variance = scale_factor, num_tps, kp_variance

    @staticmethod
    def tps_grid(kp_1, kp_2, h, w):
        """TPS(mode='kp') + transform_frame: util.py:355-410."""
        B, G, n, _ = kp_1.shape
        K = torch.norm(kp_1[:, :, :, None] - kp_1[:, :, None, :], dim=4, p=2) ** 2
        K = K * torch.log(K + 1e-9)
        kp1p = torch.cat([kp_1, torch.ones(B, G, n, 1)], 3)
        L = torch.cat([torch.cat([K, kp1p.permute(0, 1, 3, 2)], 2),
                       torch.cat([kp1p, torch.zeros(B, G, 3, 3)], 2)], 3)
        L = L + torch.eye(n + 3).expand(L.shape) * 0.01
        param = torch.matmul(torch.inverse(L), torch.cat([kp_2, torch.zeros(B, G, 3, 2)], 2))
        theta, weights = param[:, :, n:, :].permute(0, 1, 3, 2), param[:, :, :n, :]
        pts = make_coordinate_grid((h, w)).view(1, h * w, 2)
        aff = torch.matmul(theta[..., :2], pts.permute(0, 2, 1)) + theta[..., 2:]
        d2 = ((pts.view(1, 1, 1, -1, 2) - kp_1.view(B, G, -1, 1, 2)) ** 2).sum(-1)
        rbf = torch.matmul((d2 * torch.log(d2 + 1e-9)).permute(0, 1, 3, 2), weights)
        return (aff.permute(0, 1, 3, 2) + rbf).view(B, G, h, w, 2)

    def forward(self, source_image, kp_driving, kp_source, bg_param=None):
        if self.scale_factor != 1:
            source_image = self.down(source_image)
        B, _, h, w = source_image.shape
        G1 = self.num_tps + 1
        heat = kp2gaussian(kp_driving["kp"], (h, w), self.kp_variance) \
            - kp2gaussian(kp_source["kp"], (h, w), self.kp_variance)
        heat = torch.cat([torch.zeros_like(heat[:, :1]), heat], dim=1)
        moved = self.tps_grid(kp_driving["kp"].view(B, -1, 5, 2), kp_source["kp"].view(B, -1, 5, 2), h, w)
        bg = make_coordinate_grid((h, w)).view(1, 1, h, w, 2).repeat(B, 1, 1, 1, 1)
        if bg_param is not None:
            hom = torch.cat([bg, torch.ones_like(bg[..., :1])], dim=-1)
            hom = torch.matmul(bg_param.view(B, 1, 1, 1, 3, 3), hom.unsqueeze(-1)).squeeze(-1)
            bg = hom[..., :2] / hom[..., 2:3]
        motions = torch.cat([bg, moved], dim=1)
        rep = source_image[:, None].expand(B, G1, -1, h, w).reshape(B * G1, -1, h, w)
        deformed = F.grid_sample(rep, motions.view(B * G1, h, w, 2), align_corners=True).view(B, G1, -1, h, w)
        pred = self.hourglass(torch.cat([heat, deformed.view(B, -1, h, w)], dim=1))
        maps = F.softmax(self.maps(pred), dim=1)
        return {"deformed_source": deformed, "contribution_maps": maps, "mask": maps,
                "deformation": (motions.permute(0, 1, 4, 2, 3) * maps.unsqueeze(2)).sum(1).permute(0, 2, 3, 1),
                "occlusion": self.occlusion[0](pred), "transformations": motions}


# ------------------------------------------------------------------ refinement decoder
class _MotionEncoder(nn.Module):
    """BasicMotionEncoder raft.py:50-68."""

    def __init__(self, cor_planes=98):
        super().__init__()
        self.convc1 = nn.Conv2d(cor_planes, 128, 1)
        self.convc2 = nn.Conv2d(128, 96, 3, padding=1)
        self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
        self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
        self.conv = nn.Conv2d(160, 126, 3, padding=1)

    def forward(self, flow, corr):
        c = F.relu(self.convc2(F.relu(self.convc1(corr))))
        f = F.relu(self.convf2(F.relu(self.convf1(flow))))
        return torch.cat([F.relu(self.conv(torch.cat([c, f], dim=1))), flow], dim=1)


class _Refiner(nn.Module):
    """RefineFlow raft.py:70-87 -> (dflow(2) ++ docc(1))."""

    def __init__(self):
        super().__init__()
        self.convc1 = nn.Conv2d(192, 128, 3, padding=1)
        self.conv1 = nn.Conv2d(256, 128, 3, padding=1)
        self.conv2 = nn.Conv2d(128, 2, 3, padding=1)
        self.convo1 = nn.Conv2d(256, 128, 3, padding=1)
        self.convo2 = nn.Conv2d(128, 1, 3, padding=1)

    def forward(self, motion_feat, context):
        x = torch.cat([motion_feat, F.relu(self.convc1(context))], dim=1)
        return torch.cat([self.conv2(F.relu(self.conv1(x))), self.convo2(F.relu(self.convo1(x)))], dim=1)


class RaftFlowOracle(nn.Module):
    """RaftFlow (raft.py:91-311)."""

    FEATURE_WIDTHS = (512, 512, 512, 256, 128, 64)          # raft.py:105-113

    def __init__(self, prior_only=False, num_kp=10, dim=256, size=256, generator=None,
                 driving_encoder=None, source_encoder=None):
        super().__init__()
        self.scale = dim ** -0.5
        self.size, self.h, self.w, self.prior_only = size, size // 4, size // 4, prior_only
        self.generator = OcclusionAwareGenerator(**generator)
        self.levels = 6                                        # log2(32) + 1, raft.py:114-115
        self.base = int(math.log((size // 4) // (size // 32), 2))
        if prior_only:
            return
        self.kp = Hourglass(**driving_encoder)
        self.kp_img = Hourglass(**source_encoder)
        self.kp_head = nn.Conv2d(self.kp.out_filters, dim, 1)
        self.kp_img_head = nn.Conv2d(self.kp_img.out_filters, dim, 1)
        self.pos_embedding = nn.Parameter(torch.zeros(1, num_kp, self.h, self.w))
        nn.init.trunc_normal_(self.pos_embedding, std=.02)
        self.corr_enc = _MotionEncoder()
        self.refine = _Refiner()
        self.to_context = nn.ModuleList(nn.Conv2d(c, 192, 1) for c in self.FEATURE_WIDTHS)

    # -- raft.py:155-173
    def _prior_only(self, feature, dense_motion, img_full):
        grid, occ = dense_motion["deformation"], dense_motion["occlusion"]
        warps, occs = [], []
        g = grid
        for f in feature:
            R = f.shape[2:]
            g = grid if grid.shape[2] == f.shape[2] else _up(grid.permute(0, 3, 1, 2), R).permute(0, 2, 3, 1)
            o = occ if grid.shape[2] == f.shape[2] else _up(occ, R)
            warps.append(F.grid_sample(f, g))
            occs.append(torch.sigmoid(o))
        warp_img = F.grid_sample(img_full, g)
        out = self.generator.decode(warps, warp_img, occs)
        vis = torch.cat([_up(o, (self.size, self.size)) for o in occs], dim=3)
        return out, warp_img, vis

    def correlation(self, kp_s, kp_d, img):
        """raft.py:177-185 -> (B, N_drv, N_src) fp32."""
        h, w = img.shape[2:]
        g_s = kp2gaussian(kp_s, (h, w), 0.1) + self.pos_embedding
        g_d = kp2gaussian(kp_d, (h, w), 0.1) + self.pos_embedding
        k_s = self.kp_img_head(self.kp_img(torch.cat([g_s, img], dim=1)))
        q_d = self.kp_head(self.kp(g_d))
        return torch.einsum("bic,bjc->bij", q_d.flatten(2).transpose(1, 2), k_s.flatten(2).transpose(1, 2)) * self.scale, q_d, k_s

    def forward(self, kp_s, kp_d, dense_motion, img, img_full, return_trace=False):
        feature = self.generator.encode(img_full)
        if self.prior_only:
            return self._prior_only(feature, dense_motion, img_full)
        B = img.shape[0]
        h, w, N = self.h, self.w, self.h * self.w
        trace = {}
        volume, q_d, k_s = self.correlation(kp_s, kp_d, img)
        prior = dense_motion["deformation"]
        prior_occ = dense_motion["occlusion"]
        flow0 = (h - 1) * (prior.permute(0, 3, 1, 2) + 1) / 2.0 - coords_grid(B, h, w)       # raft.py:190
        flow = F.interpolate(flow0, scale_factor=1 / 8, mode="bilinear", align_corners=True) / 8.0
        occ = F.interpolate(prior_occ, scale_factor=1 / 8, mode="bilinear", align_corners=True)
        # rows = driving pixel i, viewed as maps over the driving plane for pooling (raft.py:208)
        by_source = volume.transpose(1, 2).reshape(B * N, 1, h, w)
        warps, occs, warps_c, occs_c = [], [], [], []
        acc_flow = acc_occ = None
        for i in range(self.levels):
            R = self.size // 32 * 2 ** i
            ident = coords_grid(B, R, R)
            if i < self.base:                                    # coarser than the volume
                k = 2 ** (self.base - i)
                pooled = F.avg_pool2d(by_source, k, stride=k)
                query, mult = flow + ident, float(k)
            else:
                pooled = by_source
                if i == self.base:
                    query, mult = flow + ident, 1.0
                else:                                            # finer: sample at basic res
                    query = _up(flow, (h, h)) * 0.5 ** (i - self.base) + coords_grid(B, h, w)
                    mult = 1.0
            Rq = pooled.shape[-1]
            maps = pooled.view(B, N, Rq * Rq).transpose(1, 2).reshape(B * Rq * Rq, 1, h, w)
            corr = corr_lookup(maps, query * mult)
            if i > self.base:
                corr = _up(corr, (R, R))
            m_f = self.corr_enc(flow, corr)
            ctx = F.relu(self.to_context[i](bilinear_sampler(feature[i], (flow + ident).permute(0, 2, 3, 1))))
            delta = self.refine(m_f, ctx)
            d_flow, d_occ = delta[:, :2], delta[:, 2:]
            flow_w = flow + d_flow
            occ = occ + d_occ
            warps.append(bilinear_sampler(feature[i], (flow_w + ident).permute(0, 2, 3, 1)))
            occs.append(torch.sigmoid(occ))
            if i != self.base:
                g = _up(prior.permute(0, 3, 1, 2), (R, R)).permute(0, 2, 3, 1)
                o = _up(prior_occ, (R, R))
            else:
                g, o = prior, prior_occ
            warps_c.append(F.grid_sample(feature[i], g))
            occs_c.append(torch.sigmoid(o))
            if return_trace:
                trace[f"corr{i}"], trace[f"flow{i}"] = corr, flow
            if i < self.levels - 1:                              # raft.py:276-295
                R2 = 2 * R
                up_d = _up(d_flow, (R2, R2)) * 2
                flow = up_d + _up(flow0, (R2, R2)) / (2 ** (self.base - i) / 2.0)
                up_o = _up(d_occ, (R2, R2))
                occ = up_o + _up(prior_occ, (R2, R2))
                if acc_flow is None:
                    acc_flow, acc_occ = up_d, up_o
                else:
                    carried_f = _up(acc_flow, (R2, R2)) * 2
                    carried_o = _up(acc_occ, (R2, R2))
                    flow, occ = flow + carried_f, occ + carried_o
                    acc_flow, acc_occ = up_d + carried_f, up_o + carried_o
        warp_img = bilinear_sampler(img_full, (flow + ident).permute(0, 2, 3, 1))
        out = self.generator.decode(warps, warp_img, occs, warps_c, occs_c)
        vis = torch.cat([_up(o, (self.size, self.size)) for o in occs + [torch.sigmoid(prior_occ)]], dim=3)
        if return_trace:
            trace.update(volume=volume, q_d=q_d, k_s=k_s, flow_final=flow, warps=warps, warps_c=warps_c)
            return out, warp_img, vis, trace
        return out, warp_img, vis
