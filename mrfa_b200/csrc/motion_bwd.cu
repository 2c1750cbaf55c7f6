// Backward of the prior dense-motion synthesis (motion.cu): gradients of the fused heat-map / sparse-motion /
// deformed-source kernels with respect to the key-points, Jacobians, background affine, thin-plate-spline
// parameters and the source image (reference forward: dense_motion.py:36-85, :200-243, util.py:59-87, :355-410;
// the reference's backward is autograd over ~20 (B,K,h,w)-sized eager ops).
//
// Every thread re-derives its pixel's forward quantities (the forward keeps nothing but its inputs), forms the
// per-pixel contributions, and the block reduces them with warp shuffles before ONE atomicAdd per value and block
// into a small caller-zeroed accumulator; a finishing kernel turns the accumulated d(J) into d(jac_s), d(jac_d)
// (J = jac_s * inverse(jac_d)) and, for thin-plate splines, runs the adjoint 8x8 solve in fp64.
// Source-image gradients are a bilinear scatter (red.global.add.f32), pre-reduced over nothing: K+1 motion fields
// scatter into 3 planes of 64 x 64, the atomics stay in L2.
#include "common.cuh"

namespace mrfa {

// Bilinear sample geometry with the pieces the coordinate gradient needs (zeros padding).
struct TapsG {
  int o_nw, o_ne, o_sw, o_se;
  float ax, ay, bx, by;
  bool v_nw, v_ne, v_sw, v_se;
};

__device__ __forceinline__ TapsG make_taps_g(float ix, float iy, int H, int W) {
  TapsG t;
  const float fx = floorf(ix), fy = floorf(iy);
  const bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
  const int x0 = finite ? (int)fx : -2, y0 = finite ? (int)fy : -2;
  const int x1 = x0 + 1, y1 = y0 + 1;
  t.ax = ix - fx; t.ay = iy - fy;
  t.bx = (fx + 1.f) - ix; t.by = (fy + 1.f) - iy;
  const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
  const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
  const int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  const int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  t.o_nw = cy0 * W + cx0; t.o_ne = cy0 * W + cx1;
  t.o_sw = cy1 * W + cx0; t.o_se = cy1 * W + cx1;
  t.v_nw = finite & vx0 & vy0; t.v_ne = finite & vx1 & vy0;
  t.v_sw = finite & vx0 & vy1; t.v_se = finite & vx1 & vy1;
  return t;
}

// d(loss)/d(sample position in pixels) of out[c] = bilinear(src[c]) for the C planes of one pixel, and the
// scatter of the output gradients into grad_source.  g points at the first plane's gradient (plane stride hw).
__device__ __forceinline__ float2 sample_source_bwd(const float* __restrict__ src, float* __restrict__ gsrc,
                                                    const float* __restrict__ g, int C, int hw, const TapsG& t) {
  float gix = 0.f, giy = 0.f;
  for (int c = 0; c < C; ++c) {
    const float go = __ldg(g + (int64_t)c * hw);
    const float* s = src + (int64_t)c * hw;
    const float nw = t.v_nw ? __ldg(s + t.o_nw) : 0.f, ne = t.v_ne ? __ldg(s + t.o_ne) : 0.f;
    const float sw = t.v_sw ? __ldg(s + t.o_sw) : 0.f, se = t.v_se ? __ldg(s + t.o_se) : 0.f;
    gix = fmaf(go, fmaf(t.by, ne - nw, t.ay * (se - sw)), gix);
    giy = fmaf(go, fmaf(t.bx, sw - nw, t.ax * (se - ne)), giy);
    if (gsrc != nullptr && go != 0.f) {
      float* d = gsrc + (int64_t)c * hw;
      if (t.v_nw) atomicAdd(d + t.o_nw, go * (t.bx * t.by));
      if (t.v_ne) atomicAdd(d + t.o_ne, go * (t.ax * t.by));
      if (t.v_sw) atomicAdd(d + t.o_sw, go * (t.bx * t.ay));
      if (t.v_se) atomicAdd(d + t.o_se, go * (t.ax * t.ay));
    }
  }
  return make_float2(gix, giy);
}

// Block reduction of NV per-thread values (blockDim.x == 256) and one atomicAdd per value into dst.
template <int NV>
__device__ __forceinline__ void block_accumulate(float (&v)[NV], float* __restrict__ dst) {
  __shared__ float part[8][NV];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) part[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += part[wv][threadIdx.x];
    if (s != 0.f) atomicAdd(dst + threadIdx.x, s);
  }
}

// background affine m = (t0/t2, t1/t2), t = P (gx, gy, 1): contributions to the 9 entries of P
__device__ __forceinline__ void bg_affine_bwd(const float* __restrict__ P, float gx, float gy, float2 gm, float (&v)[9]) {
  float t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) t[r] = fmaf(__ldg(P + 3 * r), gx, fmaf(__ldg(P + 3 * r + 1), gy, __ldg(P + 3 * r + 2)));
  const float it2 = 1.f / t[2];
  const float d0 = gm.x * it2, d1 = gm.y * it2;
  const float d2 = -(gm.x * t[0] + gm.y * t[1]) * it2 * it2;
  v[0] = d0 * gx; v[1] = d0 * gy; v[2] = d0;
  v[3] = d1 * gx; v[4] = d1 * gy; v[5] = d1;
  v[6] = d2 * gx; v[7] = d2 * gy; v[8] = d2;
}

__device__ __forceinline__ float2 bg_affine_fwd(const float* __restrict__ P, float gx, float gy) {
  float t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    t[r] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(P + 3 * r), gx), __fmul_rn(__ldg(P + 3 * r + 1), gy)), __ldg(P + 3 * r + 2));
  return make_float2(__fdiv_rn(t[0], t[2]), __fdiv_rn(t[1], t[2]));
}

constexpr int kAccStride = 12;       // accumulator floats per (b, k): FOMM [d kp_d 2 | d kp_s 2 | dJ 4], background [dP 9]

struct PriorKpB {
  float kdx, kdy, ksx, ksy, j00, j01, j10, j11;
};

// Same thread mapping as dense_motion_prior_kernel: block = 256 pixels of one (b, k).
__global__ void __launch_bounds__(256)
dense_motion_prior_bwd_kernel(const float* __restrict__ grad_motions, const float* __restrict__ grad_hg,
                              const float* __restrict__ kp_d, const float* __restrict__ kp_s,
                              const float* __restrict__ jac_d, const float* __restrict__ jac_s,
                              const float* __restrict__ bg_param, const float* __restrict__ source,
                              float* __restrict__ acc, float* __restrict__ grad_source,
                              int B, int K, int C, int h, int w, float variance) {
  __shared__ PriorKpB sk;
  const int hw = h * w;
  const int bk = blockIdx.y;
  const int b = bk / (K + 1), k = bk - b * (K + 1);
  if (threadIdx.x == 0 && k > 0) {
    const int kk = b * K + (k - 1);
    PriorKpB p;
    p.kdx = __ldg(kp_d + 2 * kk); p.kdy = __ldg(kp_d + 2 * kk + 1);
    p.ksx = __ldg(kp_s + 2 * kk); p.ksy = __ldg(kp_s + 2 * kk + 1);
    p.j00 = 1.f; p.j01 = 0.f; p.j10 = 0.f; p.j11 = 1.f;
    if (jac_d != nullptr) {
      const float a = __ldg(jac_d + 4 * kk), bb = __ldg(jac_d + 4 * kk + 1);
      const float c = __ldg(jac_d + 4 * kk + 2), d = __ldg(jac_d + 4 * kk + 3);
      const float det = __fsub_rn(__fmul_rn(a, d), __fmul_rn(bb, c));
      const float i00 = __fdiv_rn(d, det), i01 = __fdiv_rn(-bb, det), i10 = __fdiv_rn(-c, det), i11 = __fdiv_rn(a, det);
      const float s00 = __ldg(jac_s + 4 * kk), s01 = __ldg(jac_s + 4 * kk + 1);
      const float s10 = __ldg(jac_s + 4 * kk + 2), s11 = __ldg(jac_s + 4 * kk + 3);
      p.j00 = __fadd_rn(__fmul_rn(s00, i00), __fmul_rn(s01, i10));
      p.j01 = __fadd_rn(__fmul_rn(s00, i01), __fmul_rn(s01, i11));
      p.j10 = __fadd_rn(__fmul_rn(s10, i00), __fmul_rn(s11, i10));
      p.j11 = __fadd_rn(__fmul_rn(s10, i01), __fmul_rn(s11, i11));
    }
    sk = p;
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < hw;
  const int rr = live ? r : 0;
  const int y = rr / w, x = rr - y * w;
  const float gx = norm_coord(x, w), gy = norm_coord(y, h);
  const float* ghg = grad_hg + (int64_t)bk * (C + 1) * hw + rr;
  float2 gm = make_float2(0.f, 0.f);
  if (live && grad_motions != nullptr) gm = __ldg(reinterpret_cast<const float2*>(grad_motions) + (int64_t)bk * hw + rr);
  float* gsrc = grad_source ? grad_source + (int64_t)b * C * hw : nullptr;
  const float* src = source + (int64_t)b * C * hw;
  const float sx_pix = to_pixel_grad<MRFA_COORD_NORM_ACF>(w), sy_pix = to_pixel_grad<MRFA_COORD_NORM_ACF>(h);

  if (k == 0) {
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = 0.f;
    if (live) {
      const float2 m = bg_param ? bg_affine_fwd(bg_param + 9 * b, gx, gy) : make_float2(gx, gy);
      const TapsG t = make_taps_g(to_pixel<MRFA_COORD_NORM_ACF>(m.x, w), to_pixel<MRFA_COORD_NORM_ACF>(m.y, h), h, w);
      const float2 gp = sample_source_bwd(src, gsrc, ghg + hw, C, hw, t);
      gm.x = fmaf(gp.x, sx_pix, gm.x);
      gm.y = fmaf(gp.y, sy_pix, gm.y);
      if (bg_param) bg_affine_bwd(bg_param + 9 * b, gx, gy, gm, v);
    }
    if (bg_param) block_accumulate<9>(v, acc + (int64_t)bk * kAccStride);
    return;
  }

  float v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = 0.f;
  if (live) {
    const PriorKpB p = sk;
    const float cx = __fsub_rn(gx, p.kdx), cy = __fsub_rn(gy, p.kdy);
    const float sx = __fsub_rn(gx, p.ksx), sy = __fsub_rn(gy, p.ksy);
    const float dd = __fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), ds = __fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy));
    const float ed = expf(__fdiv_rn(__fmul_rn(-0.5f, dd), variance)), es = expf(__fdiv_rn(__fmul_rn(-0.5f, ds), variance));
    float nx = cx, ny = cy;
    if (jac_d != nullptr) {
      nx = __fadd_rn(__fmul_rn(p.j00, cx), __fmul_rn(p.j01, cy));
      ny = __fadd_rn(__fmul_rn(p.j10, cx), __fmul_rn(p.j11, cy));
    }
    const float2 m = make_float2(__fadd_rn(nx, p.ksx), __fadd_rn(ny, p.ksy));
    const TapsG t = make_taps_g(to_pixel<MRFA_COORD_NORM_ACF>(m.x, w), to_pixel<MRFA_COORD_NORM_ACF>(m.y, h), h, w);
    const float2 gp = sample_source_bwd(src, gsrc, ghg + hw, C, hw, t);
    gm.x = fmaf(gp.x, sx_pix, gm.x);
    gm.y = fmaf(gp.y, sy_pix, gm.y);
    // heat = exp(-0.5 |g - kd|^2 / var) - exp(-0.5 |g - ks|^2 / var)
    const float gh = __ldg(ghg);
    const float hd = gh * ed / variance, hs = gh * es / variance;
    // m = J (g - kd) + ks
    v[0] = fmaf(hd, cx, -(p.j00 * gm.x + p.j10 * gm.y));
    v[1] = fmaf(hd, cy, -(p.j01 * gm.x + p.j11 * gm.y));
    v[2] = fmaf(-hs, sx, gm.x);
    v[3] = fmaf(-hs, sy, gm.y);
    v[4] = gm.x * cx; v[5] = gm.x * cy; v[6] = gm.y * cx; v[7] = gm.y * cy;
  }
  block_accumulate<8>(v, acc + (int64_t)bk * kAccStride);
}

// One thread per (b, k): accumulator -> grad_kp_d, grad_kp_s, grad_jac_d, grad_jac_s; thread (b, 0) -> grad_bg.
__global__ void __launch_bounds__(128)
dense_motion_prior_bwd_finish_kernel(const float* __restrict__ acc, const float* __restrict__ jac_d,
                                     const float* __restrict__ jac_s, float* __restrict__ grad_kp_d,
                                     float* __restrict__ grad_kp_s, float* __restrict__ grad_jac_d,
                                     float* __restrict__ grad_jac_s, float* __restrict__ grad_bg, int B, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * (K + 1)) return;
  const int b = i / (K + 1), k = i - b * (K + 1);
  const float* a = acc + (int64_t)i * kAccStride;
  if (k == 0) {
    if (grad_bg != nullptr)
      for (int j = 0; j < 9; ++j) grad_bg[9 * b + j] = a[j];
    return;
  }
  const int kk = b * K + (k - 1);
  grad_kp_d[2 * kk] = a[0]; grad_kp_d[2 * kk + 1] = a[1];
  grad_kp_s[2 * kk] = a[2]; grad_kp_s[2 * kk + 1] = a[3];
  if (jac_d != nullptr && grad_jac_d != nullptr) {
    // J = S inv(D):  dS = dJ inv(D)^T,  d(inv) = S^T dJ,  dD = -inv(D)^T d(inv) inv(D)^T
    const float d00 = jac_d[4 * kk], d01 = jac_d[4 * kk + 1], d10 = jac_d[4 * kk + 2], d11 = jac_d[4 * kk + 3];
    const float det = d00 * d11 - d01 * d10;
    const float i00 = d11 / det, i01 = -d01 / det, i10 = -d10 / det, i11 = d00 / det;
    const float s00 = jac_s[4 * kk], s01 = jac_s[4 * kk + 1], s10 = jac_s[4 * kk + 2], s11 = jac_s[4 * kk + 3];
    const float g00 = a[4], g01 = a[5], g10 = a[6], g11 = a[7];
    grad_jac_s[4 * kk + 0] = g00 * i00 + g01 * i01;
    grad_jac_s[4 * kk + 1] = g00 * i10 + g01 * i11;
    grad_jac_s[4 * kk + 2] = g10 * i00 + g11 * i01;
    grad_jac_s[4 * kk + 3] = g10 * i10 + g11 * i11;
    const float e00 = s00 * g00 + s10 * g10, e01 = s00 * g01 + s10 * g11;
    const float e10 = s01 * g00 + s11 * g10, e11 = s01 * g01 + s11 * g11;
    // T = inv^T E
    const float t00 = i00 * e00 + i10 * e10, t01 = i00 * e01 + i10 * e11;
    const float t10 = i01 * e00 + i11 * e10, t11 = i01 * e01 + i11 * e11;
    // dD = -T inv^T
    grad_jac_d[4 * kk + 0] = -(t00 * i00 + t01 * i01);
    grad_jac_d[4 * kk + 1] = -(t00 * i10 + t01 * i11);
    grad_jac_d[4 * kk + 2] = -(t10 * i00 + t11 * i01);
    grad_jac_d[4 * kk + 3] = -(t10 * i10 + t11 * i11);
  }
}

// ---- key-point heat-maps (util.py:59-87): out = exp(-0.5 |g - kp|^2 / var) -> d kp ------------------------------
// block = 256 pixels of one heat-map p; grad (P,h,w) -> grad_kp (P,2), caller-zeroed.  `sign_pair`: the TPS prior
// stores heat = gauss(kp_d) - gauss(kp_s) in channel 1 + p of a (chan_total) channel block; then grad_kp2 receives
// the source side.
__global__ void __launch_bounds__(256)
kp2gaussian_bwd_kernel(const float* __restrict__ grad, const float* __restrict__ kp, const float* __restrict__ kp2,
                       float* __restrict__ grad_kp, float* __restrict__ grad_kp2, int per_batch, int64_t batch_stride,
                       int64_t chan_off, int h, int w, float variance) {
  const int hw = h * w;
  const int p = blockIdx.y;
  const int b = p / per_batch, k = p - b * per_batch;
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  if (r < hw) {
    const int y = r / w, x = r - y * w;
    const float gx = norm_coord(x, w), gy = norm_coord(y, h);
    const float g = __ldg(grad + (int64_t)b * batch_stride + (chan_off + k) * hw + r);
    {
      const float dx = __fsub_rn(gx, __ldg(kp + 2 * p)), dy = __fsub_rn(gy, __ldg(kp + 2 * p + 1));
      const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      const float e = g * expf(__fdiv_rn(__fmul_rn(-0.5f, s), variance)) / variance;
      v[0] = e * dx; v[1] = e * dy;
    }
    if (kp2 != nullptr) {
      const float dx = __fsub_rn(gx, __ldg(kp2 + 2 * p)), dy = __fsub_rn(gy, __ldg(kp2 + 2 * p + 1));
      const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      const float e = -g * expf(__fdiv_rn(__fmul_rn(-0.5f, s), variance)) / variance;
      v[2] = e * dx; v[3] = e * dy;
    }
  }
  __shared__ float part[8][4];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float s = warp_sum(v[i]);
    if (lane == 0) part[warp][i] = s;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += part[wv][threadIdx.x];
    if (threadIdx.x < 2) { if (s != 0.f) atomicAdd(grad_kp + 2 * p + threadIdx.x, s); }
    else if (kp2 != nullptr && s != 0.f) atomicAdd(grad_kp2 + 2 * p + threadIdx.x - 2, s);
  }
}

// ---- thin-plate splines -----------------------------------------------------------------------------------------
// accumulator per (b, g), 36 floats: g == 0 -> dP (9); g > 0 -> [d theta 6 | d control_params 10 | d control points 10]
constexpr int kTpsAccStride = 36;

__global__ void __launch_bounds__(256)
tps_motion_prior_bwd_kernel(const float* __restrict__ grad_motions, const float* __restrict__ grad_hg,
                            const float* __restrict__ kp_d, const float* __restrict__ theta,
                            const float* __restrict__ control_params, const float* __restrict__ bg_param,
                            const float* __restrict__ source, float* __restrict__ acc, float* __restrict__ grad_source,
                            int B, int G, int C, int chan_total, int chan_off, int h, int w) {
  const int hw = h * w;
  const int bg = blockIdx.y;
  const int b = bg / (G + 1), g = bg - b * (G + 1);
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = r < hw;
  const int rr = live ? r : 0;
  const int y = rr / w, x = rr - y * w;
  const float gx = norm_coord(x, w), gy = norm_coord(y, h);
  const float* ghg = grad_hg + ((int64_t)b * chan_total + chan_off + (int64_t)g * C) * hw + rr;
  float2 gm = make_float2(0.f, 0.f);
  if (live && grad_motions != nullptr) gm = __ldg(reinterpret_cast<const float2*>(grad_motions) + (int64_t)bg * hw + rr);
  float* gsrc = grad_source ? grad_source + (int64_t)b * C * hw : nullptr;
  const float* src = source + (int64_t)b * C * hw;
  const float sx_pix = to_pixel_grad<MRFA_COORD_NORM_ACT>(w), sy_pix = to_pixel_grad<MRFA_COORD_NORM_ACT>(h);

  if (g == 0) {
    float v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) v[i] = 0.f;
    if (live) {
      const float2 m = bg_param ? bg_affine_fwd(bg_param + 9 * b, gx, gy) : make_float2(gx, gy);
      const TapsG t = make_taps_g(to_pixel<MRFA_COORD_NORM_ACT>(m.x, w), to_pixel<MRFA_COORD_NORM_ACT>(m.y, h), h, w);
      const float2 gp = sample_source_bwd(src, gsrc, ghg, C, hw, t);
      gm.x = fmaf(gp.x, sx_pix, gm.x);
      gm.y = fmaf(gp.y, sy_pix, gm.y);
      if (bg_param) bg_affine_bwd(bg_param + 9 * b, gx, gy, gm, v);
    }
    if (bg_param) block_accumulate<9>(v, acc + (int64_t)bg * kTpsAccStride);
    return;
  }

  const int sys = b * G + (g - 1);
  const float* th = theta + (int64_t)sys * 6;
  float u[5], du[5], dx[5], dy[5];
  float ox = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 0), gx), __fmul_rn(__ldg(th + 1), gy)), __ldg(th + 2));
  float oy = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 3), gx), __fmul_rn(__ldg(th + 4), gy)), __ldg(th + 5));
  float rx = 0.f, ry = 0.f;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float* cp = kp_d + ((int64_t)sys * 5 + k) * 2;
    dx[k] = __fsub_rn(gx, __ldg(cp)); dy[k] = __fsub_rn(gy, __ldg(cp + 1));
    const float d2 = __fadd_rn(__fmul_rn(dx[k], dx[k]), __fmul_rn(dy[k], dy[k]));
    const float lg = logf(__fadd_rn(d2, 1e-9f));
    u[k] = __fmul_rn(d2, lg);
    du[k] = lg + d2 / (d2 + 1e-9f);                        // dU/d(d2)
    const float* cw = control_params + ((int64_t)sys * 5 + k) * 2;
    rx = fmaf(u[k], __ldg(cw), rx);
    ry = fmaf(u[k], __ldg(cw + 1), ry);
  }
  float v[26];
#pragma unroll
  for (int i = 0; i < 26; ++i) v[i] = 0.f;
  if (live) {
    const float2 m = make_float2(__fadd_rn(ox, rx), __fadd_rn(oy, ry));
    const TapsG t = make_taps_g(to_pixel<MRFA_COORD_NORM_ACT>(m.x, w), to_pixel<MRFA_COORD_NORM_ACT>(m.y, h), h, w);
    const float2 gp = sample_source_bwd(src, gsrc, ghg, C, hw, t);
    gm.x = fmaf(gp.x, sx_pix, gm.x);
    gm.y = fmaf(gp.y, sy_pix, gm.y);
    v[0] = gm.x * gx; v[1] = gm.x * gy; v[2] = gm.x;
    v[3] = gm.y * gx; v[4] = gm.y * gy; v[5] = gm.y;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float* cw = control_params + ((int64_t)sys * 5 + k) * 2;
      v[6 + 2 * k] = u[k] * gm.x;
      v[7 + 2 * k] = u[k] * gm.y;
      // d(d2)/d(c_k) = -2 (p - c_k)
      const float s = -2.f * du[k] * (gm.x * __ldg(cw) + gm.y * __ldg(cw + 1));
      v[16 + 2 * k] = s * dx[k];
      v[17 + 2 * k] = s * dy[k];
    }
  }
  block_accumulate<26>(v, acc + (int64_t)bg * kTpsAccStride);
}

// Adjoint of tps_solve_kernel, one warp per (b, g) system, fp64.  param = inv(L) Y with L symmetric, so
// dY = inv(L) dparam (the same elimination on a new right-hand side) and dL = -dY param^T.
//   L = [[Kmat, P], [P^T, 0]] + 0.01 I,  Kmat_rj = U(|p_r - p_j|^2),  P_r = (x_r, y_r, 1),  Y = [kp_2; 0].
// Inputs: accumulator slice of tps_motion_prior_bwd_kernel (d theta, d control_params, direct d control points).
// Outputs: grad_kp_1 += (direct + through L), grad_kp_2 += dY[:5]  (both caller-zeroed or pre-filled with the
// heat-map gradients).
__global__ void __launch_bounds__(128)
tps_solve_bwd_kernel(const float* __restrict__ acc, const float* __restrict__ kp_1, const float* __restrict__ theta,
                     const float* __restrict__ control_params, float* __restrict__ grad_kp_1,
                     float* __restrict__ grad_kp_2, float* __restrict__ grad_bg, int B, int G) {
  const int sysg = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32);   // over B * (G + 1)
  const int lane = threadIdx.x % 32;
  if (sysg >= B * (G + 1)) return;
  const int b = sysg / (G + 1), g = sysg - b * (G + 1);
  const float* a = acc + (int64_t)sysg * kTpsAccStride;
  if (g == 0) {
    if (grad_bg != nullptr && lane < 9) grad_bg[9 * b + lane] = a[lane];
    return;
  }
  const int sys = b * G + (g - 1);
  constexpr int n = 5, m = 8;
  const float* p1 = kp_1 + (int64_t)sys * n * 2;
  double row[m + 2];
#pragma unroll
  for (int j = 0; j < m + 2; ++j) row[j] = 0.0;
  const int r = lane;
  if (r < n) {
    const float xr = __ldg(p1 + 2 * r), yr = __ldg(p1 + 2 * r + 1);
#pragma unroll
    for (int j = 0; j < n; ++j) {
      const float dx = __fsub_rn(xr, __ldg(p1 + 2 * j)), dy = __fsub_rn(yr, __ldg(p1 + 2 * j + 1));
      float d = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      d = __fmul_rn(d, d);
      row[j] = (double)__fmul_rn(d, logf(__fadd_rn(d, 1e-9f)));
    }
    row[n] = xr; row[n + 1] = yr; row[n + 2] = 1.0;
    // d param rows 0..4 = d control_params
    row[m] = a[6 + 2 * r]; row[m + 1] = a[7 + 2 * r];
  } else if (r < m) {
    const int c = r - n;
#pragma unroll
    for (int j = 0; j < n; ++j) row[j] = (c == 0) ? __ldg(p1 + 2 * j) : (c == 1) ? __ldg(p1 + 2 * j + 1) : 1.f;
    // d param rows 5..7 = d theta^T: param[n + c, i] = theta[i, c]
    row[m] = a[c]; row[m + 1] = a[3 + c];
  }
  if (r < m) {
#pragma unroll
    for (int j = 0; j < m; ++j) if (j == r) row[j] += (double)0.01f;
  }
  int my_col = -1;
#pragma unroll
  for (int col = 0; col < m; ++col) {
    double cand = (r < m && my_col < 0) ? fabs(row[col]) : -1.0;
    int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double oc = __shfl_xor_sync(0xffffffffu, cand, o);
      const int ow = __shfl_xor_sync(0xffffffffu, who, o);
      if (oc > cand || (oc == cand && ow < who)) { cand = oc; who = ow; }
    }
    const double pivot = __shfl_sync(0xffffffffu, row[col], who);
    const double f = (lane == who) ? 0.0 : row[col] / pivot;
#pragma unroll
    for (int j = 0; j < m + 2; ++j) {
      const double pv = __shfl_sync(0xffffffffu, row[j], who);
      if (lane == who) row[j] = pv / pivot;
      else row[j] -= f * pv;
    }
    if (lane == who) my_col = col;
  }
  // dY row my_col lives in this lane; gather the 8 x 2 matrix into every lane, indexed by row
  double dY[m][2];
#pragma unroll
  for (int q = 0; q < m; ++q) { dY[q][0] = 0.0; dY[q][1] = 0.0; }
#pragma unroll
  for (int src = 0; src < m; ++src) {
    const int c = __shfl_sync(0xffffffffu, my_col, src);
    const double y0 = __shfl_sync(0xffffffffu, row[m], src), y1 = __shfl_sync(0xffffffffu, row[m + 1], src);
#pragma unroll
    for (int q = 0; q < m; ++q)
      if (q == c) { dY[q][0] = y0; dY[q][1] = y1; }
  }
  if (r < n) {
    // param rows: control weights W (5 x 2) and theta^T (3 x 2)
    double W[n][2], T[3][2];
#pragma unroll
    for (int j = 0; j < n; ++j) {
      W[j][0] = __ldg(control_params + ((int64_t)sys * n + j) * 2);
      W[j][1] = __ldg(control_params + ((int64_t)sys * n + j) * 2 + 1);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      T[c][0] = __ldg(theta + (int64_t)sys * 6 + c);
      T[c][1] = __ldg(theta + (int64_t)sys * 6 + 3 + c);
    }
    // dL[i][j] = -(dY[i] . param[j]).  Control point r enters L through Kmat row r / column r and P row r / column r.
    double gxr = 0.0, gyr = 0.0;
    double dYr0 = 0.0, dYr1 = 0.0, Wr0 = 0.0, Wr1 = 0.0;
#pragma unroll
    for (int q = 0; q < n; ++q)
      if (q == r) { dYr0 = dY[q][0]; dYr1 = dY[q][1]; Wr0 = W[q][0]; Wr1 = W[q][1]; }
    const float xr = __ldg(p1 + 2 * r), yr = __ldg(p1 + 2 * r + 1);
#pragma unroll
    for (int j = 0; j < n; ++j) {
      // dK_rj + dK_jr
      const double dk = -(dYr0 * W[j][0] + dYr1 * W[j][1]) - (dY[j][0] * Wr0 + dY[j][1] * Wr1);
      const double ddx = (double)xr - (double)__ldg(p1 + 2 * j), ddy = (double)yr - (double)__ldg(p1 + 2 * j + 1);
      const double d2 = ddx * ddx + ddy * ddy;
      const double du = (j == r) ? 0.0 : log(d2 + 1e-9) + d2 / (d2 + 1e-9);
      gxr += dk * du * 2.0 * ddx;
      gyr += dk * du * 2.0 * ddy;
    }
    // P block: L[r][n + c] = P_rc and L[n + c][r] = P_rc for c = 0 (x), 1 (y)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double dp = -(dYr0 * T[c][0] + dYr1 * T[c][1]) - (dY[n + c][0] * Wr0 + dY[n + c][1] * Wr1);
      if (c == 0) gxr += dp; else gyr += dp;
    }
    grad_kp_1[((int64_t)sys * n + r) * 2 + 0] += (float)(gxr + (double)a[16 + 2 * r]);
    grad_kp_1[((int64_t)sys * n + r) * 2 + 1] += (float)(gyr + (double)a[17 + 2 * r]);
    grad_kp_2[((int64_t)sys * n + r) * 2 + 0] += (float)dYr0;
    grad_kp_2[((int64_t)sys * n + r) * 2 + 1] += (float)dYr1;
  }
}

}  // namespace mrfa

using namespace mrfa;

extern "C" int64_t mrfa_dense_motion_prior_bwd_workspace(int B, int K) { return (int64_t)B * (K + 1) * kAccStride; }
extern "C" int64_t mrfa_tps_motion_prior_bwd_workspace(int B, int G) { return (int64_t)B * (G + 1) * kTpsAccStride; }

extern "C" int mrfa_dense_motion_prior_bwd(const float* grad_motions, const float* grad_hg, const float* kp_d,
                                           const float* kp_s, const float* jac_d, const float* jac_s,
                                           const float* bg_param, const float* source, float* workspace,
                                           float* grad_kp_d, float* grad_kp_s, float* grad_jac_d, float* grad_jac_s,
                                           float* grad_bg, float* grad_source, int B, int K, int C, int h, int w,
                                           float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_hg && kp_d && kp_s && source && workspace && grad_kp_d && grad_kp_s);
  MRFA_CHECK_ARG((jac_d == nullptr) == (jac_s == nullptr));
  MRFA_CHECK_ARG(jac_d == nullptr || (grad_jac_d != nullptr && grad_jac_s != nullptr));
  MRFA_CHECK_ARG(grad_bg == nullptr || bg_param != nullptr);
  MRFA_CHECK_ARG(B >= 0 && K > 0 && C > 0 && h > 1 && w > 1 && variance > 0.f);
  if (B == 0) return 0;
  MRFA_CHECK_SHAPE((int64_t)B * (K + 1) <= 65535 && (int64_t)h * w < ((int64_t)1 << 31));
  cudaStream_t st = as_stream(stream);
  const dim3 grid((unsigned)cdiv64((int64_t)h * w, 256), (unsigned)(B * (K + 1)));
  dense_motion_prior_bwd_kernel<<<grid, 256, 0, st>>>(grad_motions, grad_hg, kp_d, kp_s, jac_d, jac_s, bg_param, source,
                                                      workspace, grad_source, B, K, C, h, w, variance);
  int rc = MRFA_LAUNCH_RESULT();
  if (rc) return rc;
  dense_motion_prior_bwd_finish_kernel<<<(unsigned)cdiv64((int64_t)B * (K + 1), 128), 128, 0, st>>>(
      workspace, jac_d, jac_s, grad_kp_d, grad_kp_s, grad_jac_d, grad_jac_s, grad_bg, B, K);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_kp2gaussian_bwd(const float* grad, const float* kp, float* grad_kp, int P, int h, int w,
                                    float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad && kp && grad_kp && P >= 0 && h > 1 && w > 1 && variance > 0.f);
  if (P == 0) return 0;
  MRFA_CHECK_SHAPE(P <= 65535);
  const dim3 grid((unsigned)cdiv64((int64_t)h * w, 256), (unsigned)P);
  kp2gaussian_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(grad, kp, nullptr, grad_kp, nullptr, P, 0, 0, h, w, variance);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_tps_motion_prior_bwd(const float* grad_motions, const float* grad_hg, const float* kp_d,
                                         const float* kp_s, const float* theta, const float* control_params,
                                         const float* bg_param, const float* source, float* workspace,
                                         float* grad_kp_d, float* grad_kp_s, float* grad_bg, float* grad_source,
                                         int B, int G, int C, int h, int w, float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_hg && kp_d && kp_s && theta && control_params && source && workspace && grad_kp_d && grad_kp_s);
  MRFA_CHECK_ARG(grad_bg == nullptr || bg_param != nullptr);
  MRFA_CHECK_ARG(B >= 0 && G > 0 && C > 0 && h > 1 && w > 1 && variance > 0.f);
  if (B == 0) return 0;
  const int KP = G * 5;
  MRFA_CHECK_SHAPE((int64_t)B * KP <= 65535 && (int64_t)h * w < ((int64_t)1 << 31));
  cudaStream_t st = as_stream(stream);
  const int chan_total = (KP + 1) + (G + 1) * C;
  const unsigned gx = (unsigned)cdiv64((int64_t)h * w, 256);
  // heat-map channels 1..KP: d kp_d, d kp_s (channel 0 is the constant background map)
  kp2gaussian_bwd_kernel<<<dim3(gx, (unsigned)(B * KP)), 256, 0, st>>>(grad_hg, kp_d, kp_s, grad_kp_d, grad_kp_s, KP,
                                                                      (int64_t)chan_total * h * w, 1, h, w, variance);
  int rc = MRFA_LAUNCH_RESULT();
  if (rc) return rc;
  tps_motion_prior_bwd_kernel<<<dim3(gx, (unsigned)(B * (G + 1))), 256, 0, st>>>(
      grad_motions, grad_hg, kp_d, theta, control_params, bg_param, source, workspace, grad_source, B, G, C, chan_total,
      KP + 1, h, w);
  rc = MRFA_LAUNCH_RESULT();
  if (rc) return rc;
  tps_solve_bwd_kernel<<<(unsigned)cdiv64((int64_t)B * (G + 1) * 32, 128), 128, 0, st>>>(
      workspace, kp_d, theta, control_params, grad_kp_d, grad_kp_s, grad_bg, B, G);
  return MRFA_LAUNCH_RESULT();
}
