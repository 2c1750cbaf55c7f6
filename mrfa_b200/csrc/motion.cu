// Prior dense-motion synthesis (K6/K7/K8 of SURVEY.md, a12-a15): key-point heat-maps, the
// K+1 sparse motion fields (affine / Jacobian / background / thin-plate spline) and the K+1
// warped copies of the quarter-resolution source, fused so that neither the identity grid,
// the (B,K,h,w,2,2) Jacobian repeat nor the (K+1)x repeated source of the reference is ever
// materialised.  HBM-bound; algorithmic bytes per pair = outputs (motions + hourglass input)
// + one read of the source.
#include "common.cuh"

namespace mrfa {

__device__ __forceinline__ float gauss(float gx, float gy, const float* __restrict__ kp, float variance) {
  const float dx = __fsub_rn(gx, __ldg(kp)), dy = __fsub_rn(gy, __ldg(kp + 1));
  const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  return expf(__fdiv_rn(__fmul_rn(-0.5f, s), variance));
}

// to_homogeneous -> 3x3 matmul -> from_homogeneous (util.py:329-338, dense_motion.py:69-73)
__device__ __forceinline__ float2 bg_affine(const float* __restrict__ P, float gx, float gy) {
  float t[3];
#pragma unroll
  for (int r = 0; r < 3; ++r)
    t[r] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(P + 3 * r), gx), __fmul_rn(__ldg(P + 3 * r + 1), gy)), __ldg(P + 3 * r + 2));
  return make_float2(__fdiv_rn(t[0], t[2]), __fdiv_rn(t[1], t[2]));
}

template <int MODE>
__device__ __forceinline__ void sample_source(const float* __restrict__ src, float* __restrict__ dst, int C, int h,
                                              int w, float2 m, int r) {
  const Taps t = make_taps(to_pixel<MODE>(m.x, w), to_pixel<MODE>(m.y, h), h, w);
  const int hw = h * w;
  for (int c = 0; c < C; ++c) {
    const float* s = src + (int64_t)c * hw;
    float acc = __ldg(s + t.o_nw) * t.w_nw;
    acc = fmaf(__ldg(s + t.o_ne), t.w_ne, acc);
    acc = fmaf(__ldg(s + t.o_sw), t.w_sw, acc);
    acc = fmaf(__ldg(s + t.o_se), t.w_se, acc);
    dst[(int64_t)c * hw + r] = acc;
  }
}

// One thread per (b, y, x) walking k = 0 .. K (k == 0 is the background channel); a block covers 128 pixels of ONE sample.
// The per-key-point algebra -- J = jac_s * inverse(jac_d) with its four IEEE divisions, the key-point coordinates -- is
// evaluated once per block into shared memory (thread k owns key-point k); the identity-grid value of the pixel, its
// two IEEE divisions and the reciprocal of the variance are evaluated once per thread instead of once per (k, pixel)
// (round 2: the one-thread-per-(b,k,pixel) form issued 425 instructions per thread, 60 % issue-bound at 3 % of DRAM).
struct PriorKp {
  float kdx, kdy, ksx, ksy, j00, j01, j10, j11;
};
constexpr int kPriorThreads = 128;
constexpr int kPriorMaxK = 64;                       // key-points held in shared memory per sample

template <int CT>                                    // CT > 0: compile-time channel count (3), 0: runtime C
__global__ void __launch_bounds__(kPriorThreads)
dense_motion_prior_kernel(const float* __restrict__ kp_d, const float* __restrict__ kp_s,
                          const float* __restrict__ jac_d, const float* __restrict__ jac_s,
                          const float* __restrict__ bg_param, const float* __restrict__ source,
                          float* __restrict__ motions, float* __restrict__ hg_input,
                          int B, int K, int Crt, int h, int w, float variance) {
  __shared__ PriorKp sk[kPriorMaxK];
  const int C = CT > 0 ? CT : Crt;
  const int hw = h * w;
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < K; k += kPriorThreads) {
    const int kk = b * K + k;
    PriorKp p;
    p.kdx = __ldg(kp_d + 2 * kk); p.kdy = __ldg(kp_d + 2 * kk + 1);
    p.ksx = __ldg(kp_s + 2 * kk); p.ksy = __ldg(kp_s + 2 * kk + 1);
    p.j00 = 1.f; p.j01 = 0.f; p.j10 = 0.f; p.j11 = 1.f;
    if (jac_d != nullptr) {
      // J = jac_s * inverse(jac_d)  (dense_motion.py:54)
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
    sk[k] = p;
  }
  __syncthreads();
  const int r = blockIdx.x * kPriorThreads + threadIdx.x;
  if (r >= hw) return;
  const int y = r / w, x = r - y * w;
  const float gx = norm_coord(x, w), gy = norm_coord(y, h);
  const float rvar = __frcp_rn(variance);
  const float* src = source + (int64_t)b * C * hw;
  float2* mo = reinterpret_cast<float2*>(motions) + (int64_t)b * (K + 1) * hw + r;
  float* dst = hg_input + (int64_t)b * (K + 1) * (C + 1) * hw + r;

#pragma unroll 2
  for (int k = 0; k <= K; ++k) {
    float heat = 0.f;
    float2 m;
    if (k == 0) {
      m = bg_param ? bg_affine(bg_param + 9 * b, gx, gy) : make_float2(gx, gy);
    } else {
      const PriorKp p = sk[k - 1];
      float cx = __fsub_rn(gx, p.kdx), cy = __fsub_rn(gy, p.kdy);
      {
        const float sx = __fsub_rn(gx, p.ksx), sy = __fsub_rn(gy, p.ksy);
        const float dd = __fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), ds = __fadd_rn(__fmul_rn(sx, sx), __fmul_rn(sy, sy));
        // exp((-0.5 * s) / variance), util.py:85: the quotient through div_by_const (same bits as the IEEE division)
        heat = __fsub_rn(expf(div_by_const(__fmul_rn(-0.5f, dd), variance, rvar)),
                         expf(div_by_const(__fmul_rn(-0.5f, ds), variance, rvar)));
      }
      if (jac_d != nullptr) {
        const float nx = __fadd_rn(__fmul_rn(p.j00, cx), __fmul_rn(p.j01, cy));
        const float ny = __fadd_rn(__fmul_rn(p.j10, cx), __fmul_rn(p.j11, cy));
        cx = nx; cy = ny;
      }
      m = make_float2(__fadd_rn(cx, p.ksx), __fadd_rn(cy, p.ksy));
    }
    mo[(int64_t)k * hw] = m;
    float* dk = dst + (int64_t)k * (C + 1) * hw;
    dk[0] = heat;
    const Taps t = make_taps(to_pixel<MRFA_COORD_NORM_ACF>(m.x, w), to_pixel<MRFA_COORD_NORM_ACF>(m.y, h), h, w);
#pragma unroll
    for (int c = 0; c < (CT > 0 ? CT : 1); ++c) {
      if (CT > 0) {
        const float* s = src + c * hw;
        float acc = __ldg(s + t.o_nw) * t.w_nw;
        acc = fmaf(__ldg(s + t.o_ne), t.w_ne, acc);
        acc = fmaf(__ldg(s + t.o_sw), t.w_sw, acc);
        acc = fmaf(__ldg(s + t.o_se), t.w_se, acc);
        dk[(c + 1) * hw] = acc;
      }
    }
    if (CT == 0) {
      for (int c = 0; c < C; ++c) {
        const float* s = src + (int64_t)c * hw;
        float acc = __ldg(s + t.o_nw) * t.w_nw;
        acc = fmaf(__ldg(s + t.o_ne), t.w_ne, acc);
        acc = fmaf(__ldg(s + t.o_sw), t.w_sw, acc);
        acc = fmaf(__ldg(s + t.o_se), t.w_se, acc);
        dk[(int64_t)(c + 1) * hw] = acc;
      }
    }
  }
}

// ---- thin-plate splines ---------------------------------------------------------------------
// One warp per (b,g) system: lane r < 8 owns row r of [L | Y] (8 + 2 columns).  Gauss-Jordan
// with partial pivoting through warp shuffles, in fp64 (the 8x8 systems are conditioned at
// ~1e3-1e4, so an fp32 elimination would carry ~1e-4 of its own round-off; fp64 keeps this
// kernel's contribution below fp32 resolution).
__global__ void __launch_bounds__(128)
tps_solve_kernel(const float* __restrict__ kp_1, const float* __restrict__ kp_2, float* __restrict__ theta,
                 float* __restrict__ control_params, int BG) {
  const int sys = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) / 32);
  const int lane = threadIdx.x % 32;
  if (sys >= BG) return;
  constexpr int n = 5, m = 8;
  const float* p1 = kp_1 + (int64_t)sys * n * 2;
  const float* p2 = kp_2 + (int64_t)sys * n * 2;
  double row[m + 2];
#pragma unroll
  for (int j = 0; j < m + 2; ++j) row[j] = 0.0;
  const int r = lane;
  if (r < n) {
    const float xr = __ldg(p1 + 2 * r), yr = __ldg(p1 + 2 * r + 1);
#pragma unroll
    for (int j = 0; j < n; ++j) {
      // K = |p_r - p_j|^2 * log(|p_r - p_j|^2 + 1e-9), evaluated in fp32 like util.py:362-364
      const float dx = __fsub_rn(xr, __ldg(p1 + 2 * j)), dy = __fsub_rn(yr, __ldg(p1 + 2 * j + 1));
      float d = sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
      d = __fmul_rn(d, d);
      row[j] = (double)__fmul_rn(d, logf(__fadd_rn(d, 1e-9f)));
    }
    row[n] = xr; row[n + 1] = yr; row[n + 2] = 1.0;
    row[m] = __ldg(p2 + 2 * r); row[m + 1] = __ldg(p2 + 2 * r + 1);
  } else if (r < m) {
    const int c = r - n;              // rows of [P^T 0]: x's, y's, ones
#pragma unroll
    for (int j = 0; j < n; ++j) row[j] = (c == 0) ? __ldg(p1 + 2 * j) : (c == 1) ? __ldg(p1 + 2 * j + 1) : 1.f;
  }
  if (r < m) {
#pragma unroll
    for (int j = 0; j < m; ++j) if (j == r) row[j] += (double)0.01f;
  }
  int my_col = -1;                     // pivot column this lane ended up owning
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
  if (r < m && my_col >= 0) {
    // param row my_col: rows [0,n) are the control weights, rows [n,n+3) are theta^T
    if (my_col < n) {
      control_params[((int64_t)sys * n + my_col) * 2 + 0] = (float)row[m];
      control_params[((int64_t)sys * n + my_col) * 2 + 1] = (float)row[m + 1];
    } else {
      theta[(int64_t)sys * 6 + 0 * 3 + (my_col - n)] = (float)row[m];
      theta[(int64_t)sys * 6 + 1 * 3 + (my_col - n)] = (float)row[m + 1];
    }
  }
}

__global__ void __launch_bounds__(256)
tps_heatmap_kernel(const float* __restrict__ kp_d, const float* __restrict__ kp_s, float* __restrict__ hg_input,
                   int B, int KP, int chan_total, int h, int w, float variance) {
  const int hw = h * w;
  const int64_t total = (int64_t)B * (KP + 1) * hw;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = (int)(i % hw);
  const int k = (int)((i / hw) % (KP + 1));
  const int b = (int)(i / ((int64_t)hw * (KP + 1)));
  const int y = r / w, x = r - y * w;
  float heat = 0.f;
  if (k > 0) {
    const float gx = norm_coord(x, w), gy = norm_coord(y, h);
    const int kk = b * KP + (k - 1);
    heat = __fsub_rn(gauss(gx, gy, kp_d + 2 * kk, variance), gauss(gx, gy, kp_s + 2 * kk, variance));
  }
  hg_input[((int64_t)b * chan_total + k) * hw + r] = heat;
}

__global__ void __launch_bounds__(256)
tps_motion_prior_kernel(const float* __restrict__ kp_d, const float* __restrict__ theta,
                        const float* __restrict__ control_params, const float* __restrict__ bg_param,
                        const float* __restrict__ source, float* __restrict__ motions, float* __restrict__ hg_input,
                        int B, int G, int C, int chan_total, int chan_off, int h, int w) {
  const int hw = h * w;
  const int64_t total = (int64_t)B * (G + 1) * hw;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = (int)(i % hw);
  const int g = (int)((i / hw) % (G + 1));
  const int b = (int)(i / ((int64_t)hw * (G + 1)));
  const int y = r / w, x = r - y * w;
  const float gx = norm_coord(x, w), gy = norm_coord(y, h);
  float2 m;
  if (g == 0) {
    m = bg_param ? bg_affine(bg_param + 9 * b, gx, gy) : make_float2(gx, gy);
  } else {
    const int sys = b * G + (g - 1);
    const float* th = theta + (int64_t)sys * 6;
    // theta[:, :2] @ p + theta[:, 2]   (util.py:402)
    float ox = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 0), gx), __fmul_rn(__ldg(th + 1), gy)), __ldg(th + 2));
    float oy = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 3), gx), __fmul_rn(__ldg(th + 4), gy)), __ldg(th + 5));
    float rx = 0.f, ry = 0.f;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const float* cp = kp_d + ((int64_t)sys * 5 + k) * 2;
      const float dx = __fsub_rn(gx, __ldg(cp)), dy = __fsub_rn(gy, __ldg(cp + 1));
      const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
      const float u = __fmul_rn(d2, logf(__fadd_rn(d2, 1e-9f)));
      const float* cw = control_params + ((int64_t)sys * 5 + k) * 2;
      rx = fmaf(u, __ldg(cw), rx);
      ry = fmaf(u, __ldg(cw + 1), ry);
    }
    m = make_float2(__fadd_rn(ox, rx), __fadd_rn(oy, ry));
  }
  reinterpret_cast<float2*>(motions)[i] = m;
  float* dst = hg_input + ((int64_t)b * chan_total + chan_off + (int64_t)g * C) * hw;
  sample_source<MRFA_COORD_NORM_ACT>(source + (int64_t)b * C * hw, dst, C, h, w, m, r);
}

// Random affine + thin-plate warp of the identity grid: the equivariance branch of training
// (SURVEY.md 8(f) N4).  metric 0 = Transform.warp_coordinates (model.py:50-70): d = |dx| + |dy|,
// U = d^2 * log(d + 1e-6);  metric 1 = TPS mode 'random' (util.py:412-423): r2 = dx^2 + dy^2,
// U = r2 * log(r2 + 1e-9).  The reference materialises (B, HW, P, 2) temporaries; here one
// thread per grid point walks the P control points held in shared memory.
constexpr int kMaxCtrl = 256;

__global__ void __launch_bounds__(256)
random_warp_grid_kernel(const float* __restrict__ theta, const float* __restrict__ control_points,
                        const float* __restrict__ control_params, float* __restrict__ grid, int P, int h, int w,
                        int metric) {
  __shared__ float2 cp[kMaxCtrl];
  __shared__ float cw[kMaxCtrl];
  const int b = blockIdx.y;
  for (int k = threadIdx.x; k < P; k += blockDim.x) {
    cp[k] = make_float2(__ldg(control_points + 2 * k), __ldg(control_points + 2 * k + 1));
    cw[k] = control_params ? __ldg(control_params + (int64_t)b * P + k) : 0.f;
  }
  __syncthreads();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= h * w) return;
  const int y = r / w, x = r - y * w;
  const float gx = norm_coord(x, w), gy = norm_coord(y, h);
  const float* th = theta + (int64_t)b * 6;
  float ox = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 0), gx), __fmul_rn(__ldg(th + 1), gy)), __ldg(th + 2));
  float oy = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(th + 3), gx), __fmul_rn(__ldg(th + 4), gy)), __ldg(th + 5));
  if (control_params != nullptr) {
    float acc = 0.f;
    for (int k = 0; k < P; ++k) {
      const float dx = __fsub_rn(gx, cp[k].x), dy = __fsub_rn(gy, cp[k].y);
      float u;
      if (metric == 0) {
        const float d = __fadd_rn(fabsf(dx), fabsf(dy));
        u = __fmul_rn(__fmul_rn(d, d), logf(__fadd_rn(d, 1e-6f)));
      } else {
        const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        u = __fmul_rn(d2, logf(__fadd_rn(d2, 1e-9f)));
      }
      acc = __fadd_rn(acc, __fmul_rn(u, cw[k]));
    }
    ox = __fadd_rn(ox, acc);
    oy = __fadd_rn(oy, acc);
  }
  reinterpret_cast<float2*>(grid)[(int64_t)b * h * w + r] = make_float2(ox, oy);
}


}  // namespace mrfa

using namespace mrfa;

extern "C" int mrfa_dense_motion_prior(const float* kp_d, const float* kp_s, const float* jac_d, const float* jac_s,
                                       const float* bg_param, const float* source, float* motions, float* hg_input,
                                       int B, int K, int C, int h, int w, float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(kp_d && kp_s && source && motions && hg_input);
  MRFA_CHECK_ARG((jac_d == nullptr) == (jac_s == nullptr));
  MRFA_CHECK_ARG(B >= 0 && K > 0 && C > 0 && h > 1 && w > 1 && variance > 0.f);
  if (B == 0) return 0;
  MRFA_CHECK_SHAPE(B <= 65535 && K <= kPriorMaxK && (int64_t)h * w * (C + 1) < ((int64_t)1 << 31));
  const dim3 grid((unsigned)cdiv64((int64_t)h * w, kPriorThreads), (unsigned)B);
  if (C == 3)
    dense_motion_prior_kernel<3><<<grid, kPriorThreads, 0, as_stream(stream)>>>(
        kp_d, kp_s, jac_d, jac_s, bg_param, source, motions, hg_input, B, K, C, h, w, variance);
  else
    dense_motion_prior_kernel<0><<<grid, kPriorThreads, 0, as_stream(stream)>>>(
        kp_d, kp_s, jac_d, jac_s, bg_param, source, motions, hg_input, B, K, C, h, w, variance);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_tps_solve(const float* kp_1, const float* kp_2, float* theta, float* control_params, int BG,
                              mrfa_stream_t stream) {
  MRFA_CHECK_ARG(kp_1 && kp_2 && theta && control_params && BG >= 0);
  if (BG == 0) return 0;
  tps_solve_kernel<<<(unsigned)cdiv64((int64_t)BG * 32, 128), 128, 0, as_stream(stream)>>>(kp_1, kp_2, theta, control_params, BG);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_tps_motion_prior(const float* kp_d, const float* kp_s, const float* theta,
                                     const float* control_params, const float* bg_param, const float* source,
                                     float* motions, float* hg_input, int B, int G, int C, int h, int w,
                                     float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(kp_d && kp_s && theta && control_params && source && motions && hg_input);
  MRFA_CHECK_ARG(B >= 0 && G > 0 && C > 0 && h > 1 && w > 1 && variance > 0.f);
  if (B == 0) return 0;
  const int KP = G * 5;
  const int chan_total = (KP + 1) + (G + 1) * C;
  const int64_t t1 = (int64_t)B * (KP + 1) * h * w;
  tps_heatmap_kernel<<<(unsigned)cdiv64(t1, 256), 256, 0, as_stream(stream)>>>(kp_d, kp_s, hg_input, B, KP, chan_total, h, w, variance);
  int rc = MRFA_LAUNCH_RESULT();
  if (rc) return rc;
  const int64_t t2 = (int64_t)B * (G + 1) * h * w;
  tps_motion_prior_kernel<<<(unsigned)cdiv64(t2, 256), 256, 0, as_stream(stream)>>>(
      kp_d, theta, control_params, bg_param, source, motions, hg_input, B, G, C, chan_total, KP + 1, h, w);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_random_warp_grid(const float* theta, const float* control_points, const float* control_params,
                                     float* grid, int B, int P, int h, int w, int metric, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(theta && grid && B >= 0 && h > 0 && w > 0 && (metric == 0 || metric == 1));
  MRFA_CHECK_ARG(control_params == nullptr || (control_points != nullptr && P > 0));
  MRFA_CHECK_SHAPE(P <= kMaxCtrl && B <= 65535);
  if (B == 0) return 0;
  dim3 g((unsigned)cdiv64((int64_t)h * w, 256), (unsigned)B);
  random_warp_grid_kernel<<<g, 256, 0, as_stream(stream)>>>(theta, control_points, control_params, grid,
                                                            control_params ? P : 0, h, w, metric);
  return MRFA_LAUNCH_RESULT();
}
