// Bilinear warps (K5/K6 of SURVEY.md): forward, backward and the fused dual warp.
//
// Memory-bound gathers over NCHW fp32 planes.  One thread owns one output pixel and a chunk
// of CPT channels: the four tap offsets/weights are computed once and reused for every
// channel of the chunk, the 4*CPT tap loads are issued back to back (memory-level
// parallelism), and lanes map to consecutive output pixels so both the taps (smooth motion
// fields) and the stores are coalesced.  Roofline: HBM, algorithmic bytes per warp call =
// (2*C*Ho*Wo + 2*Ho*Wo) * 4.
#include "common.cuh"

namespace mrfa {

constexpr int kWarpPix = 128;          // pixels per block (4 warps along the pixel axis)
constexpr int kThreads = 256;          // 2 channel-chunk slots per block
constexpr int kCPT = 8;                // channels per thread

template <int MODE, int PAD, bool ADD_ID>
__device__ __forceinline__ void load_sample_point(const float* __restrict__ grid, const mrfa_grid_strides_t& gs,
                                                  int n, int y, int x, int H, int W,
                                                  float& ix, float& iy, float& mx, float& my) {
  const float* g = grid + n * gs.sn + y * gs.sy + x * gs.sx;
  float gx, gy;
  if (gs.sc == 1 && ((reinterpret_cast<uintptr_t>(g) & 7) == 0)) {
    float2 v = __ldg(reinterpret_cast<const float2*>(g));
    gx = v.x; gy = v.y;
  } else {
    gx = __ldg(g); gy = __ldg(g + gs.sc);
  }
  if (ADD_ID) { gx = __fadd_rn(gx, (float)x); gy = __fadd_rn(gy, (float)y); }
  ix = source_index<MODE, PAD>(gx, W, &mx);
  iy = source_index<MODE, PAD>(gy, H, &my);
}

template <int MODE, int PAD, bool ADD_ID>
__global__ void __launch_bounds__(kThreads)
grid_sample_fwd_kernel(const float* __restrict__ in, const float* __restrict__ grid, mrfa_grid_strides_t gs,
                       float* __restrict__ out, int N, int C, int H, int W, int Ho, int Wo, int in_batch_div) {
  const int HoWo = Ho * Wo;
  const int64_t gp = (int64_t)blockIdx.x * kWarpPix + (threadIdx.x % kWarpPix);
  if (gp >= (int64_t)N * HoWo) return;
  const int n = (int)(gp / HoWo);
  const int p = (int)(gp - (int64_t)n * HoWo);
  const int y = p / Wo, x = p - y * Wo;
  const int c0 = (blockIdx.y * (kThreads / kWarpPix) + threadIdx.x / kWarpPix) * kCPT;
  if (c0 >= C) return;

  float ix, iy, mx, my;
  load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
  const Taps t = make_taps(ix, iy, H, W);

  const int64_t HW = (int64_t)H * W;
  const float* src = in + ((int64_t)(n / in_batch_div) * C + c0) * HW;
  float* dst = out + ((int64_t)n * C + c0) * HoWo + p;
  const int nc = min(kCPT, C - c0);
  if (nc == kCPT) {
    float v[kCPT][4];
#pragma unroll
    for (int c = 0; c < kCPT; ++c) {
      const float* s = src + c * HW;
      v[c][0] = __ldg(s + t.o_nw); v[c][1] = __ldg(s + t.o_ne);
      v[c][2] = __ldg(s + t.o_sw); v[c][3] = __ldg(s + t.o_se);
    }
#pragma unroll
    for (int c = 0; c < kCPT; ++c) {
      float acc = v[c][0] * t.w_nw;
      acc = fmaf(v[c][1], t.w_ne, acc);
      acc = fmaf(v[c][2], t.w_sw, acc);
      acc = fmaf(v[c][3], t.w_se, acc);
      __stcs(dst + (int64_t)c * HoWo, acc);
    }
  } else {
    for (int c = 0; c < nc; ++c) {
      const float* s = src + c * HW;
      float acc = __ldg(s + t.o_nw) * t.w_nw;
      acc = fmaf(__ldg(s + t.o_ne), t.w_ne, acc);
      acc = fmaf(__ldg(s + t.o_sw), t.w_sw, acc);
      acc = fmaf(__ldg(s + t.o_se), t.w_se, acc);
      dst[(int64_t)c * HoWo] = acc;
    }
  }
}

// refined warp (pixel flow + identity) and coarse warp (normalised prior grid, a.c.=False)
// of the same feature map: the feature tile is pulled through L1/L2 once for both outputs.
__global__ void __launch_bounds__(kThreads)
dual_warp_fwd_kernel(const float* __restrict__ in, const float* __restrict__ flow, const float* __restrict__ prior,
                     float* __restrict__ out_r, float* __restrict__ out_c, int N, int C, int H, int W) {
  const int HW = H * W;
  const int64_t gp = (int64_t)blockIdx.x * kWarpPix + (threadIdx.x % kWarpPix);
  if (gp >= (int64_t)N * HW) return;
  const int n = (int)(gp / HW);
  const int p = (int)(gp - (int64_t)n * HW);
  const int y = p / W, x = p - y * W;
  const int c0 = (blockIdx.y * (kThreads / kWarpPix) + threadIdx.x / kWarpPix) * kCPT;
  if (c0 >= C) return;

  const float fx = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 0) * HW + p), (float)x);
  const float fy = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 1) * HW + p), (float)y);
  const Taps tr = make_taps(to_pixel<MRFA_COORD_PIXEL>(fx, W), to_pixel<MRFA_COORD_PIXEL>(fy, H), H, W);
  const float2 pg = __ldg(reinterpret_cast<const float2*>(prior) + (int64_t)n * HW + p);
  const Taps tc = make_taps(to_pixel<MRFA_COORD_NORM_ACF>(pg.x, W), to_pixel<MRFA_COORD_NORM_ACF>(pg.y, H), H, W);

  const float* src = in + ((int64_t)n * C + c0) * HW;
  float* dr = out_r + ((int64_t)n * C + c0) * HW + p;
  float* dc = out_c + ((int64_t)n * C + c0) * HW + p;
  const int nc = min(kCPT, C - c0);
#pragma unroll 4
  for (int c = 0; c < nc; ++c) {
    const float* s = src + (int64_t)c * HW;
    float a0 = __ldg(s + tr.o_nw), a1 = __ldg(s + tr.o_ne), a2 = __ldg(s + tr.o_sw), a3 = __ldg(s + tr.o_se);
    float b0 = __ldg(s + tc.o_nw), b1 = __ldg(s + tc.o_ne), b2 = __ldg(s + tc.o_sw), b3 = __ldg(s + tc.o_se);
    float ar = a0 * tr.w_nw; ar = fmaf(a1, tr.w_ne, ar); ar = fmaf(a2, tr.w_sw, ar); ar = fmaf(a3, tr.w_se, ar);
    float ac = b0 * tc.w_nw; ac = fmaf(b1, tc.w_ne, ac); ac = fmaf(b2, tc.w_sw, ac); ac = fmaf(b3, tc.w_se, ac);
    __stcs(dr + (int64_t)c * HW, ar);
    __stcs(dc + (int64_t)c * HW, ac);
  }
}

// Backward.  Block = 32 pixels x 8 channel slices; slice s walks channels s, s+8, ...; the
// per-slice partial d(out)/d(coord) sums are reduced through shared memory so grad_grid is
// written once per pixel (no atomics on it); grad_in is a red.global scatter-add.
constexpr int kBwdPix = 32;
constexpr int kBwdSlices = 8;

template <int MODE, int PAD, bool ADD_ID>
__global__ void __launch_bounds__(kBwdPix * kBwdSlices)
grid_sample_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ in,
                       const float* __restrict__ grid, mrfa_grid_strides_t gs,
                       float* __restrict__ grad_in, float* __restrict__ grad_grid,
                       int N, int C, int H, int W, int Ho, int Wo, int in_batch_div) {
  __shared__ float red[2][kBwdSlices][kBwdPix];
  const int HoWo = Ho * Wo;
  const int lane = threadIdx.x % kBwdPix, slice = threadIdx.x / kBwdPix;
  const int64_t gp = (int64_t)blockIdx.x * kBwdPix + lane;
  const bool live = gp < (int64_t)N * HoWo;
  float gix = 0.f, giy = 0.f;
  float mx = 0.f, my = 0.f;
  if (live) {
    const int n = (int)(gp / HoWo);
    const int p = (int)(gp - (int64_t)n * HoWo);
    const int y = p / Wo, x = p - y * Wo;
    float ix, iy;
    load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
    const Taps t = make_taps(ix, iy, H, W);
    // distances reused by the coordinate gradient (ATen grid_sampler_2d_backward)
    const float fx = floorf(ix), fy = floorf(iy);
    const float ax = ix - fx, ay = iy - fy, bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
    const int64_t HW = (int64_t)H * W;
    const int nin = n / in_batch_div;
    for (int c = slice; c < C; c += kBwdSlices) {
      const float go = __ldg(grad_out + ((int64_t)n * C + c) * HoWo + p);
      if (grad_in != nullptr) {
        float* gi = grad_in + ((int64_t)nin * C + c) * HW;
        if (t.w_nw != 0.f) atomicAdd(gi + t.o_nw, t.w_nw * go);
        if (t.w_ne != 0.f) atomicAdd(gi + t.o_ne, t.w_ne * go);
        if (t.w_sw != 0.f) atomicAdd(gi + t.o_sw, t.w_sw * go);
        if (t.w_se != 0.f) atomicAdd(gi + t.o_se, t.w_se * go);
      }
      if (grad_grid != nullptr) {
        const float* s = in + ((int64_t)nin * C + c) * HW;
        // a tap outside the image contributes 0 (its value is the zero padding)
        const bool fin = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
        const int x0 = fin ? (int)fx : -2, y0 = fin ? (int)fy : -2;
        const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
        const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y0 + 1 >= 0) & (y0 + 1 < H);
        const float nw = (vx0 & vy0) ? __ldg(s + t.o_nw) : 0.f;
        const float ne = (vx1 & vy0) ? __ldg(s + t.o_ne) : 0.f;
        const float sw = (vx0 & vy1) ? __ldg(s + t.o_sw) : 0.f;
        const float se = (vx1 & vy1) ? __ldg(s + t.o_se) : 0.f;
        gix += go * ((ne - nw) * by + (se - sw) * ay);
        giy += go * ((sw - nw) * bx + (se - ne) * ax);
      }
    }
  }
  if (grad_grid == nullptr) return;
  red[0][slice][lane] = gix;
  red[1][slice][lane] = giy;
  __syncthreads();
  if (slice == 0 && live) {
    float sx = 0.f, sy = 0.f;
#pragma unroll
    for (int s = 0; s < kBwdSlices; ++s) { sx += red[0][s][lane]; sy += red[1][s][lane]; }
    reinterpret_cast<float2*>(grad_grid)[gp] = make_float2(sx * mx, sy * my);
  }
}

template <int MODE, int PAD>
static int launch_fwd(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N, int C, int H,
                      int W, int Ho, int Wo, int div, int add_id, cudaStream_t st) {
  dim3 g((unsigned)cdiv64((int64_t)N * Ho * Wo, kWarpPix), (unsigned)cdiv64(C, kCPT * (kThreads / kWarpPix)));
  if (add_id) grid_sample_fwd_kernel<MODE, PAD, true><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div);
  else grid_sample_fwd_kernel<MODE, PAD, false><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div);
  return MRFA_LAUNCH_RESULT();
}

template <int MODE, int PAD>
static int launch_bwd(const float* go, const float* in, const float* grid, mrfa_grid_strides_t gs, float* gi,
                      float* gg, int N, int C, int H, int W, int Ho, int Wo, int div, int add_id, cudaStream_t st) {
  dim3 g((unsigned)cdiv64((int64_t)N * Ho * Wo, kBwdPix));
  if (add_id)
    grid_sample_bwd_kernel<MODE, PAD, true><<<g, kBwdPix * kBwdSlices, 0, st>>>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div);
  else
    grid_sample_bwd_kernel<MODE, PAD, false><<<g, kBwdPix * kBwdSlices, 0, st>>>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div);
  return MRFA_LAUNCH_RESULT();
}

}  // namespace mrfa

using namespace mrfa;

#define DISPATCH_MODE_PAD(FN, ...)                                                                   \
  do {                                                                                               \
    if (padding_mode == MRFA_PAD_ZEROS) {                                                            \
      if (coord_mode == MRFA_COORD_NORM_ACF) return FN<MRFA_COORD_NORM_ACF, MRFA_PAD_ZEROS>(__VA_ARGS__); \
      if (coord_mode == MRFA_COORD_NORM_ACT) return FN<MRFA_COORD_NORM_ACT, MRFA_PAD_ZEROS>(__VA_ARGS__); \
      if (coord_mode == MRFA_COORD_PIXEL) return FN<MRFA_COORD_PIXEL, MRFA_PAD_ZEROS>(__VA_ARGS__);  \
    } else if (padding_mode == MRFA_PAD_REFLECTION) {                                                \
      if (coord_mode == MRFA_COORD_NORM_ACF) return FN<MRFA_COORD_NORM_ACF, MRFA_PAD_REFLECTION>(__VA_ARGS__); \
      if (coord_mode == MRFA_COORD_NORM_ACT) return FN<MRFA_COORD_NORM_ACT, MRFA_PAD_REFLECTION>(__VA_ARGS__); \
    }                                                                                                \
    return MRFA_E_BADARG;                                                                            \
  } while (0)

extern "C" int mrfa_grid_sample_fwd(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N,
                                    int C, int H, int W, int Ho, int Wo, int in_batch_div, int coord_mode,
                                    int padding_mode, int add_identity, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && grid && out);
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && in_batch_div >= 1);
  MRFA_CHECK_ARG(N % in_batch_div == 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31) && (int64_t)Ho * Wo < (1ll << 31));
  if (N == 0) return 0;
  DISPATCH_MODE_PAD(launch_fwd, in, grid, gs, out, N, C, H, W, Ho, Wo, in_batch_div, add_identity, as_stream(stream));
}

extern "C" int mrfa_grid_sample_bwd(const float* grad_out, const float* in, const float* grid,
                                    mrfa_grid_strides_t gs, float* grad_in, float* grad_grid, int N, int C, int H,
                                    int W, int Ho, int Wo, int in_batch_div, int coord_mode, int padding_mode,
                                    int add_identity, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_out && in && grid && (grad_in || grad_grid));
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && in_batch_div >= 1);
  MRFA_CHECK_ARG(N % in_batch_div == 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31) && (int64_t)Ho * Wo < (1ll << 31));
  if (N == 0) return 0;
  DISPATCH_MODE_PAD(launch_bwd, grad_out, in, grid, gs, grad_in, grad_grid, N, C, H, W, Ho, Wo, in_batch_div,
                    add_identity, as_stream(stream));
}

extern "C" int mrfa_dual_warp_fwd(const float* in, const float* flow, const float* prior_grid, float* out_refined,
                                  float* out_coarse, int N, int C, int H, int W, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && flow && prior_grid && out_refined && out_coarse);
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31));
  if (N == 0) return 0;
  dim3 g((unsigned)cdiv64((int64_t)N * H * W, kWarpPix), (unsigned)cdiv64(C, kCPT * (kThreads / kWarpPix)));
  dual_warp_fwd_kernel<<<g, kThreads, 0, as_stream(stream)>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W);
  return MRFA_LAUNCH_RESULT();
}
