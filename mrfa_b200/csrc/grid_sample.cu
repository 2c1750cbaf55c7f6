// Bilinear warps (K5/K6 of SURVEY.md): forward, backward and the fused dual warp.
//
// Memory-bound gathers over NCHW fp32 planes.  One thread owns one output pixel and a chunk
// of CPT channels: the four tap offsets/weights are computed once and reused for every
// channel of the chunk, the 4*CPT tap loads are issued back to back (memory-level
// parallelism), and lanes map to consecutive output pixels so both the taps (smooth motion
// fields) and the stores are coalesced.  Roofline: HBM, algorithmic bytes per warp call =
// (2*C*Ho*Wo + 2*Ho*Wo) * 4.
#include <stdlib.h>
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
      *(dst + (int64_t)c * HoWo) = acc;
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

// Few-channel NCHW maps (C <= 4: the full-resolution image warp raft.py:302, C = 3).  With so few channels the generic
// kernel above is bound by the latency of its dependent load chain (grid value -> taps -> store: ncu showed 14 % of the
// DRAM peak at 47 % occupancy), not by bytes: each thread here owns kFewPix pixels 128 apart (lanes stay on consecutive
// pixels, so grid loads, taps and stores coalesce as before) and issues all grid loads, then all 4 * C * kFewPix tap
// loads, before the first use -- four times the bytes in flight per thread.
constexpr int kFewPix = 4;
constexpr int kFewThreads = 128;

template <int MODE, int PAD, bool ADD_ID, int CC>
__global__ void __launch_bounds__(kFewThreads)
grid_sample_fwd_fewc_kernel(const float* __restrict__ in, const float* __restrict__ grid, mrfa_grid_strides_t gs,
                            float* __restrict__ out, int H, int W, int Ho, int Wo, FastDiv wo_div, int in_batch_div) {
  const int HoWo = Ho * Wo;
  const int n = blockIdx.y;                            // one sample per grid row: no division by the plane size
  const int p0 = blockIdx.x * (kFewThreads * kFewPix) + threadIdx.x;
  const int HW = H * W;
  float gx[kFewPix], gy[kFewPix];
  int nn[kFewPix], pp[kFewPix], xx[kFewPix], yy[kFewPix];
#pragma unroll
  for (int j = 0; j < kFewPix; ++j) {
    const int pj = p0 + j * kFewThreads;
    const bool live = pj < HoWo;
    const int p = live ? pj : 0;
    nn[j] = live ? n : -1; pp[j] = p;
    yy[j] = (int)fast_div((uint32_t)p, wo_div); xx[j] = p - yy[j] * Wo;
    const float* g = grid + n * gs.sn + yy[j] * gs.sy + xx[j] * gs.sx;
    gx[j] = __ldg(g); gy[j] = __ldg(g + gs.sc);
  }
  Taps t[kFewPix];
#pragma unroll
  for (int j = 0; j < kFewPix; ++j) {
    float vx = gx[j], vy = gy[j], mx, my;
    if (ADD_ID) { vx = __fadd_rn(vx, (float)xx[j]); vy = __fadd_rn(vy, (float)yy[j]); }
    t[j] = make_taps(source_index<MODE, PAD>(vx, W, &mx), source_index<MODE, PAD>(vy, H, &my), H, W);
  }
  float v[kFewPix][CC][4];
#pragma unroll
  for (int j = 0; j < kFewPix; ++j) {
    const float* src = in + (int64_t)(n / in_batch_div) * CC * HW;
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      const float* s = src + c * HW;
      v[j][c][0] = __ldg(s + t[j].o_nw); v[j][c][1] = __ldg(s + t[j].o_ne);
      v[j][c][2] = __ldg(s + t[j].o_sw); v[j][c][3] = __ldg(s + t[j].o_se);
    }
  }
#pragma unroll
  for (int j = 0; j < kFewPix; ++j) {
    if (nn[j] < 0) continue;
    float* dst = out + (int64_t)nn[j] * CC * HoWo + pp[j];
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      float acc = v[j][c][0] * t[j].w_nw;                 // same evaluation order as the generic kernel (and ATen)
      acc = fmaf(v[j][c][1], t[j].w_ne, acc);
      acc = fmaf(v[j][c][2], t[j].w_sw, acc);
      acc = fmaf(v[j][c][3], t[j].w_se, acc);
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
    *(dr + (int64_t)c * HW) = ar;
    *(dc + (int64_t)c * HW) = ac;
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


// ---------------------------------------------------------------------------------------------
// channels-last (NHWC) variants: the memory layout cuDNN's tensor-core convolutions produce on
// sm_100.  A tap is C contiguous floats, so 16 lanes x float4 read one tap of one pixel as 256
// contiguous bytes (C = 64) and the store is a full line -- no scalar gathers at all.
// Thread group = 16 lanes per output pixel (2 pixels per warp); lane j walks channel quads
// j, j+16, ...  (C % 4 == 0).
// ---------------------------------------------------------------------------------------------
constexpr int kNhwcGroup = 16;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stcs4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }  // default policy: measured 4-5 % faster than .cs for streams
__device__ __forceinline__ float4 blend4(float4 a, float4 b, float4 c, float4 d, const Taps& t) {
  float4 r;
  r.x = fmaf(d.x, t.w_se, fmaf(c.x, t.w_sw, fmaf(b.x, t.w_ne, a.x * t.w_nw)));
  r.y = fmaf(d.y, t.w_se, fmaf(c.y, t.w_sw, fmaf(b.y, t.w_ne, a.y * t.w_nw)));
  r.z = fmaf(d.z, t.w_se, fmaf(c.z, t.w_sw, fmaf(b.z, t.w_ne, a.z * t.w_nw)));
  r.w = fmaf(d.w, t.w_se, fmaf(c.w, t.w_sw, fmaf(b.w, t.w_ne, a.w * t.w_nw)));
  return r;
}

// Forward NHWC kernels are warp-cooperative: lane l evaluates the coordinates / taps of pixel
// (warp base + l) ONCE, then the warp walks its 32 pixels, broadcasting each pixel's taps with
// shuffles while all lanes stream that pixel's channels as float4 (LPP lanes per pixel, 32/LPP
// pixels per step).  The IEEE coordinate arithmetic is thereby amortised over the channel axis
// instead of being repeated by every lane of a pixel.
struct TapsB {              // taps + input-plane base of one pixel, as shuffled between lanes
  int o_nw, o_ne, o_sw, o_se, n_in;
  float w_nw, w_ne, w_sw, w_se;
};
__device__ __forceinline__ TapsB shfl_taps(const TapsB& t, int src) {
  TapsB r;
  r.o_nw = __shfl_sync(0xffffffffu, t.o_nw, src); r.o_ne = __shfl_sync(0xffffffffu, t.o_ne, src);
  r.o_sw = __shfl_sync(0xffffffffu, t.o_sw, src); r.o_se = __shfl_sync(0xffffffffu, t.o_se, src);
  r.n_in = __shfl_sync(0xffffffffu, t.n_in, src);
  r.w_nw = __shfl_sync(0xffffffffu, t.w_nw, src); r.w_ne = __shfl_sync(0xffffffffu, t.w_ne, src);
  r.w_sw = __shfl_sync(0xffffffffu, t.w_sw, src); r.w_se = __shfl_sync(0xffffffffu, t.w_se, src);
  return r;
}
__device__ __forceinline__ float4 blend4b(float4 a, float4 b, float4 c, float4 d, const TapsB& t) {
  float4 r;
  r.x = fmaf(d.x, t.w_se, fmaf(c.x, t.w_sw, fmaf(b.x, t.w_ne, a.x * t.w_nw)));
  r.y = fmaf(d.y, t.w_se, fmaf(c.y, t.w_sw, fmaf(b.y, t.w_ne, a.y * t.w_nw)));
  r.z = fmaf(d.z, t.w_se, fmaf(c.z, t.w_sw, fmaf(b.z, t.w_ne, a.z * t.w_nw)));
  r.w = fmaf(d.w, t.w_se, fmaf(c.w, t.w_sw, fmaf(b.w, t.w_ne, a.w * t.w_nw)));
  return r;
}
// offsets are pre-multiplied by C (element offsets inside one image, < 2^31 by the ABI check)
__device__ __forceinline__ TapsB to_tapsb(const Taps& t, int n_in, int C) {
  TapsB r;
  r.o_nw = t.o_nw * C; r.o_ne = t.o_ne * C; r.o_sw = t.o_sw * C; r.o_se = t.o_se * C; r.n_in = n_in;
  r.w_nw = t.w_nw; r.w_ne = t.w_ne; r.w_sw = t.w_sw; r.w_se = t.w_se;
  return r;
}

// LPP = lanes per pixel (power of two, LPP * 4 <= C or LPP == 1)
// PW = pixels per warp (32 for large problems; 4 keeps enough warps in flight for small ones)
template <int MODE, int PAD, bool ADD_ID, int LPP, int PW, int UNROLL>
__global__ void __launch_bounds__(kThreads)
grid_sample_fwd_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ grid, mrfa_grid_strides_t gs,
                            float* __restrict__ out, int N, int C, int H, int W, int Ho, int Wo, int in_batch_div) {
  constexpr int PPS = (32 / LPP) < PW ? (32 / LPP) : PW;   // pixels per step
  constexpr int LPPE = 32 / PPS;                            // lanes actually assigned to one pixel
  const int HoWo = Ho * Wo;
  const int lane = threadIdx.x % 32;
  const int64_t total = (int64_t)N * HoWo;
  const int64_t gp0 = ((int64_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32) * PW;
  if (gp0 >= total) return;
  TapsB mine;
  {
    const int64_t gp = min(gp0 + (lane % PW), total - 1);
    const int n = (int)(gp / HoWo);
    const int p = (int)(gp - (int64_t)n * HoWo);
    const int y = p / Wo, x = p - y * Wo;
    float ix, iy, mx, my;
    load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
    mine = to_tapsb(make_taps(ix, iy, H, W), n / in_batch_div, C);
  }
  const int sub = lane / LPPE, cl = (lane % LPPE) * 4;
  const int64_t plane = (int64_t)H * W * C;
#pragma unroll UNROLL
  for (int s = 0; s < PW; s += PPS) {
    const TapsB t = shfl_taps(mine, s + sub);
    const int64_t gp = gp0 + s + sub;
    if (gp >= total) continue;
    const float* src = in + (int64_t)t.n_in * plane + cl;
    float* dst = out + gp * C + cl;
    for (int c = 0; c < C - cl; c += LPPE * 4) {
      const float4 a = ldg4(src + (t.o_nw + c)), b = ldg4(src + (t.o_ne + c));
      const float4 d = ldg4(src + (t.o_sw + c)), e = ldg4(src + (t.o_se + c));
      stcs4(dst + c, blend4b(a, b, d, e, t));
    }
  }
}

// Two passes over the warp's pixels -- refined (pixel flow + identity), then coarse (normalised prior grid) --
// so only one tap set is live at a time (64 registers, 4 blocks per SM, no spills); the second pass finds
// the feature rows of its neighbourhood in L1/L2, so `in` still leaves HBM once.
template <int LPP, int PW>
__global__ void __launch_bounds__(kThreads, 4)
dual_warp_fwd_nhwc_kernel(const float* __restrict__ in, const float* __restrict__ flow, const float* __restrict__ prior,
                          float* __restrict__ out_r, float* __restrict__ out_c, int N, int C, int H, int W,
                          int64_t cstride) {
  constexpr int PPS = (32 / LPP) < PW ? (32 / LPP) : PW;
  constexpr int LPPE = 32 / PPS;
  const int HW = H * W;
  const int lane = threadIdx.x % 32;
  const int64_t total = (int64_t)N * HW;
  const int64_t gp0 = ((int64_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32) * PW;
  if (gp0 >= total) return;
  const int sub = lane / LPPE, cl = (lane % LPPE) * 4;
  const int64_t plane = (int64_t)HW * C;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    TapsB m;
    {
      const int64_t gp = min(gp0 + (lane % PW), total - 1);
      const int n = (int)(gp / HW);
      const int p = (int)(gp - (int64_t)n * HW);
      if (pass == 0) {
        const int y = p / W, x = p - y * W;
        const float fx = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 0) * HW + p), (float)x);
        const float fy = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 1) * HW + p), (float)y);
        m = to_tapsb(make_taps(to_pixel<MRFA_COORD_PIXEL>(fx, W), to_pixel<MRFA_COORD_PIXEL>(fy, H), H, W), n, C);
      } else {
        const float2 pg = __ldg(reinterpret_cast<const float2*>(prior) + gp);
        m = to_tapsb(make_taps(to_pixel<MRFA_COORD_NORM_ACF>(pg.x, W), to_pixel<MRFA_COORD_NORM_ACF>(pg.y, H), H, W), n, C);
      }
    }
    float* const out = pass == 0 ? out_r : out_c;     // the coarse warp may land in a channel slice of a wider NHWC buffer
    const int64_t ostride = pass == 0 ? (int64_t)C : cstride;
#pragma unroll 2
    for (int s = 0; s < PW; s += PPS) {
      const TapsB t = shfl_taps(m, s + sub);
      const int64_t gp = gp0 + s + sub;
      if (gp >= total) continue;
      const float* src = in + (int64_t)t.n_in * plane + cl;
      float* d = out + gp * ostride + cl;
      for (int c = 0; c < C - cl; c += LPPE * 4) {
        const float4 a0 = ldg4(src + (t.o_nw + c)), a1 = ldg4(src + (t.o_ne + c));
        const float4 a2 = ldg4(src + (t.o_sw + c)), a3 = ldg4(src + (t.o_se + c));
        stcs4(d + c, blend4b(a0, a1, a2, a3, t));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Run-walk forward kernels (the production path for >= 100k output pixels).
//
// The plain NHWC kernels above pull 4 taps x C floats through L1 for every output pixel: 5 units of L1 / LSU traffic
// (4 loads + 1 store) per unit of output, which at C = 64, R = 256 is more than the 128 B/clk/SM L1 data path delivers
// in the time HBM needs for the algorithmic bytes -- the kernels were L1-bound at ~0.55 of the HBM roofline.
// Motion fields are smooth: along a row the sample position advances by ~1 source pixel per output pixel, so the right
// taps (ne, se) of pixel x are the left taps (nw, sw) of pixel x + 1.  Each lane group (the lanes that cover the C channels
// of one pixel) therefore walks a RUN of consecutive pixels and keeps the previous right column in registers: when the
// tap offsets chain (checked per pixel, exact -- any flow is still handled, it just reloads) a pixel costs 2 loads + 1
// store.  V = floats per lane access: 8 -> 256-bit LDG.E.ENL2.256 / STG.E.ENL2.256 (sm_100), 4 -> 128-bit.
// ---------------------------------------------------------------------------------------------
template <int V> struct VecF { float v[V]; };

template <int V> __device__ __forceinline__ VecF<V> ldgv(const float* p);
template <> __device__ __forceinline__ VecF<4> ldgv<4>(const float* p) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  VecF<4> r; r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  return r;
}
template <> __device__ __forceinline__ VecF<8> ldgv<8>(const float* p) {
  VecF<8> r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}
template <int V> __device__ __forceinline__ void stv(float* p, const VecF<V>& r);
template <> __device__ __forceinline__ void stv<4>(float* p, const VecF<4>& r) {
  *reinterpret_cast<float4*>(p) = make_float4(r.v[0], r.v[1], r.v[2], r.v[3]);
}
template <> __device__ __forceinline__ void stv<8>(float* p, const VecF<8>& r) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]),
               "f"(r.v[3]), "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}
template <int V>
__device__ __forceinline__ VecF<V> blendv(const VecF<V>& a, const VecF<V>& b, const VecF<V>& c, const VecF<V>& d, const TapsB& t) {
  VecF<V> r;     // same evaluation order as ATen / blend4b: nw, ne, sw, se
#pragma unroll
  for (int i = 0; i < V; ++i) r.v[i] = fmaf(d.v[i], t.w_se, fmaf(c.v[i], t.w_sw, fmaf(b.v[i], t.w_ne, a.v[i] * t.w_nw)));
  return r;
}

// One pass of a warp over its `pw` consecutive pixels (pw a power of two, G <= pw <= 32): lane l holds the taps of pixel
// gp0 + (l & (pw - 1)) in `mine`; group `grp` (LPPE lanes, G = 32 / LPPE groups) walks pixels [grp * run, (grp + 1) * run),
// run = pw / G, channel chunk by channel chunk.  Fewer pixels per warp = more warps for the mid-sized levels.
template <int V, int LPPE>
__device__ __forceinline__ void run_walk(const TapsB& mine, const float* __restrict__ in, float* __restrict__ out,
                                         int64_t gp0, int64_t total, int C, int64_t plane, int64_t ostride, int lane, int pw) {
  constexpr int G = 32 / LPPE;
  const int run = pw / G;
  const int grp = lane / LPPE, cl = (lane % LPPE) * V;
  for (int c0 = 0; c0 < C; c0 += LPPE * V) {                     // warp-uniform trip count: the shuffles below need every lane
    const int c = c0 + cl;
    const bool act = c < C;
    VecF<V> pr_ne, pr_se;
    int po_ne = -1, po_se = -1, pn = -1;
#pragma unroll 4
    for (int i = 0; i < run; ++i) {
      const TapsB t = shfl_taps(mine, grp * run + i);
      const int64_t gp = gp0 + grp * run + i;
      if (gp >= total || !act) continue;
      const float* src = in + (int64_t)t.n_in * plane + c;
      const bool chain = (t.o_nw == po_ne) & (t.o_sw == po_se) & (t.n_in == pn);
      VecF<V> a, d;
      if (chain) { a = pr_ne; d = pr_se; }
      else { a = ldgv<V>(src + t.o_nw); d = ldgv<V>(src + t.o_sw); }
      const VecF<V> b = ldgv<V>(src + t.o_ne), e = ldgv<V>(src + t.o_se);
      stv<V>(out + gp * ostride + c, blendv<V>(a, b, d, e, t));
      pr_ne = b; pr_se = e; po_ne = t.o_ne; po_se = t.o_se; pn = t.n_in;
    }
  }
}

template <int MODE, int PAD, bool ADD_ID, int V, int LPPE>
__global__ void __launch_bounds__(kThreads)
grid_sample_fwd_nhwc_run_kernel(const float* __restrict__ in, const float* __restrict__ grid, mrfa_grid_strides_t gs,
                                float* __restrict__ out, int N, int C, int H, int W, int Ho, int Wo, int in_batch_div, int pw) {
  const int HoWo = Ho * Wo;
  const int lane = threadIdx.x % 32;
  const int64_t total = (int64_t)N * HoWo;
  const int64_t gp0 = ((int64_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32) * pw;
  if (gp0 >= total) return;
  TapsB mine;
  {
    const int64_t gp = min(gp0 + (lane & (pw - 1)), total - 1);
    const int n = (int)(gp / HoWo);
    const int p = (int)(gp - (int64_t)n * HoWo);
    const int y = p / Wo, x = p - y * Wo;
    float ix, iy, mx, my;
    load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
    mine = to_tapsb(make_taps(ix, iy, H, W), n / in_batch_div, C);
  }
  run_walk<V, LPPE>(mine, in, out, gp0, total, C, (int64_t)H * W * C, C, lane, pw);
}

// OCC4: cap the kernel at 64 registers (4 blocks = 32 resident warps per SM) -- ncu shows the walk is long-scoreboard bound
// at 3 blocks, so more warps means more tap loads in flight; costs a 16-byte spill
template <int V, int LPPE, bool OCC4>
__global__ void __launch_bounds__(kThreads, OCC4 ? 4 : 1)
dual_warp_fwd_nhwc_run_kernel(const float* __restrict__ in, const float* __restrict__ flow, const float* __restrict__ prior,
                              float* __restrict__ out_r, float* __restrict__ out_c, int N, int C, int H, int W,
                              int64_t cstride, int pw) {
  const int HW = H * W;
  const int lane = threadIdx.x % 32;
  const int64_t total = (int64_t)N * HW;
  const int64_t gp0 = ((int64_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32) * pw;
  if (gp0 >= total) return;
  const int64_t gp = min(gp0 + (lane & (pw - 1)), total - 1);
  const int n = (int)(gp / HW);
  const int p = (int)(gp - (int64_t)n * HW);
  const int64_t plane = (int64_t)HW * C;
  {   // refined warp: pixel flow + identity (raft.py:247)
    const int y = p / W, x = p - y * W;
    const float fx = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 0) * HW + p), (float)x);
    const float fy = __fadd_rn(__ldg(flow + ((int64_t)n * 2 + 1) * HW + p), (float)y);
    const TapsB m = to_tapsb(make_taps(to_pixel<MRFA_COORD_PIXEL>(fx, W), to_pixel<MRFA_COORD_PIXEL>(fy, H), H, W), n, C);
    run_walk<V, LPPE>(m, in, out_r, gp0, total, C, plane, C, lane, pw);
  }
  {   // coarse warp: normalised prior grid, align_corners=False (raft.py:271); may land in a channel slice of a wider buffer
    const float2 pg = __ldg(reinterpret_cast<const float2*>(prior) + gp);
    const TapsB m = to_tapsb(make_taps(to_pixel<MRFA_COORD_NORM_ACF>(pg.x, W), to_pixel<MRFA_COORD_NORM_ACF>(pg.y, H), H, W), n, C);
    run_walk<V, LPPE>(m, in, out_c, gp0, total, C, plane, cstride, lane, pw);
  }
}

// NHWC backward: one thread group (16 lanes) per output pixel; the grad_input scatter is a
// vector red.global.add.v4.f32 per tap and channel quad; the coordinate gradient is reduced over
// the channel axis with shuffles inside the 16-lane group.
template <int MODE, int PAD, bool ADD_ID>
__global__ void __launch_bounds__(kThreads)
grid_sample_bwd_nhwc_kernel(const float* __restrict__ grad_out, const float* __restrict__ in,
                            const float* __restrict__ grid, mrfa_grid_strides_t gs, float* __restrict__ grad_in,
                            float* __restrict__ grad_grid, int N, int C, int H, int W, int Ho, int Wo,
                            int in_batch_div) {
  const int HoWo = Ho * Wo;
  const int64_t gp = ((int64_t)blockIdx.x * kThreads + threadIdx.x) / kNhwcGroup;
  const int j = threadIdx.x % kNhwcGroup;
  const bool live = gp < (int64_t)N * HoWo;
  float gix = 0.f, giy = 0.f, mx = 0.f, my = 0.f;
  if (live) {
    const int n = (int)(gp / HoWo);
    const int p = (int)(gp - (int64_t)n * HoWo);
    const int y = p / Wo, x = p - y * Wo;
    float ix, iy;
    load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
    const Taps t = make_taps(ix, iy, H, W);
    const float fx = floorf(ix), fy = floorf(iy);
    const float ax = ix - fx, ay = iy - fy, bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;
    const bool fin = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
    const int x0 = fin ? (int)fx : -2, y0 = fin ? (int)fy : -2;
    const bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
    const bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y0 + 1 >= 0) & (y0 + 1 < H);
    const int64_t ibase = (int64_t)(n / in_batch_div) * H * W * C;
    const float* go = grad_out + gp * C;
    for (int c = j * 4; c < C; c += kNhwcGroup * 4) {
      const float4 g = ldg4(go + c);
      if (grad_in != nullptr) {
        float* gi = grad_in + ibase + c;
        if (t.w_nw != 0.f) atomicAdd(reinterpret_cast<float4*>(gi + (int64_t)t.o_nw * C), make_float4(g.x * t.w_nw, g.y * t.w_nw, g.z * t.w_nw, g.w * t.w_nw));
        if (t.w_ne != 0.f) atomicAdd(reinterpret_cast<float4*>(gi + (int64_t)t.o_ne * C), make_float4(g.x * t.w_ne, g.y * t.w_ne, g.z * t.w_ne, g.w * t.w_ne));
        if (t.w_sw != 0.f) atomicAdd(reinterpret_cast<float4*>(gi + (int64_t)t.o_sw * C), make_float4(g.x * t.w_sw, g.y * t.w_sw, g.z * t.w_sw, g.w * t.w_sw));
        if (t.w_se != 0.f) atomicAdd(reinterpret_cast<float4*>(gi + (int64_t)t.o_se * C), make_float4(g.x * t.w_se, g.y * t.w_se, g.z * t.w_se, g.w * t.w_se));
      }
      if (grad_grid != nullptr) {
        const float* s = in + ibase + c;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 nw = (vx0 & vy0) ? ldg4(s + (int64_t)t.o_nw * C) : z;
        const float4 ne = (vx1 & vy0) ? ldg4(s + (int64_t)t.o_ne * C) : z;
        const float4 sw = (vx0 & vy1) ? ldg4(s + (int64_t)t.o_sw * C) : z;
        const float4 se = (vx1 & vy1) ? ldg4(s + (int64_t)t.o_se * C) : z;
        gix += g.x * ((ne.x - nw.x) * by + (se.x - sw.x) * ay) + g.y * ((ne.y - nw.y) * by + (se.y - sw.y) * ay) +
               g.z * ((ne.z - nw.z) * by + (se.z - sw.z) * ay) + g.w * ((ne.w - nw.w) * by + (se.w - sw.w) * ay);
        giy += g.x * ((sw.x - nw.x) * bx + (se.x - ne.x) * ax) + g.y * ((sw.y - nw.y) * bx + (se.y - ne.y) * ax) +
               g.z * ((sw.z - nw.z) * bx + (se.z - ne.z) * ax) + g.w * ((sw.w - nw.w) * bx + (se.w - ne.w) * ax);
      }
    }
  }
  if (grad_grid == nullptr) return;
#pragma unroll
  for (int o = kNhwcGroup / 2; o > 0; o >>= 1) {
    gix += __shfl_xor_sync(0xffffffffu, gix, o);
    giy += __shfl_xor_sync(0xffffffffu, giy, o);
  }
  if (live && j == 0) reinterpret_cast<float2*>(grad_grid)[gp] = make_float2(gix * mx, giy * my);
}

// ---------------------------------------------------------------------------------------------
// Run-walk backward (NHWC): the scatter-add pre-reduction.  ATen issues one atomic per tap (4 per output pixel and
// channel); along a row of a smooth motion field the right-hand cells (ne, se) of pixel x are the left-hand cells (nw, sw)
// of pixel x + 1, so a lane group that walks a run of consecutive pixels carries the right column's contributions in
// registers and merges them with the next pixel's left column before issuing ONE red.global.add.v4.f32 per cell:
// 2 vector atomics per pixel and channel quad instead of 4 when the cells chain (checked per pixel on the exact cell
// addresses; anything else flushes and falls back to 4).  The input taps needed for the coordinate gradient chain the same
// way (2 loads instead of 4); that gradient is reduced over the channel axis with warp shuffles and written once per pixel.
// ---------------------------------------------------------------------------------------------
struct TapsG {              // TapsB + what the coordinate gradient needs
  TapsB t;
  float ax, ay, bx, by;     // distances to the west / north and east / south cell centres
  int valid;                // bit 0 nw, 1 ne, 2 sw, 3 se: tap inside the image (zeros padding contributes nothing)
};
__device__ __forceinline__ TapsG shfl_tapsg(const TapsG& g, int src) {
  TapsG r;
  r.t = shfl_taps(g.t, src);
  r.ax = __shfl_sync(0xffffffffu, g.ax, src); r.ay = __shfl_sync(0xffffffffu, g.ay, src);
  r.bx = __shfl_sync(0xffffffffu, g.bx, src); r.by = __shfl_sync(0xffffffffu, g.by, src);
  r.valid = __shfl_sync(0xffffffffu, g.valid, src);
  return r;
}
__device__ __forceinline__ float4 scale4(float4 g, float w) { return make_float4(g.x * w, g.y * w, g.z * w, g.w * w); }
__device__ __forceinline__ float4 fma4(float4 g, float w, float4 a) {
  return make_float4(fmaf(g.x, w, a.x), fmaf(g.y, w, a.y), fmaf(g.z, w, a.z), fmaf(g.w, w, a.w));
}
__device__ __forceinline__ bool nonzero4(float4 a) { return a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f; }

template <int MODE, int PAD, bool ADD_ID, int LPPE>
__global__ void __launch_bounds__(kThreads)
grid_sample_bwd_nhwc_run_kernel(const float* __restrict__ grad_out, const float* __restrict__ in,
                                const float* __restrict__ grid, mrfa_grid_strides_t gs, float* __restrict__ grad_in,
                                float* __restrict__ grad_grid, int N, int C, int H, int W, int Ho, int Wo, int in_batch_div,
                                int pw) {
  constexpr int G = 32 / LPPE;
  const int HoWo = Ho * Wo;
  const int lane = threadIdx.x % 32;
  const int64_t total = (int64_t)N * HoWo;
  const int64_t gp0 = ((int64_t)blockIdx.x * (kThreads / 32) + threadIdx.x / 32) * pw;
  if (gp0 >= total) return;
  const int own = lane & (pw - 1);                       // the pixel of the warp this lane evaluated the geometry of
  TapsG mine;
  float mx, my;
  {
    const int64_t gp = min(gp0 + own, total - 1);
    const int n = (int)(gp / HoWo);
    const int p = (int)(gp - (int64_t)n * HoWo);
    const int y = p / Wo, x = p - y * Wo;
    float ix, iy;
    load_sample_point<MODE, PAD, ADD_ID>(grid, gs, n, y, x, H, W, ix, iy, mx, my);
    mine.t = to_tapsb(make_taps(ix, iy, H, W), n / in_batch_div, C);
    const float fx = floorf(ix), fy = floorf(iy);
    mine.ax = ix - fx; mine.ay = iy - fy; mine.bx = (fx + 1.f) - ix; mine.by = (fy + 1.f) - iy;
    const bool fin = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);
    const int x0 = fin ? (int)fx : -2, y0 = fin ? (int)fy : -2;
    const int vx0 = (x0 >= 0) & (x0 < W), vx1 = (x0 + 1 >= 0) & (x0 + 1 < W);
    const int vy0 = (y0 >= 0) & (y0 < H), vy1 = (y0 + 1 >= 0) & (y0 + 1 < H);
    mine.valid = (vx0 & vy0) | ((vx1 & vy0) << 1) | ((vx0 & vy1) << 2) | ((vx1 & vy1) << 3);
  }
  const int run = pw / G;
  const int grp = lane / LPPE, cl = (lane % LPPE) * 4;
  const int64_t plane = (int64_t)H * W * C;
  float gix_own = 0.f, giy_own = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = 0; c0 < C; c0 += LPPE * 4) {              // warp-uniform trip count (shuffles inside)
    const int c = c0 + cl;
    const bool act = c < C;
    float4 cn = z4, cs = z4;                              // carried contributions to the cells (o_ne, o_se) of the previous pixel
    int ko_n = -1, ko_s = -1, kn = -1;
    float4 pv_ne = z4, pv_se = z4;                        // carried input taps (coordinate gradient)
    int lo_n = -1, lo_s = -1, ln = -1;
#pragma unroll 2
    for (int i = 0; i < run; ++i) {
      const TapsG tg = shfl_tapsg(mine, grp * run + i);
      const TapsB& t = tg.t;
      const int64_t gp = gp0 + grp * run + i;
      const bool live = act && gp < total;
      float gix = 0.f, giy = 0.f;
      if (live) {
        const float4 g = ldg4(grad_out + gp * C + c);
        if (grad_in != nullptr) {
          float* gi = grad_in + (int64_t)t.n_in * plane + c;
          float4 an = scale4(g, t.w_nw), as = scale4(g, t.w_sw);
          if ((t.o_nw == ko_n) & (t.o_sw == ko_s) & (t.n_in == kn)) {
            an.x += cn.x; an.y += cn.y; an.z += cn.z; an.w += cn.w;
            as.x += cs.x; as.y += cs.y; as.z += cs.z; as.w += cs.w;
          } else if (kn >= 0) {                            // the chain broke: flush the carried column
            float* gk = grad_in + (int64_t)kn * plane + c;
            if (nonzero4(cn)) atomicAdd(reinterpret_cast<float4*>(gk + ko_n), cn);
            if (nonzero4(cs)) atomicAdd(reinterpret_cast<float4*>(gk + ko_s), cs);
          }
          if (nonzero4(an)) atomicAdd(reinterpret_cast<float4*>(gi + t.o_nw), an);
          if (nonzero4(as)) atomicAdd(reinterpret_cast<float4*>(gi + t.o_sw), as);
          cn = scale4(g, t.w_ne); cs = scale4(g, t.w_se);
          ko_n = t.o_ne; ko_s = t.o_se; kn = t.n_in;
        }
        if (grad_grid != nullptr) {
          const float* s = in + (int64_t)t.n_in * plane + c;
          float4 nw, sw;
          if ((t.o_nw == lo_n) & (t.o_sw == lo_s) & (t.n_in == ln)) { nw = pv_ne; sw = pv_se; }
          else { nw = ldg4(s + t.o_nw); sw = ldg4(s + t.o_sw); }
          const float4 ne = ldg4(s + t.o_ne), se = ldg4(s + t.o_se);
          pv_ne = ne; pv_se = se; lo_n = t.o_ne; lo_s = t.o_se; ln = t.n_in;
          // a tap outside the image contributes 0 (its value is the zero padding, the clamped address holds something else)
          const float m0 = (tg.valid & 1) ? 1.f : 0.f, m1 = (tg.valid & 2) ? 1.f : 0.f;
          const float m2 = (tg.valid & 4) ? 1.f : 0.f, m3 = (tg.valid & 8) ? 1.f : 0.f;
          const float* pnw = &nw.x; const float* pne = &ne.x; const float* psw = &sw.x; const float* pse = &se.x;
          const float* pg = &g.x;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float vnw = pnw[e] * m0, vne = pne[e] * m1, vsw = psw[e] * m2, vse = pse[e] * m3;
            gix += pg[e] * ((vne - vnw) * tg.by + (vse - vsw) * tg.ay);
            giy += pg[e] * ((vsw - vnw) * tg.bx + (vse - vne) * tg.ax);
          }
        }
      }
      if (grad_grid != nullptr) {
        // reduce over the channel lanes of the group, then hand the sums to the lane that owns this pixel's geometry
#pragma unroll
        for (int o = LPPE / 2; o > 0; o >>= 1) {
          gix += __shfl_xor_sync(0xffffffffu, gix, o);
          giy += __shfl_xor_sync(0xffffffffu, giy, o);
        }
        const int src = (own / run) * LPPE;                // leader lane of the group that walks this lane's pixel
        const float rx = __shfl_sync(0xffffffffu, gix, src), ry = __shfl_sync(0xffffffffu, giy, src);
        if (own % run == i) { gix_own += rx; giy_own += ry; }
      }
    }
    if (grad_in != nullptr && kn >= 0) {                    // end of the run: flush the last right-hand column
      float* gk = grad_in + (int64_t)kn * plane + c;
      if (nonzero4(cn)) atomicAdd(reinterpret_cast<float4*>(gk + ko_n), cn);
      if (nonzero4(cs)) atomicAdd(reinterpret_cast<float4*>(gk + ko_s), cs);
    }
  }
  if (grad_grid != nullptr && lane < pw && gp0 + lane < total)
    reinterpret_cast<float2*>(grad_grid)[gp0 + lane] = make_float2(gix_own * mx, giy_own * my);
}

static inline int lanes_per_pixel(int C) {
  int lpp = 1;
  while (lpp < 32 && lpp * 2 * 4 <= C) lpp *= 2;
  return lpp;
}

constexpr int64_t kSmallPixels = 100000;     // below this, 4 pixels per warp keep the SMs busy

// run-walk configuration: lanes per pixel for V floats per lane (power of two, <= 32)
static inline int lanes_per_pixel_v(int C, int V) {
  int l = 1;
  while (l < 32 && l * 2 * V <= C) l *= 2;
  return l;
}
// MRFA_WARP_RUN: 0 = plain (4 loads per pixel) kernels everywhere, 1 = run-walk for >= 100k output pixels, 2 (default) =
// run-walk for every size; MRFA_WARP_VEC=4 forces the 128-bit run-walk (A/B measurements only)
static int warp_run_mode() {
  static const int v = []() { const char* e = getenv("MRFA_WARP_RUN"); return e ? atoi(e) : 2; }();
  return v;
}
static int warp_vec_pref() {
  static const int v = []() { const char* e = getenv("MRFA_WARP_VEC"); return e ? atoi(e) : 8; }();
  return v;
}
static inline int pick_vec(int C, const void* a, const void* b, const void* c, int64_t ostride) {
  const bool al32 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 31) == 0;
  return (warp_vec_pref() == 8 && C % 8 == 0 && ostride % 8 == 0 && al32) ? 8 : 4;
}

// pixels per warp: enough warps (>= ~16k) to fill 148 SMs a few times over at the mid-sized levels, never fewer pixels than
// lane groups.  MRFA_WARP_PW overrides (A/B measurements only).
// `max_run`: cap on the pixels one lane group walks (0 = none).  The dual warp makes two passes over its pixels and wants the
// input neighbourhood of pass 1 still in L1 for pass 2: measured best at runs of 8 (pw = 8 * G), the single warp at 32.
static inline int pick_pw(int64_t pixels, int lanes_pp, int max_run) {
  static const int forced = []() { const char* e = getenv("MRFA_WARP_PW"); return e ? atoi(e) : 0; }();
  const int G = 32 / lanes_pp;
  int pw = 32;
  if (max_run > 0 && max_run * G < pw) pw = max_run * G;
  while (pw > 1 && pixels / pw < 16384) pw >>= 1;
  if (forced > 0) pw = forced;
  if (pw < G) pw = G;
  if (pw > 32) pw = 32;
  return pw;
}

template <int MODE, int PAD, bool ADD_ID>
static int launch_fwd_nhwc_run(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N, int C,
                               int H, int W, int Ho, int Wo, int div, cudaStream_t st) {
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int V = pick_vec(C, in, out, out, C);
  const int pw = pick_pw(pixels, lanes_per_pixel_v(C, V), 0);
  dim3 g((unsigned)cdiv64(cdiv64(pixels, pw), kThreads / 32));
#define MRFA_GSR_CASE(VV, L)                                                                                          \
  case L: grid_sample_fwd_nhwc_run_kernel<MODE, PAD, ADD_ID, VV, L><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div, pw); break;
  if (V == 8) {
    switch (lanes_per_pixel_v(C, 8)) { MRFA_GSR_CASE(8, 1) MRFA_GSR_CASE(8, 2) MRFA_GSR_CASE(8, 4) MRFA_GSR_CASE(8, 8) MRFA_GSR_CASE(8, 16) MRFA_GSR_CASE(8, 32) }
  } else {
    switch (lanes_per_pixel_v(C, 4)) { MRFA_GSR_CASE(4, 1) MRFA_GSR_CASE(4, 2) MRFA_GSR_CASE(4, 4) MRFA_GSR_CASE(4, 8) MRFA_GSR_CASE(4, 16) MRFA_GSR_CASE(4, 32) }
  }
#undef MRFA_GSR_CASE
  return MRFA_LAUNCH_RESULT();
}

template <int MODE, int PAD, bool ADD_ID>
static int launch_fwd_nhwc_lpp(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N, int C,
                               int H, int W, int Ho, int Wo, int div, cudaStream_t st) {
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const bool small = pixels < kSmallPixels;
  if (warp_run_mode() && (!small || warp_run_mode() == 2))
    return launch_fwd_nhwc_run<MODE, PAD, ADD_ID>(in, grid, gs, out, N, C, H, W, Ho, Wo, div, st);
  static const int unroll = []() { const char* e = getenv("MRFA_WARP_UNROLL"); return e ? atoi(e) : 2; }();
  dim3 g((unsigned)cdiv64(cdiv64(pixels, small ? 4 : 32), kThreads / 32));
#define MRFA_GS_CASE(L)                                                                                            \
  case L:                                                                                                          \
    if (small) grid_sample_fwd_nhwc_kernel<MODE, PAD, ADD_ID, L, 4, 2><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div); \
    else if (unroll == 1) grid_sample_fwd_nhwc_kernel<MODE, PAD, ADD_ID, L, 32, 1><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div); \
    else if (unroll == 4) grid_sample_fwd_nhwc_kernel<MODE, PAD, ADD_ID, L, 32, 4><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div); \
    else grid_sample_fwd_nhwc_kernel<MODE, PAD, ADD_ID, L, 32, 2><<<g, kThreads, 0, st>>>(in, grid, gs, out, N, C, H, W, Ho, Wo, div);    \
    break;
  switch (lanes_per_pixel(C)) {
    MRFA_GS_CASE(1) MRFA_GS_CASE(2) MRFA_GS_CASE(4) MRFA_GS_CASE(8) MRFA_GS_CASE(16) MRFA_GS_CASE(32)
  }
#undef MRFA_GS_CASE
  return MRFA_LAUNCH_RESULT();
}

template <int MODE, int PAD>
static int launch_fwd_nhwc(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N, int C, int H,
                           int W, int Ho, int Wo, int div, int add_id, cudaStream_t st) {
  if (add_id) return launch_fwd_nhwc_lpp<MODE, PAD, true>(in, grid, gs, out, N, C, H, W, Ho, Wo, div, st);
  return launch_fwd_nhwc_lpp<MODE, PAD, false>(in, grid, gs, out, N, C, H, W, Ho, Wo, div, st);
}

template <int MODE, int PAD, bool ADD_ID>
static int launch_bwd_nhwc_run(const float* go, const float* in, const float* grid, mrfa_grid_strides_t gs, float* gi,
                               float* gg, int N, int C, int H, int W, int Ho, int Wo, int div, cudaStream_t st) {
  const int64_t pixels = (int64_t)N * Ho * Wo;
  const int lpp = lanes_per_pixel_v(C, 4);
  const int pw = pick_pw(pixels, lpp, 0);
  dim3 g((unsigned)cdiv64(cdiv64(pixels, pw), kThreads / 32));
#define MRFA_GSB_CASE(L)                                                                                              \
  case L: grid_sample_bwd_nhwc_run_kernel<MODE, PAD, ADD_ID, L><<<g, kThreads, 0, st>>>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div, pw); break;
  switch (lpp) { MRFA_GSB_CASE(1) MRFA_GSB_CASE(2) MRFA_GSB_CASE(4) MRFA_GSB_CASE(8) MRFA_GSB_CASE(16) MRFA_GSB_CASE(32) }
#undef MRFA_GSB_CASE
  return MRFA_LAUNCH_RESULT();
}

// MRFA_BWD_RUN=0 selects the plain backward (4 atomics per pixel and channel quad; A/B measurements only)
static int bwd_run_mode() {
  static const int v = []() { const char* e = getenv("MRFA_BWD_RUN"); return e ? atoi(e) : 1; }();
  return v;
}

template <int MODE, int PAD>
static int launch_bwd_nhwc(const float* go, const float* in, const float* grid, mrfa_grid_strides_t gs, float* gi,
                           float* gg, int N, int C, int H, int W, int Ho, int Wo, int div, int add_id, cudaStream_t st) {
  if (bwd_run_mode()) {
    if (add_id) return launch_bwd_nhwc_run<MODE, PAD, true>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div, st);
    return launch_bwd_nhwc_run<MODE, PAD, false>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div, st);
  }
  dim3 g((unsigned)cdiv64((int64_t)N * Ho * Wo * kNhwcGroup, kThreads));
  if (add_id)
    grid_sample_bwd_nhwc_kernel<MODE, PAD, true><<<g, kThreads, 0, st>>>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div);
  else
    grid_sample_bwd_nhwc_kernel<MODE, PAD, false><<<g, kThreads, 0, st>>>(go, in, grid, gs, gi, gg, N, C, H, W, Ho, Wo, div);
  return MRFA_LAUNCH_RESULT();
}

template <int MODE, int PAD>
static int launch_fwd(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out, int N, int C, int H,
                      int W, int Ho, int Wo, int div, int add_id, cudaStream_t st) {
  if (C <= 4 && N <= 65535 && (int64_t)N * Ho * Wo >= (1 << 16)) {        // few channels, many pixels: the latency-hiding variant
    const dim3 gb((unsigned)cdiv64((int64_t)Ho * Wo, kFewThreads * kFewPix), (unsigned)N);
    const FastDiv wd = make_fastdiv((uint32_t)Wo);
#define MRFA_FEWC(CC)                                                                                                    \
  case CC:                                                                                                               \
    if (add_id) grid_sample_fwd_fewc_kernel<MODE, PAD, true, CC><<<gb, kFewThreads, 0, st>>>(in, grid, gs, out, H, W, Ho, Wo, wd, div); \
    else grid_sample_fwd_fewc_kernel<MODE, PAD, false, CC><<<gb, kFewThreads, 0, st>>>(in, grid, gs, out, H, W, Ho, Wo, wd, div);       \
    break;
    switch (C) { MRFA_FEWC(1) MRFA_FEWC(2) MRFA_FEWC(3) MRFA_FEWC(4) }
#undef MRFA_FEWC
    return MRFA_LAUNCH_RESULT();
  }
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
                                    int padding_mode, int add_identity, int channels_last, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && grid && out);
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && in_batch_div >= 1);
  MRFA_CHECK_ARG(N % in_batch_div == 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31) && (int64_t)Ho * Wo < (1ll << 31));
  if (N == 0) return 0;
  if (channels_last) {
    MRFA_CHECK_SHAPE(C % 4 == 0 && (int64_t)H * W * C < (1ll << 31));
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) != 0) return MRFA_E_ALIGN;
    DISPATCH_MODE_PAD(launch_fwd_nhwc, in, grid, gs, out, N, C, H, W, Ho, Wo, in_batch_div, add_identity, as_stream(stream));
  }
  DISPATCH_MODE_PAD(launch_fwd, in, grid, gs, out, N, C, H, W, Ho, Wo, in_batch_div, add_identity, as_stream(stream));
}

extern "C" int mrfa_grid_sample_bwd(const float* grad_out, const float* in, const float* grid,
                                    mrfa_grid_strides_t gs, float* grad_in, float* grad_grid, int N, int C, int H,
                                    int W, int Ho, int Wo, int in_batch_div, int coord_mode, int padding_mode,
                                    int add_identity, int channels_last, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_out && in && grid && (grad_in || grad_grid));
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && in_batch_div >= 1);
  MRFA_CHECK_ARG(N % in_batch_div == 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31) && (int64_t)Ho * Wo < (1ll << 31));
  if (N == 0) return 0;
  if (channels_last) {
    MRFA_CHECK_SHAPE(C % 4 == 0);
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(grad_out) | reinterpret_cast<uintptr_t>(grad_in)) & 15) != 0)
      return MRFA_E_ALIGN;
    DISPATCH_MODE_PAD(launch_bwd_nhwc, grad_out, in, grid, gs, grad_in, grad_grid, N, C, H, W, Ho, Wo, in_batch_div,
                      add_identity, as_stream(stream));
  }
  DISPATCH_MODE_PAD(launch_bwd, grad_out, in, grid, gs, grad_in, grad_grid, N, C, H, W, Ho, Wo, in_batch_div,
                    add_identity, as_stream(stream));
}

extern "C" int mrfa_dual_warp_fwd(const float* in, const float* flow, const float* prior_grid, float* out_refined,
                                  float* out_coarse, int N, int C, int H, int W, int channels_last,
                                  int64_t coarse_pixel_stride, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(coarse_pixel_stride == 0 || (channels_last && coarse_pixel_stride >= C && coarse_pixel_stride % 4 == 0));
  MRFA_CHECK_ARG(in && flow && prior_grid && out_refined && out_coarse);
  MRFA_CHECK_ARG(N >= 0 && C > 0 && H > 0 && W > 0);
  MRFA_CHECK_SHAPE((int64_t)H * W < (1ll << 31));
  if (N == 0) return 0;
  if (channels_last) {
    MRFA_CHECK_SHAPE(C % 4 == 0 && (int64_t)H * W * C < (1ll << 31));
    if (((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out_refined) | reinterpret_cast<uintptr_t>(out_coarse)) & 15) != 0)
      return MRFA_E_ALIGN;
    const int64_t pixels = (int64_t)N * H * W;
    const bool small = pixels < kSmallPixels;
    dim3 gn((unsigned)cdiv64(cdiv64(pixels, small ? 4 : 32), kThreads / 32));
    cudaStream_t st = as_stream(stream);
    const int64_t cs = coarse_pixel_stride > 0 ? coarse_pixel_stride : C;
    if (warp_run_mode() && (!small || warp_run_mode() == 2)) {
      const int V = pick_vec(C, in, out_refined, out_coarse, cs);
      const int pw = pick_pw(pixels, lanes_per_pixel_v(C, V), 8);
      dim3 gr((unsigned)cdiv64(cdiv64(pixels, pw), kThreads / 32));
      static const bool occ4 = []() { const char* e = getenv("MRFA_WARP_OCC4"); return e ? atoi(e) != 0 : true; }();
#define MRFA_DWR_CASE(VV, L)                                                                                         \
  case L:                                                                                                            \
    if (occ4) dual_warp_fwd_nhwc_run_kernel<VV, L, true><<<gr, kThreads, 0, st>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W, cs, pw); \
    else dual_warp_fwd_nhwc_run_kernel<VV, L, false><<<gr, kThreads, 0, st>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W, cs, pw);   \
    break;
      if (V == 8) {
        switch (lanes_per_pixel_v(C, 8)) { MRFA_DWR_CASE(8, 1) MRFA_DWR_CASE(8, 2) MRFA_DWR_CASE(8, 4) MRFA_DWR_CASE(8, 8) MRFA_DWR_CASE(8, 16) MRFA_DWR_CASE(8, 32) }
      } else {
        switch (lanes_per_pixel_v(C, 4)) { MRFA_DWR_CASE(4, 1) MRFA_DWR_CASE(4, 2) MRFA_DWR_CASE(4, 4) MRFA_DWR_CASE(4, 8) MRFA_DWR_CASE(4, 16) MRFA_DWR_CASE(4, 32) }
      }
#undef MRFA_DWR_CASE
      return MRFA_LAUNCH_RESULT();
    }
#define MRFA_DW_CASE(L)                                                                                              \
  case L:                                                                                                            \
    if (small) dual_warp_fwd_nhwc_kernel<L, 4><<<gn, kThreads, 0, st>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W, cs); \
    else dual_warp_fwd_nhwc_kernel<L, 32><<<gn, kThreads, 0, st>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W, cs);   \
    break;
    switch (lanes_per_pixel(C)) {
      MRFA_DW_CASE(1) MRFA_DW_CASE(2) MRFA_DW_CASE(4) MRFA_DW_CASE(8) MRFA_DW_CASE(16) MRFA_DW_CASE(32)
    }
#undef MRFA_DW_CASE
    return MRFA_LAUNCH_RESULT();
  }
  dim3 g((unsigned)cdiv64((int64_t)N * H * W, kWarpPix), (unsigned)cdiv64(C, kCPT * (kThreads / kWarpPix)));
  dual_warp_fwd_kernel<<<g, kThreads, 0, as_stream(stream)>>>(in, flow, prior_grid, out_refined, out_coarse, N, C, H, W);
  return MRFA_LAUNCH_RESULT();
}
