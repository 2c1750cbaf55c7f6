// Shared device helpers for libmrfa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/mrfa_b200.h"

#define MRFA_CHECK_ARG(cond) do { if (!(cond)) return MRFA_E_BADARG; } while (0)
#define MRFA_CHECK_SHAPE(cond) do { if (!(cond)) return MRFA_E_SHAPE; } while (0)
#define MRFA_LAUNCH_RESULT() ((int)cudaGetLastError())

static inline cudaStream_t as_stream(mrfa_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

namespace mrfa {

// ---- coordinate conventions ---------------------------------------
// Replays the exact fp32 operation order of the reference so results stay inside 1e-5:
//  PIXEL   : g = 2*x/(W-1) - 1            (util.py:30-31), then align_corners=True un-normalise
//  NORM_ACT: ((g + 1) / 2) * (size - 1)   (ATen grid_sampler_unnormalize, align_corners)
//  NORM_ACF: ((g + 1) * size - 1) / 2
template <int MODE>
__device__ __forceinline__ float to_pixel(float g, int size) {
  // x / 2 is written as x * 0.5f: bit-identical, without the IEEE division sequence
  if (MODE == MRFA_COORD_PIXEL) {
    g = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, g), (float)(size - 1)), 1.f);
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  } else if (MODE == MRFA_COORD_NORM_ACT) {
    return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), (float)(size - 1));
  } else {
    // ATen evaluates (g + 1) * size - 1 with ONE rounding: nvcc contracts it to an FMA in grid_sampler_unnormalize and the
    // vectorised CPU kernel does the same with (g + 1) * (size / 2) - 0.5 (measured: stock CUDA == stock CPU bit for bit,
    // a separately rounded product differs by up to 4e-5 on a 200-pixel map).  Replay the FMA.
    return __fmul_rn(__fmaf_rn(__fadd_rn(g, 1.f), (float)size, -1.f), 0.5f);
  }
}

// d(pixel)/d(grid value) for the backward pass.
template <int MODE>
__device__ __forceinline__ float to_pixel_grad(int size) {
  if (MODE == MRFA_COORD_PIXEL) return 1.f;
  if (MODE == MRFA_COORD_NORM_ACT) return (float)(size - 1) / 2.f;
  return (float)size / 2.f;
}

// ATen reflect_coordinates + clip (padding_mode="reflection"), returns d(out)/d(in) in *mult.
__device__ __forceinline__ float reflect_coord(float x, int twice_low, int twice_high, float* mult) {
  if (twice_low == twice_high) { *mult = 0.f; return 0.f; }
  float mn = (float)twice_low / 2.f;
  float span = (float)(twice_high - twice_low) / 2.f;
  float m = 1.f;
  x = x - mn;
  if (x < 0.f) { x = -x; m = -1.f; }
  float extra = fmodf(x, span);
  int flips = (int)floorf(x / span);
  if (flips % 2 == 0) { *mult = m; return extra + mn; }
  *mult = -m;
  return span - extra + mn;
}

template <int MODE, int PAD>
__device__ __forceinline__ float source_index(float g, int size, float* mult) {
  float p = to_pixel<MODE>(g, size);
  float m = to_pixel_grad<MODE>(size);
  if (PAD == MRFA_PAD_REFLECTION) {
    float r;
    if (MODE == MRFA_COORD_NORM_ACF) p = reflect_coord(p, -1, 2 * size - 1, &r);
    else p = reflect_coord(p, 0, 2 * (size - 1), &r);
    m *= r;
    if (p <= 0.f) { p = 0.f; m = 0.f; }                      // clip_coordinates_set_grad
    else if (p >= (float)(size - 1)) { p = (float)(size - 1); m = 0.f; }
  }
  *mult = m;
  return p;
}

// Four bilinear taps of one sample: flat offsets inside an H x W plane (clamped to a legal
// address) and weights (zeroed for out-of-image taps = zeros padding).
struct Taps {
  int o_nw, o_ne, o_sw, o_se;
  float w_nw, w_ne, w_sw, w_se;
};

__device__ __forceinline__ Taps make_taps(float ix, float iy, int H, int W) {
  Taps t;
  // clamp so the int conversion is defined for wild / non-finite coordinates
  float fx = floorf(ix), fy = floorf(iy);
  bool finite = (fabsf(ix) < 1e9f) && (fabsf(iy) < 1e9f);   // false for NaN / inf
  int x0 = finite ? (int)fx : -2;
  int y0 = finite ? (int)fy : -2;
  int x1 = x0 + 1, y1 = y0 + 1;
  float ax = ix - fx, ay = iy - fy;                          // == ix - ix_nw
  float bx = (fx + 1.f) - ix, by = (fy + 1.f) - iy;          // == ix_se - ix
  bool vx0 = (x0 >= 0) & (x0 < W), vx1 = (x1 >= 0) & (x1 < W);
  bool vy0 = (y0 >= 0) & (y0 < H), vy1 = (y1 >= 0) & (y1 < H);
  int cx0 = min(max(x0, 0), W - 1), cx1 = min(max(x1, 0), W - 1);
  int cy0 = min(max(y0, 0), H - 1), cy1 = min(max(y1, 0), H - 1);
  t.o_nw = cy0 * W + cx0; t.o_ne = cy0 * W + cx1;
  t.o_sw = cy1 * W + cx0; t.o_se = cy1 * W + cx1;
  t.w_nw = (finite & vx0 & vy0) ? bx * by : 0.f;
  t.w_ne = (finite & vx1 & vy0) ? ax * by : 0.f;
  t.w_sw = (finite & vx0 & vy1) ? bx * ay : 0.f;
  t.w_se = (finite & vx1 & vy1) ? ax * ay : 0.f;
  return t;
}

// Element offset of (y, x) inside one H x W map (include/mrfa_b200.h, "Map layouts").
//   ROWMAJOR: y * W + x.
//   TILED   : 64-byte tiles of 4 rows x 8 columns (bf16), so an (2r+2)^2 footprint touches 4-6 DRAM granules
//             instead of 8 separate rows.  Level 0 groups 2 x 2 tiles into a 16 x 8-pixel super-tile (the unit
//             one level-1 tile is pooled from in the GEMM epilogue); level 1 orders plain tiles row-major.
template <bool TILED>
__host__ __device__ __forceinline__ int64_t map_offset(int lvl, int y, int x, int Wl) {
  if (!TILED) return (int64_t)y * Wl + x;
  if (lvl == 0)
    return (int64_t)((y >> 3) * (Wl >> 4) + (x >> 4)) * 128 + ((y >> 2) & 1) * 64 + ((x >> 3) & 1) * 32 + (y & 3) * 8 + (x & 7);
  return (int64_t)((y >> 2) * (Wl >> 3) + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7);
}

// Division by a launch-time constant without the integer-division sequence (the streaming kernels decompose a flat
// element index into (channel quad, x, y, sample) per element: with 64-bit `/` and `%` that was most of their issued
// instructions).  q = umulhi(x, mul) >> shr is exact for x < 2^31 (ceil(2^(31+L) / d) with L = ceil(log2 d)).
struct FastDiv {
  uint32_t d, mul, shr;
};
static inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f;
  f.d = d; f.mul = 0; f.shr = 0;
  if (d > 1) {
    uint32_t l = 0;
    while ((1u << l) < d) ++l;                                   // ceil(log2 d), 1..31 for the sizes that occur
    const uint32_t p = 31 + l;
    f.mul = (uint32_t)((((uint64_t)1 << p) + d - 1) / d);
    f.shr = p - 32;
  }
  return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t x, const FastDiv& f) {
  return f.d == 1 ? x : (__umulhi(x, f.mul) >> f.shr);
}
// i = ((n * h + y) * w + x) * c + ci  ->  (ci, x, y, n) and pix = i / c; 32-bit multiply-shift when i < 2^31
struct IndexSplit {
  FastDiv c, w, h;
};
__device__ __forceinline__ void split_index(int64_t i, const IndexSplit& s, int& ci, int& x, int& y, int64_t& n, int64_t& pix) {
  if (i < ((int64_t)1 << 31)) {
    const uint32_t u = (uint32_t)i;
    const uint32_t p = fast_div(u, s.c);
    const uint32_t r = fast_div(p, s.w);
    const uint32_t nn = fast_div(r, s.h);
    ci = (int)(u - p * s.c.d); x = (int)(p - r * s.w.d); y = (int)(r - nn * s.h.d);
    n = nn; pix = p;
  } else {
    ci = (int)(i % s.c.d);
    pix = i / s.c.d;
    x = (int)(pix % s.w.d);
    y = (int)((pix / s.w.d) % s.h.d);
    n = pix / ((int64_t)s.w.d * s.h.d);
  }
}
static inline IndexSplit make_index_split(int c, int w, int h) {
  IndexSplit s;
  s.c = make_fastdiv((uint32_t)c); s.w = make_fastdiv((uint32_t)w); s.h = make_fastdiv((uint32_t)h);
  return s;
}

// x / d for a divisor d that is constant over many quotients, with rcp = RN(1 / d) (__frcp_rn): q0 = RN(x * rcp),
// r = x - q0 * d (exact in one FMA), q = RN(q0 + r * rcp) is the correctly rounded quotient (Markstein) -- the same bits as
// __fdiv_rn for every finite x whose quotient is a normal number, without the reciprocal refinement and the slow-path
// check of the IEEE division sequence; x = 0 gives 0, non-finite x gives NaN instead of inf.  (Checked against the
// correctly rounded quotient on 1.1e8 (x, d) pairs on the host.)
__device__ __forceinline__ float div_by_const(float x, float d, float rcp) {
  const float q0 = __fmul_rn(x, rcp);
  const float r = __fmaf_rn(-q0, d, x);
  return __fmaf_rn(r, rcp, q0);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// identity grid value, bit-exact with util.make_coordinate_grid: 2*(j/(n-1)) - 1
__device__ __forceinline__ float norm_coord(int j, int n) {
  return __fsub_rn(__fmul_rn(2.f, __fdiv_rn((float)j, (float)(n - 1))), 1.f);
}

}  // namespace mrfa
