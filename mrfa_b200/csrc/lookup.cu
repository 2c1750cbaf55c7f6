// Correlation pyramid lookup (K4 of SURVEY.md, CorrBlock.__call__ raft.py:23-48) and the
// generic level-1 build (raft.py:19-21).
//
// Per query the (2r+1)^2 window of a level shares one fractional offset, so its bilinear taps
// live in an (2r+2)^2 integer footprint.  A block takes 32 consecutive queries of one sample:
//   phase 1  warp-per-query gather of the two footprints (each footprint row is one 16/32-byte
//            contiguous piece of that query's own volume row) into shared memory;
//   phase 2  lane-per-query evaluation of the 2*(2r+1)^2 outputs from shared memory, so every
//            store is a 128-byte line of the (B, 98, Q) output.
// HBM-bound: algorithmic bytes per query = 2 levels * 64 * elt + 8 (coords) + 98*4 (out).
#include "common.cuh"

namespace mrfa {

constexpr int kQPB = 32;              // queries per block
constexpr int kLookupThreads = 128;
constexpr int kMaxR = 4;

template <typename T> __device__ __forceinline__ float ld_elem(const T* p);
template <> __device__ __forceinline__ float ld_elem<float>(const float* p) { return __ldg(p); }
template <> __device__ __forceinline__ float ld_elem<__nv_bfloat16>(const __nv_bfloat16* p) {
  return __bfloat162float(__ldg(p));
}

struct QueryGeom {
  int x0, y0;       // integer position of footprint cell (0,0)
  float fx, fy;     // shared fractional offsets
};

// centre of the window on level `lvl`, replaying util.bilinear_sampler's coordinate round trip
__device__ __forceinline__ QueryGeom query_geom(float cx, float cy, int lvl, int Hl, int Wl, int r) {
  const float s = (lvl == 0) ? 1.f : 2.f;
  const float px = to_pixel<MRFA_COORD_PIXEL>(__fdiv_rn(cx, s), Wl);
  const float py = to_pixel<MRFA_COORD_PIXEL>(__fdiv_rn(cy, s), Hl);
  QueryGeom g;
  const bool fin = (fabsf(px) < 1e8f) && (fabsf(py) < 1e8f);
  const float flx = floorf(px), fly = floorf(py);
  g.x0 = fin ? (int)flx - r : -(1 << 20);
  g.y0 = fin ? (int)fly - r : -(1 << 20);
  g.fx = fin ? px - flx : 0.f;
  g.fy = fin ? py - fly : 0.f;
  return g;
}

template <typename T, int R, bool TILED>
__global__ void __launch_bounds__(kLookupThreads)
corr_lookup_fwd_kernel(const T* __restrict__ level0, const T* __restrict__ level1, const float* __restrict__ coords,
                       float* __restrict__ out, int Q, int H, int W, int64_t map_batch_stride, int64_t row_offset,
                       int out_channels_last) {
  constexpr int n = 2 * R + 1, F = n + 1, FF = F * F;
  constexpr int kStride = 2 * FF + 1;                 // odd -> conflict-free lane-per-query reads
  __shared__ float foot[kQPB * kStride];
  __shared__ float frac[kQPB][2][2][n];               // per-tap fractional offsets [level][axis][tap] (see lookup_geometry)
  __shared__ int org[kQPB][4];                        // footprint origin (x0, y0) per level

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * kQPB;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int H1 = H / 2, W1 = W / 2;

  // ---- phase 0: one thread per (query, level, axis).  The reference adds the integer window offsets before the
  //      coordinate round trip of util.bilinear_sampler (raft.py:31-37), so every tap gets its own fraction.
  {
    const int t = threadIdx.x;
    const int qi = t >> 2, lvl = (t >> 1) & 1, axis = t & 1;
    if (q0 + qi < Q) {
      const float c = __ldg(coords + ((int64_t)b * 2 + axis) * Q + q0 + qi);
      const int size = axis ? (lvl ? H1 : H) : (lvl ? W1 : W);
      const float cs = __fdiv_rn(c, lvl ? 2.f : 1.f);
      const float pc = to_pixel<MRFA_COORD_PIXEL>(cs, size);
      const bool fin = fabsf(pc) < 1e8f;
      const int base = fin ? (int)floorf(pc) - R : -(1 << 20);
      org[qi][2 * lvl + axis] = base;
#pragma unroll
      for (int a = 0; a < n; ++a) {
        const float pa = to_pixel<MRFA_COORD_PIXEL>(__fadd_rn(cs, (float)(a - R)), size);
        frac[qi][lvl][axis][a] = fin ? pa - (float)(base + a) : 0.f;
      }
    }
  }
  __syncthreads();

  // ---- phase 1: warp per query, gather both footprints.  kGroup queries are gathered per
  //      round so every lane has 2*kGroup*ceil(FF/32) independent loads in flight before the
  //      first shared-memory store: the gather is latency-bound (each 2/4-byte load pulls its
  //      own 32-byte sector), so bytes in flight per SM set the achieved bandwidth. ----
  constexpr int kWarps = kLookupThreads / 32;
  constexpr int kPerWarp = kQPB / kWarps;
  constexpr int kGroup = 4;
  constexpr int kIter = (FF + 31) / 32;
  static_assert(kPerWarp % kGroup == 0, "query grouping");
#pragma unroll 1
  for (int g0 = 0; g0 < kPerWarp; g0 += kGroup) {
    float v[kGroup][2][kIter];
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      const int qi = warp + (g0 + g) * kWarps;
      const bool live = q0 + qi < Q;
      const int64_t map = (int64_t)b * map_batch_stride + row_offset + q0 + qi;
#pragma unroll
      for (int lvl = 0; lvl < 2; ++lvl) {
        const int Hl = lvl ? H1 : H, Wl = lvl ? W1 : W;
        const int gx0 = org[qi][2 * lvl], gy0 = org[qi][2 * lvl + 1];
        const T* base = (lvl ? level1 : level0) + map * ((int64_t)Hl * Wl);
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
          const int e = lane + 32 * it;
          const int fy_ = e / F, fx_ = e - fy_ * F;
          const int yy = gy0 + fy_, xx = gx0 + fx_;
          float t = 0.f;
          if (live && e < FF && yy >= 0 && yy < Hl && xx >= 0 && xx < Wl) t = ld_elem<T>(base + map_offset<TILED>(lvl, yy, xx, Wl));
          v[g][lvl][it] = t;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < kGroup; ++g) {
      const int qi = warp + (g0 + g) * kWarps;
#pragma unroll
      for (int lvl = 0; lvl < 2; ++lvl)
#pragma unroll
        for (int it = 0; it < kIter; ++it) {
          const int e = lane + 32 * it;
          if (e < FF) foot[qi * kStride + lvl * FF + e] = v[g][lvl][it];
        }
    }
  }
  __syncthreads();

  if (out_channels_last) {
    // ---- phase 2 (NHWC output): warp per query, lanes along the 2*(2r+1)^2 contiguous channels
    for (int qi = warp; qi < kQPB; qi += kLookupThreads / 32) {
      const int q = q0 + qi;
      if (q >= Q) break;
      const float* fq = foot + qi * kStride;
      float* dst = out + ((int64_t)b * Q + q) * (2 * n * n);
      for (int k = lane; k < 2 * n * n; k += 32) {
        const int lvl = k / (n * n), kk = k - lvl * n * n;
        const int a = kk / n, bb = kk - a * n;
        const float fx = frac[qi][lvl][0][a], fy = frac[qi][lvl][1][bb];
        const float* c = fq + lvl * FF + bb * F + a;
        float acc = c[0] * ((1.f - fx) * (1.f - fy));
        acc = fmaf(c[1], fx * (1.f - fy), acc);
        acc = fmaf(c[F], (1.f - fx) * fy, acc);
        acc = fmaf(c[F + 1], fx * fy, acc);
        *(dst + k) = acc;
      }
    }
    return;
  }
  // ---- phase 2: lane = query, warps stride over the output channels ----------------------
  const int q = q0 + lane;
  if (q >= Q) return;
  const float* fq = foot + lane * kStride;
  float* dst = out + (int64_t)b * (2 * n * n) * Q + q;
#pragma unroll
  for (int lvl = 0; lvl < 2; ++lvl) {
    for (int k = warp; k < n * n; k += kLookupThreads / 32) {
      const int a = k / n, bb = k - a * n;             // channel a*n+b: x offset a-r, y offset b-r
      const float fx = frac[lane][lvl][0][a], fy = frac[lane][lvl][1][bb];
      // same evaluation order as ATen: nw, ne, sw, se
      const float w_nw = (1.f - fx) * (1.f - fy), w_ne = fx * (1.f - fy), w_sw = (1.f - fx) * fy, w_se = fx * fy;
      const float* c = fq + lvl * FF + bb * F + a;
      float acc = c[0] * w_nw;
      acc = fmaf(c[1], w_ne, acc);
      acc = fmaf(c[F], w_sw, acc);
      acc = fmaf(c[F + 1], w_se, acc);
      *(dst + (int64_t)(lvl * n * n + k) * Q) = acc;
    }
  }
}


// ---------------------------------------------------------------------------------------------
// Tiled bf16 pyramid maps (the volume written by mrfa_corr_volume for w = 64 / 128), r <= 3.
//
// Persistent blocks walk groups of 32 consecutive queries of one sample, software-pipelined:
//   geometry  one thread per (query, level, axis): the reference evaluates the coordinate round trip of
//             util.bilinear_sampler for every window tap separately (raft.py:31-37 adds the integer offsets BEFORE the
//             normalisation), so each tap a gets its own fractional offset frac_a = pixel(c / 2^lvl + a - r) - (x0 + a);
//             x0 = floor(pixel(c / 2^lvl)) - r is the shared footprint origin.  Bit-for-bit the reference's sample
//             positions, also for coordinates far from the origin where the rounding of c + a differs per tap;
//   gather    one warp instruction fetches BOTH footprints of a query: lane = (level, footprint row, tile column) loads
//             one 16-byte tile row, so a query costs 4-6 64-byte DRAM granules per level instead of 8 scattered row
//             pieces; the loads of group g+1 are issued before group g is evaluated and stay in flight meanwhile;
//   patch     the 8 x 8 footprint of each level is parked in shared memory as fp32, aligned to x0 (conflict-free reads);
//   evaluate  lane = query, warp = (level, half of the window columns): separable bilinear weights, each row's
//             horizontal interpolation shared by the two window rows that use it (2.3 shared loads per output);
//   write     through a staging tile so both output layouts leave as full 128-byte lines (NHWC: the 32 queries of a
//             group are one contiguous 12.5 KB run).
// ---------------------------------------------------------------------------------------------
// two levels x 8 rows x 16 bf16 columns = 128 words per query, + 4: a multiple of 4 so a gathering lane parks its 16-byte
// tile row with ONE conflict-free st.shared.v4.  The evaluating lanes (lane = query) then sit 4 banks apart, so queries
// 8, 16, 24 apart would collide on the same word: the 4 words of every 16-byte chunk are stored XOR-permuted by
// (query >> 3), which spreads those four lanes over the chunk's four banks (conflict-free for equal word indices).
constexpr int kPatchStride = 2 * 8 * 8 + 4;
constexpr int kFracStride = 2 * 2 * 7 + 1;           // [level][axis][tap] fractions per query (odd stride)
constexpr int kGroupT = 8;                           // queries per warp per group = loads in flight per lane
constexpr int kTiledBlocksPerSM = 5;                 // resident persistent blocks per SM (~37 KB of shared memory each)

template <int R>
struct LookupSmem {
  static constexpr int n = 2 * R + 1;
  static constexpr int kOut = 2 * n * n;               // outputs per query, stored back to back (float4 copy-out)
  alignas(16) float outs[kQPB * kOut];
  alignas(16) uint32_t patch[kQPB * kPatchStride];                 // per query: [level][row 0..7][8 words = tile columns tx, tx+1 as loaded]
  float frac[2][kQPB * kFracStride];
  // per (query, level): {element offset of tile column tx0 / tx0 + 1 inside a tile row (-1: not needed or outside the
  // map), footprint origin y0, footprint origin x0}; dead queries carry y0 = -2^20 so every row fails the range test
  alignas(16) int4 org[2][kQPB][2];
};

// to_pixel<MRFA_COORD_PIXEL> with div_by_const (common.cuh): the divisor size - 1 is constant per (level, axis)
__device__ __forceinline__ float to_pixel_pix(float g, float size_m1, float rcp) {
  g = __fsub_rn(div_by_const(__fmul_rn(2.f, g), size_m1, rcp), 1.f);
  return __fmul_rn(__fmul_rn(__fadd_rn(g, 1.f), 0.5f), size_m1);
}

template <int R>
__device__ __forceinline__ void lookup_geometry(LookupSmem<R>& sm, int buf, const float* __restrict__ coords, int b, int q0,
                                                int Q, int H, int W) {
  constexpr int n = 2 * R + 1, F = n + 1;
  const int t = threadIdx.x;                         // 128 threads = 32 queries x 2 levels x 2 axes
  const int qi = t >> 2, lvl = (t >> 1) & 1, axis = t & 1;
  const int q = q0 + qi;
  int* o = reinterpret_cast<int*>(&sm.org[buf][qi][lvl]);
  if (q >= Q) {
    if (axis) o[2] = -(1 << 20); else { o[0] = -1; o[1] = -1; o[3] = 0; }
    return;
  }
  const float c = __ldg(coords + ((int64_t)b * 2 + axis) * Q + q);
  const int size = (axis ? H : W) >> lvl;
  const float sm1 = (float)(size - 1), rcp = __frcp_rn(sm1);
  const float cs = __fmul_rn(c, lvl ? 0.5f : 1.f);   // coords / 2**lvl (raft.py:34), exact
  float* f = sm.frac[buf] + qi * kFracStride + (lvl * 2 + axis) * 7;
  float pa[n];
#pragma unroll
  for (int a = 0; a < n; ++a) pa[a] = to_pixel_pix(__fadd_rn(cs, (float)(a - R)), sm1, rcp);
  const float pc = pa[R];                            // tap R adds 0: the window centre itself
  const bool fin = fabsf(pc) < 1e8f;
  const int base = fin ? (int)floorf(pc) - R : -(1 << 20);
#pragma unroll
  for (int a = 0; a < n; ++a) f[a] = fin ? pa[a] - (float)(base + a) : 0.f;
  if (axis) {
    o[2] = base;
  } else {
    // tile columns tx0 = floor(x0 / 8) and tx0 + 1 (the second only when the footprint crosses an 8-column boundary):
    // element offset of the tile inside its tile row (include/mrfa_b200.h "Map layouts": level 0 in 2 x 2 super-tiles)
    const int tiles_w = size >> 3, tx0 = base >> 3;   // arithmetic shift: floor for negative origins
    const bool two = (base & 7) + F > 8;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int tx = tx0 + half;
      const bool ok = ((unsigned)tx < (unsigned)tiles_w) & ((half == 0) | two);
      o[half] = ok ? (lvl ? tx << 5 : ((tx >> 1) << 7) + ((tx & 1) << 5)) : -1;
    }
    o[3] = base;
  }
}

template <int R>
__device__ __forceinline__ void lookup_issue_loads(const LookupSmem<R>& sm, int buf, uint4 (&v)[kGroupT],
                                                   const __nv_bfloat16* __restrict__ level0,
                                                   const __nv_bfloat16* __restrict__ level1, int b, int q0, int H,
                                                   int W, int64_t map_batch_stride, int64_t row_offset) {
  constexpr int F = 2 * R + 2;
  constexpr int kWarps = kLookupThreads / 32;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int lvl = lane >> 4, r = (lane >> 1) & 7, half = lane & 1;
  const int Hl = (r < F) ? (H >> lvl) : 0, tiles_w = (W >> lvl) >> 3;       // footprint rows past F never load
  const int map_elems = (H >> lvl) * (W >> lvl);
  // tile-row pitch and row-in-tile terms of the map offset: level 0 rows of 16 x 8-pixel super-tiles, level 1 rows of tiles
  const int shift = lvl ? 2 : 3, pitch = lvl ? tiles_w << 5 : tiles_w << 6, mid = lvl ? 0 : 64;
  // map of query (b, q0 + warp) on this lane's level; the other queries of the warp are 4 maps apart
  const __nv_bfloat16* base = (lvl ? level1 : level0) + ((int64_t)b * map_batch_stride + row_offset + q0 + warp) * map_elems;
#pragma unroll
  for (int g = 0; g < kGroupT; ++g) {
    const int4 o = sm.org[buf][warp + g * kWarps][lvl];
    const int y = o.z + r, xo = half ? o.y : o.x;
    const bool ok = ((unsigned)y < (unsigned)Hl) & (xo >= 0);
    const int off = (y >> shift) * pitch + (((y >> 2) & 1) ? mid : 0) + ((y & 3) << 3) + xo + g * (kWarps * map_elems);
    v[g] = make_uint4(0u, 0u, 0u, 0u);
    if (ok) v[g] = __ldg(reinterpret_cast<const uint4*>(base + off));
  }
}

// Horizontal interpolations of one patch row: h[a] = e[dx + a] * (1 - fx[a]) + e[dx + a + 1] * fx[a].
template <int R>
__device__ __forceinline__ void lookup_hrow(const uint32_t* __restrict__ pr, int w0, uint32_t sh, int kx,
                                            const float (&fx)[2 * R + 1], const float (&wx0)[2 * R + 1],
                                            float (&h)[2 * R + 1]) {
  constexpr int n = 2 * R + 1, NP = (n + 2) / 2;       // NP aligned pairs hold elements dx .. dx + n
  uint32_t w[NP + 1];
#pragma unroll
  for (int i = 0; i <= NP; ++i) w[i] = pr[(w0 + i) ^ kx];            // chunk-permuted word positions (see kPatchStride)
  float e[2 * NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const uint32_t p = __funnelshift_r(w[i], w[i + 1], sh);          // elements dx + 2i (low half), dx + 2i + 1 (high half)
    e[2 * i] = __uint_as_float(p << 16);
    e[2 * i + 1] = __uint_as_float(p & 0xffff0000u);
  }
#pragma unroll
  for (int a = 0; a < n; ++a) h[a] = fmaf(e[a + 1], fx[a], e[a] * wx0[a]);
}

// Output rows [B0, B1) of one level for the lane's query (same operation order as the column-wise form it replaced).
template <int R, int B0, int B1>
__device__ __forceinline__ void lookup_eval_rows(const uint32_t* __restrict__ pq, const float* __restrict__ fq,
                                                 float* __restrict__ oq, int dx, int kx) {
  constexpr int n = 2 * R + 1;
  const int w0 = dx >> 1;
  const uint32_t sh = (dx & 1) * 16;
  float fx[n], wx0[n];
#pragma unroll
  for (int a = 0; a < n; ++a) { fx[a] = fq[a]; wx0[a] = 1.f - fx[a]; }
  float h[2][n];
  lookup_hrow<R>(pq + B0 * 8, w0, sh, kx, fx, wx0, h[0]);
#pragma unroll
  for (int bb = B0; bb < B1; ++bb) {
    float (&hp)[n] = h[(bb - B0) & 1];
    float (&hn)[n] = h[(bb - B0 + 1) & 1];
    lookup_hrow<R>(pq + (bb + 1) * 8, w0, sh, kx, fx, wx0, hn);
    const float fy = fq[7 + bb], wy0 = 1.f - fy;
#pragma unroll
    for (int a = 0; a < n; ++a) oq[a * n + bb] = fmaf(hn[a], fy, hp[a] * wy0);
  }
}

template <int R>
__global__ void __launch_bounds__(kLookupThreads)
corr_lookup_fwd_tiled_kernel(const __nv_bfloat16* __restrict__ level0, const __nv_bfloat16* __restrict__ level1,
                             const float* __restrict__ coords, float* __restrict__ out, int B, int Q, int H, int W,
                             int64_t map_batch_stride, int64_t row_offset, int out_channels_last) {
  constexpr int n = 2 * R + 1, F = n + 1, NN = n * n;
  static_assert(F <= 8, "the footprint must fit 8 rows x 2 tile columns");
  constexpr int kWarps = kLookupThreads / 32;
  constexpr int kOut = LookupSmem<R>::kOut;
  static_assert(kQPB == kWarps * kGroupT, "one gather round per group");
  __shared__ LookupSmem<R> sm;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // groups of 32 queries, numbered sample-major; (b, gi) advance by the grid size without 64-bit divisions
  const int groups_per_b = (Q + kQPB - 1) / kQPB;
  const int total = B * groups_per_b;                        // < 2^31 (checked by the launcher)
  int grp = blockIdx.x;
  if (grp >= total) return;
  const int step_b = gridDim.x / groups_per_b, step_g = gridDim.x - step_b * groups_per_b;
  int b = grp / groups_per_b, gi = grp - b * groups_per_b;

  uint4 v[kGroupT];
  int buf = 0;
  lookup_geometry<R>(sm, 0, coords, b, gi * kQPB, Q, H, W);
  __syncthreads();
  lookup_issue_loads<R>(sm, 0, v, level0, level1, b, gi * kQPB, H, W, map_batch_stride, row_offset);
  for (; grp < total; grp += gridDim.x, buf ^= 1) {
    const int q0 = gi * kQPB;
    const int cur_b = b;
    // ---- park the gathered tile rows as loaded (packed bf16): word (level*64 + row*8 + half*4 + i) == lane*4 + i
#pragma unroll
    for (int g = 0; g < kGroupT; ++g) {
      uint4* d = reinterpret_cast<uint4*>(sm.patch + (warp + g * kWarps) * kPatchStride + lane * 4);
      constexpr int kW4 = kWarps == 4 ? 1 : 0;
      static_assert(kW4 == 1, "the chunk permutation below assumes query = warp + 4 * g");
      switch ((g >> 1) & 3) {                           // (query >> 3) with query = warp + 4 g, warp < 4: static per g
        case 0: *d = make_uint4(v[g].x, v[g].y, v[g].z, v[g].w); break;
        case 1: *d = make_uint4(v[g].y, v[g].x, v[g].w, v[g].z); break;
        case 2: *d = make_uint4(v[g].z, v[g].w, v[g].x, v[g].y); break;
        default: *d = make_uint4(v[g].w, v[g].z, v[g].y, v[g].x); break;
      }
    }
    // ---- geometry + loads of the next group: in flight while this group is evaluated and written
    const bool more = grp + (int)gridDim.x < total;
    b += step_b; gi += step_g;
    if (gi >= groups_per_b) { gi -= groups_per_b; ++b; }
    if (more) lookup_geometry<R>(sm, buf ^ 1, coords, b, gi * kQPB, Q, H, W);
    __syncthreads();                                  // patches of this group and geometry of the next are visible
    if (more) lookup_issue_loads<R>(sm, buf ^ 1, v, level0, level1, b, gi * kQPB, H, W, map_batch_stride, row_offset);
    // ---- evaluate: lane = query, warp = (level, half of the window rows).  Row-wise: the words holding elements
    //      dx .. dx + n of a patch row (dx = x0 & 7: where the footprint starts inside the two loaded tile columns) are
    //      fetched once, aligned by one funnel shift per word pair, widened, and give the n horizontal interpolations of
    //      that row; two consecutive rows give a row of outputs.
    {
      const int lvl = warp >> 1;
      const uint32_t* pq = sm.patch + lane * kPatchStride + lvl * 64;
      const float* fq = sm.frac[buf] + lane * kFracStride + lvl * 14;
      float* oq = sm.outs + lane * kOut + lvl * NN;
      const int dx = sm.org[buf][lane][lvl].w & 7;
      constexpr int nb0 = (n + 1) / 2;
      if (warp & 1) lookup_eval_rows<R, nb0, n>(pq, fq, oq, dx, lane >> 3);
      else lookup_eval_rows<R, 0, nb0>(pq, fq, oq, dx, lane >> 3);
    }
    __syncthreads();
    // ---- write out
    const int nq = min(kQPB, Q - q0);
    if (out_channels_last) {
      // the group's outputs are one contiguous run of nq * 2*n*n floats, in shared memory as in global memory
      float* dst = out + ((int64_t)cur_b * Q + q0) * kOut;
      // float4 copies when the run starts 16-byte aligned (always for even Q); the remainder / unaligned case is scalar
      const int total4 = ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) ? nq * kOut / 4 : 0;
      for (int f4 = threadIdx.x; f4 < total4; f4 += kLookupThreads)
        reinterpret_cast<float4*>(dst)[f4] = reinterpret_cast<const float4*>(sm.outs)[f4];
      for (int e = total4 * 4 + threadIdx.x; e < nq * kOut; e += kLookupThreads) dst[e] = sm.outs[e];
    } else if (lane < nq) {
      float* dst = out + (int64_t)cur_b * kOut * Q + q0 + lane;
      for (int k = warp; k < kOut; k += kWarps) dst[(int64_t)k * Q] = sm.outs[lane * kOut + k];
    }
    // the next iteration's patch stores only touch sm.patch (last read before the barrier above); sm.outs is rewritten
    // after the next barrier, by which time every thread has left this write phase
  }
}

// Backward: footprint-cell gradients are assembled by a gather over the (at most four) window
// taps that touch a cell -- the overlapping-tap pre-reduction -- and leave the block as one
// red.global per in-image cell (128 instead of 392 atomics per query); coordinate gradients
// are reduced over the channel axis with warp shuffles.
template <typename T, int R, bool TILED>
__global__ void __launch_bounds__(kLookupThreads)
corr_lookup_bwd_kernel(const float* __restrict__ grad_out, const T* __restrict__ level0,
                       const T* __restrict__ level1, const float* __restrict__ coords,
                       float* __restrict__ grad_level0, float* __restrict__ grad_level1,
                       float* __restrict__ grad_coords, int Q, int H, int W, int64_t map_batch_stride,
                       int64_t row_offset) {
  constexpr int n = 2 * R + 1, F = n + 1, FF = F * F, NN = n * n;
  constexpr int kGoStride = 2 * NN + 1;
  constexpr int kWarps = kLookupThreads / 32;
  __shared__ float go[kQPB * kGoStride];
  __shared__ float foot[kWarps][FF];

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * kQPB;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int H1 = H / 2, W1 = W / 2;

  // stage grad_out: lanes along q (coalesced), warps along channels
  for (int k = warp; k < 2 * NN; k += kWarps) {
    const int q = q0 + lane;
    go[lane * kGoStride + k] = (q < Q) ? __ldg(grad_out + ((int64_t)b * 2 * NN + k) * Q + q) : 0.f;
  }
  __syncthreads();

  for (int qi = warp; qi < kQPB; qi += kWarps) {          // warp per query
    const int q = q0 + qi;
    if (q >= Q) break;
    const float cx = __ldg(coords + ((int64_t)b * 2 + 0) * Q + q);
    const float cy = __ldg(coords + ((int64_t)b * 2 + 1) * Q + q);
    const int64_t map = (int64_t)b * map_batch_stride + row_offset + q;
    const float* gq = go + qi * kGoStride;
    float gcx = 0.f, gcy = 0.f;
#pragma unroll
    for (int lvl = 0; lvl < 2; ++lvl) {
      const int Hl = lvl ? H1 : H, Wl = lvl ? W1 : W;
      const QueryGeom g = query_geom(cx, cy, lvl, Hl, Wl, R);
      const float fx = g.fx, fy = g.fy;
      const float w_nw = (1.f - fx) * (1.f - fy), w_ne = fx * (1.f - fy), w_sw = (1.f - fx) * fy, w_se = fx * fy;
      float* gl = lvl ? grad_level1 : grad_level0;
      const T* base = (lvl ? level1 : level0) + map * ((int64_t)Hl * Wl);
      __syncwarp();
      for (int e = lane; e < FF; e += 32) {
        const int fy_ = e / F, fx_ = e - fy_ * F;
        const int yy = g.y0 + fy_, xx = g.x0 + fx_;
        const bool inside = yy >= 0 && yy < Hl && xx >= 0 && xx < Wl;
        // (1) scatter into the map: cell (fy_, fx_) collects window taps (b',a) in
        //     {(fy_,fx_), (fy_,fx_-1), (fy_-1,fx_), (fy_-1,fx_-1)}; channel index = a*n + b'
        if (gl != nullptr && inside) {
          float acc = 0.f;
          if (fy_ < n && fx_ < n) acc = fmaf(gq[lvl * NN + fx_ * n + fy_], w_nw, acc);
          if (fy_ < n && fx_ >= 1) acc = fmaf(gq[lvl * NN + (fx_ - 1) * n + fy_], w_ne, acc);
          if (fy_ >= 1 && fx_ < n) acc = fmaf(gq[lvl * NN + fx_ * n + (fy_ - 1)], w_sw, acc);
          if (fy_ >= 1 && fx_ >= 1) acc = fmaf(gq[lvl * NN + (fx_ - 1) * n + (fy_ - 1)], w_se, acc);
          if (acc != 0.f) atomicAdd(gl + map * ((int64_t)Hl * Wl) + map_offset<TILED>(lvl, yy, xx, Wl), acc);
        }
        // (2) stage the map values for the coordinate gradient (zero outside the image)
        if (grad_coords != nullptr) foot[warp][e] = inside ? ld_elem<T>(base + map_offset<TILED>(lvl, yy, xx, Wl)) : 0.f;
      }
      if (grad_coords != nullptr) {
        __syncwarp();
        const float inv = lvl ? 0.5f : 1.f;                 // d(c / 2^lvl)/dc
        for (int k = lane; k < NN; k += 32) {
          const int a = k / n, bb = k - a * n;
          const float* c = &foot[warp][bb * F + a];
          const float dpx = (c[1] - c[0]) * (1.f - fy) + (c[F + 1] - c[F]) * fy;
          const float dpy = (c[F] - c[0]) * (1.f - fx) + (c[F + 1] - c[1]) * fx;
          const float gk = gq[lvl * NN + k] * inv;
          gcx = fmaf(gk, dpx, gcx);
          gcy = fmaf(gk, dpy, gcy);
        }
      }
    }
    if (grad_coords != nullptr) {
      gcx = warp_sum(gcx);
      gcy = warp_sum(gcy);
      if (lane == 0) {
        grad_coords[((int64_t)b * 2 + 0) * Q + q] = gcx;
        grad_coords[((int64_t)b * 2 + 1) * Q + q] = gcy;
      }
    }
  }
}

__global__ void __launch_bounds__(256)
avg_pool2x2_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t P, int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const int64_t total = P * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int64_t p = i / ((int64_t)Wo * Ho);
    const float* s = in + (p * H + 2 * y) * W + 2 * x;
    // ATen avg_pool2d accumulates row-major over the window, then divides by the pool size
    float acc = __fadd_rn(__fadd_rn(__fadd_rn(__ldg(s), __ldg(s + 1)), __ldg(s + W)), __ldg(s + W + 1));
    out[i] = __fdiv_rn(acc, 4.f);
  }
}

}  // namespace mrfa

using namespace mrfa;

template <typename T, bool TILED>
static int launch_lookup_fwd(const void* l0, const void* l1, const float* coords, float* out, int B, int Q, int H,
                             int W, int64_t mbs, int64_t ro, int radius, int ocl, cudaStream_t st) {
  dim3 g((unsigned)cdiv64(Q, kQPB), (unsigned)B);
  const T* a = static_cast<const T*>(l0);
  const T* b = static_cast<const T*>(l1);
  switch (radius) {
    case 1: corr_lookup_fwd_kernel<T, 1, TILED><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, Q, H, W, mbs, ro, ocl); break;
    case 2: corr_lookup_fwd_kernel<T, 2, TILED><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, Q, H, W, mbs, ro, ocl); break;
    case 3: corr_lookup_fwd_kernel<T, 3, TILED><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, Q, H, W, mbs, ro, ocl); break;
    case 4: corr_lookup_fwd_kernel<T, 4, TILED><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, Q, H, W, mbs, ro, ocl); break;
    default: return MRFA_E_SHAPE;
  }
  return MRFA_LAUNCH_RESULT();
}

// bf16 tiled maps, radius <= 3: the vectorised tile walk
static int launch_lookup_fwd_tiled(const void* l0, const void* l1, const float* coords, float* out, int B, int Q, int H,
                                   int W, int64_t mbs, int64_t ro, int radius, int ocl, cudaStream_t st) {
  // persistent blocks: each walks groups of 32 queries with the next group's loads in flight
  const int64_t groups = (int64_t)B * cdiv64(Q, kQPB);
  if (groups >= ((int64_t)1 << 31) - 4096) return MRFA_E_SHAPE;   // 32-bit group counters (+ one grid stride) in the kernel
  const int64_t cap = (int64_t)148 * kTiledBlocksPerSM;
  const unsigned g = (unsigned)(groups < cap ? groups : cap);
  const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(l0);
  const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(l1);
  switch (radius) {
    case 1: corr_lookup_fwd_tiled_kernel<1><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, B, Q, H, W, mbs, ro, ocl); break;
    case 2: corr_lookup_fwd_tiled_kernel<2><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, B, Q, H, W, mbs, ro, ocl); break;
    case 3: corr_lookup_fwd_tiled_kernel<3><<<g, kLookupThreads, 0, st>>>(a, b, coords, out, B, Q, H, W, mbs, ro, ocl); break;
    default: return launch_lookup_fwd<__nv_bfloat16, true>(l0, l1, coords, out, B, Q, H, W, mbs, ro, radius, ocl, st);
  }
  return MRFA_LAUNCH_RESULT();
}

template <typename T, bool TILED>
static int launch_lookup_bwd(const float* go, const void* l0, const void* l1, const float* coords, float* g0,
                             float* g1, float* gc, int B, int Q, int H, int W, int64_t mbs, int64_t ro, int radius,
                             cudaStream_t st) {
  dim3 g((unsigned)cdiv64(Q, kQPB), (unsigned)B);
  const T* a = static_cast<const T*>(l0);
  const T* b = static_cast<const T*>(l1);
  switch (radius) {
    case 1: corr_lookup_bwd_kernel<T, 1, TILED><<<g, kLookupThreads, 0, st>>>(go, a, b, coords, g0, g1, gc, Q, H, W, mbs, ro); break;
    case 2: corr_lookup_bwd_kernel<T, 2, TILED><<<g, kLookupThreads, 0, st>>>(go, a, b, coords, g0, g1, gc, Q, H, W, mbs, ro); break;
    case 3: corr_lookup_bwd_kernel<T, 3, TILED><<<g, kLookupThreads, 0, st>>>(go, a, b, coords, g0, g1, gc, Q, H, W, mbs, ro); break;
    case 4: corr_lookup_bwd_kernel<T, 4, TILED><<<g, kLookupThreads, 0, st>>>(go, a, b, coords, g0, g1, gc, Q, H, W, mbs, ro); break;
    default: return MRFA_E_SHAPE;
  }
  return MRFA_LAUNCH_RESULT();
}

static bool tiled_ok(int elem_bf16, int H, int W) {      // the tiled layout exists for bf16 maps with whole super-tiles only
  return elem_bf16 && H % 8 == 0 && W % 16 == 0;
}

extern "C" int mrfa_corr_lookup_fwd(const void* level0, const void* level1, int elem_bf16, const float* coords,
                                    float* out, int B, int Q, int H, int W, int64_t map_batch_stride,
                                    int64_t row_offset, int radius, int map_layout, int out_channels_last,
                                    mrfa_stream_t stream) {
  MRFA_CHECK_ARG(level0 && level1 && coords && out);
  MRFA_CHECK_ARG(B >= 0 && Q > 0 && H >= 2 && W >= 2 && map_batch_stride >= 0 && row_offset >= 0);
  MRFA_CHECK_ARG(map_layout == MRFA_MAP_ROWMAJOR || map_layout == MRFA_MAP_TILED);
  MRFA_CHECK_SHAPE(radius >= 1 && radius <= kMaxR && B <= 65535);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (map_layout == MRFA_MAP_TILED) {
    MRFA_CHECK_SHAPE(tiled_ok(elem_bf16, H, W));
    if (((reinterpret_cast<uintptr_t>(level0) | reinterpret_cast<uintptr_t>(level1)) & 15) != 0) return MRFA_E_ALIGN;
    return launch_lookup_fwd_tiled(level0, level1, coords, out, B, Q, H, W, map_batch_stride, row_offset, radius, out_channels_last, st);
  }
  if (elem_bf16)
    return launch_lookup_fwd<__nv_bfloat16, false>(level0, level1, coords, out, B, Q, H, W, map_batch_stride, row_offset, radius, out_channels_last, st);
  return launch_lookup_fwd<float, false>(level0, level1, coords, out, B, Q, H, W, map_batch_stride, row_offset, radius, out_channels_last, st);
}

extern "C" int mrfa_corr_lookup_bwd(const float* grad_out, const void* level0, const void* level1, int elem_bf16,
                                    const float* coords, float* grad_level0, float* grad_level1, float* grad_coords,
                                    int B, int Q, int H, int W, int64_t map_batch_stride, int64_t row_offset,
                                    int radius, int map_layout, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_out && level0 && level1 && coords);
  MRFA_CHECK_ARG((grad_level0 == nullptr) == (grad_level1 == nullptr));
  MRFA_CHECK_ARG(grad_level0 || grad_coords);
  MRFA_CHECK_ARG(B >= 0 && Q > 0 && H >= 2 && W >= 2 && map_batch_stride >= 0 && row_offset >= 0);
  MRFA_CHECK_ARG(map_layout == MRFA_MAP_ROWMAJOR || map_layout == MRFA_MAP_TILED);
  MRFA_CHECK_SHAPE(radius >= 1 && radius <= kMaxR && B <= 65535);
  if (B == 0) return 0;
  cudaStream_t st = as_stream(stream);
  if (map_layout == MRFA_MAP_TILED) {
    MRFA_CHECK_SHAPE(tiled_ok(elem_bf16, H, W));
    return launch_lookup_bwd<__nv_bfloat16, true>(grad_out, level0, level1, coords, grad_level0, grad_level1, grad_coords, B, Q, H, W, map_batch_stride, row_offset, radius, st);
  }
  if (elem_bf16)
    return launch_lookup_bwd<__nv_bfloat16, false>(grad_out, level0, level1, coords, grad_level0, grad_level1, grad_coords, B, Q, H, W, map_batch_stride, row_offset, radius, st);
  return launch_lookup_bwd<float, false>(grad_out, level0, level1, coords, grad_level0, grad_level1, grad_coords, B, Q, H, W, map_batch_stride, row_offset, radius, st);
}

extern "C" int mrfa_avg_pool2x2(const float* in, float* out, int64_t P, int H, int W, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && out && P >= 0 && H >= 2 && W >= 2);
  if (P == 0) return 0;
  int64_t blocks = cdiv64(P * (H / 2) * (W / 2), 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  avg_pool2x2_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(in, out, P, H, W);
  return MRFA_LAUNCH_RESULT();
}
