// All-pairs structure correlation (K1+K2+K3 of SURVEY.md; raft.py:183-185, :208, :219, :235-236,
// CorrBlock.__init__ raft.py:12-21).
//
//   corr_pack    fp32 NCHW conv outputs -> bf16 K-major GEMM operands, with the driving-side
//                average pooling (raft.py:219) applied to the *operand* (pooling commutes with
//                the contraction), so one GEMM yields every driving resolution.
//   corr_volume  persistent warp-specialised tcgen05 GEMM: A tile (128 rows x C) stationary in
//                shared memory, B tiles streamed by TMA through an mbarrier ring, fp32
//                accumulators double-buffered in TMEM, epilogue fuses the 1/sqrt(C) scale, the
//                bf16 cast and the source-side 2x2 average pool (level 1 of the pyramid) and
//                writes each row as whole 32-byte sectors.  The transposed copies of the
//                reference (raft.py:208,235-236) never exist: row i of the volume *is* the
//                h x w map that query i looks up.
//
// Roofline: the output (rows_total x hw x 1.25 bf16) makes the kernel HBM-write-bound once the
// operands sit in L2; algorithmic FLOPs = 2 * hw * hw * C per pair.
#include <cuda.h>
#include <stdlib.h>
#include "common.cuh"
#include "tcgen05.cuh"
#include "tensormap.cuh"

namespace mrfa {

// ============================================================================================
// pack
// ============================================================================================
constexpr int kPackRows = 8;        // spatial rows per tile (covers the 8x8 pooling block)
constexpr int kPackCh = 32;         // channels per tile
constexpr int kPackThreads = 256;

__device__ __forceinline__ int64_t level_row_offset(int hw, int lvl) {
  int64_t off = 0;
  for (int m = 0; m < lvl; ++m) off += hw >> (2 * m);
  return off;
}

// grid: (tiles_y * tiles_x, 2 * C/32, B); blockIdx.y < C/32 -> q_d (all levels), else k_s
__global__ void __launch_bounds__(kPackThreads)
corr_pack_kernel(const float* __restrict__ q_d, const float* __restrict__ k_s, __nv_bfloat16* __restrict__ a_op,
                 __nv_bfloat16* __restrict__ b_op, int C, int h, int w, int tile_w, int64_t rows_total, int tiled) {
  extern __shared__ float tile[];                      // [kPackCh][kPackRows * tile_w + 1]
  const int cblocks = C / kPackCh;
  const bool is_q = blockIdx.y < cblocks;
  const int c0 = (is_q ? blockIdx.y : blockIdx.y - cblocks) * kPackCh;
  const int b = blockIdx.z;
  const int tiles_x = w / tile_w;
  const int ty = blockIdx.x / tiles_x, tx = blockIdx.x - ty * tiles_x;
  const int y0 = ty * kPackRows, x0 = tx * tile_w;
  const int hw = h * w;
  const int cs = kPackRows * tile_w + 1;               // channel stride in smem (odd)
  const float* src = (is_q ? q_d : k_s) + ((int64_t)b * C + c0) * hw;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  // load: (c, y) segments of tile_w contiguous floats
  const int segs = kPackCh * kPackRows;
  for (int s = warp; s < segs; s += kPackThreads / 32) {
    const int c = s / kPackRows, y = s - c * kPackRows;
    for (int x = lane; x < tile_w; x += 32)
      tile[c * cs + y * tile_w + x] = __ldg(src + (int64_t)c * hw + (y0 + y) * w + x0 + x);
  }
  __syncthreads();

  __nv_bfloat16* dst = is_q ? a_op + (int64_t)b * rows_total * C : b_op + (int64_t)b * hw * C;
  // level 0: one pixel per warp-iteration, lane = channel -> 64-byte row pieces
  const int npix = kPackRows * tile_w;
  for (int p = warp; p < npix; p += kPackThreads / 32) {
    const int y = p / tile_w, x = p - y * tile_w;
    // source operand rows follow the map layout of the volume (tiled: the GEMM then writes tiled maps)
    const int64_t row = (!is_q && tiled) ? map_offset<true>(0, y0 + y, x0 + x, w) : (int64_t)(y0 + y) * w + x0 + x;
    dst[row * C + c0 + lane] = __float2bfloat16_rn(tile[lane * cs + p]);
  }
  if (!is_q) return;
  // pooled driving levels k = 2, 4, 8 (F.avg_pool2d: row-major window sum / k^2)
#pragma unroll
  for (int lvl = 1; lvl <= 3; ++lvl) {
    const int k = 1 << lvl;
    const int pw = tile_w / k, ph = kPackRows / k;
    const int64_t off = level_row_offset(hw, lvl);
    const int wl = w / k;
    for (int p = warp; p < pw * ph; p += kPackThreads / 32) {
      const int py = p / pw, px = p - py * pw;
      float acc = 0.f;
      for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) acc += tile[lane * cs + (py * k + dy) * tile_w + px * k + dx];
      const int64_t row = off + (int64_t)(y0 / k + py) * wl + x0 / k + px;
      dst[row * C + c0 + lane] = __float2bfloat16_rn(acc / (float)(k * k));
    }
  }
}

// channels-last inputs: (B,h,w,C) in memory is already "(h w) c" row-major, so the pack is a
// vectorised cast (float4 -> 4 x bf16) plus the pooled driving rows; one thread per (row, 4 ch).
__device__ __forceinline__ uint2 pack_bf16x4(float4 v) {
  return make_uint2(
      (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.y)) << 16),
      (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.z)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v.w)) << 16));
}

// Driving operand: thread = (pair, 8x8 pixel block, channel quad).  Every input pixel is read
// exactly once; the 2x2 / 4x4 / 8x8 means are built hierarchically in registers and written as
// the pooled rows of a_op.  Lanes run along the channel quads -> 512-byte loads, 256-byte stores.
__global__ void __launch_bounds__(256)
corr_pack_nhwc_q_kernel(const float* __restrict__ q_d, const float* __restrict__ bias, __nv_bfloat16* __restrict__ a_op, int C,
                        int h, int w, int rows_total) {
  const int cq = C / 4, bw = w / 8;
  const int nblk = (h / 8) * bw;
  const int b = blockIdx.y;
  const int hw = h * w;
  const int off1 = hw, off2 = hw + hw / 4, off3 = hw + hw / 4 + hw / 16;
  const float* qb = q_d + (int64_t)b * hw * C;
  __nv_bfloat16* ab = a_op + (int64_t)b * rows_total * C;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < (uint32_t)nblk * cq; i += gridDim.x * blockDim.x) {
    const int blk = (int)(i / (uint32_t)cq);
    const int c = (int)(i - (uint32_t)blk * cq) * 4;
    const int by = blk / bw, bx = blk - by * bw;
    // bias of the 1x1 head convolution that produced q_d (raft.py:181 kp_head), added here instead of by a separate
    // elementwise pass over the map; the pooled rows see it through the means
    const float4 bq = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 s8 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q8 = 0; q8 < 4; ++q8) {                      // 4x4 quadrants of the 8x8 block
      const int y4 = by * 8 + (q8 >> 1) * 4, x4 = bx * 8 + (q8 & 1) * 4;
      float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {                    // 2x2 blocks of the quadrant
        const int y2 = y4 + (q4 >> 1) * 2, x2 = x4 + (q4 & 1) * 2;
        float4 v[4];
#pragma unroll
        for (int p = 0; p < 4; ++p)
          v[p] = __ldg(reinterpret_cast<const float4*>(qb + ((int64_t)(y2 + (p >> 1)) * w + x2 + (p & 1)) * C + c));
#pragma unroll
        for (int p = 0; p < 4; ++p) { v[p].x += bq.x; v[p].y += bq.y; v[p].z += bq.z; v[p].w += bq.w; }
#pragma unroll
        for (int p = 0; p < 4; ++p)
          *reinterpret_cast<uint2*>(ab + ((int64_t)(y2 + (p >> 1)) * w + x2 + (p & 1)) * C + c) = pack_bf16x4(v[p]);
        float4 s2;
        s2.x = (v[0].x + v[1].x) + (v[2].x + v[3].x); s2.y = (v[0].y + v[1].y) + (v[2].y + v[3].y);
        s2.z = (v[0].z + v[1].z) + (v[2].z + v[3].z); s2.w = (v[0].w + v[1].w) + (v[2].w + v[3].w);
        *reinterpret_cast<uint2*>(ab + ((int64_t)off1 + (y2 / 2) * (w / 2) + x2 / 2) * C + c) =
            pack_bf16x4(make_float4(s2.x * 0.25f, s2.y * 0.25f, s2.z * 0.25f, s2.w * 0.25f));
        s4.x += s2.x; s4.y += s2.y; s4.z += s2.z; s4.w += s2.w;
      }
      *reinterpret_cast<uint2*>(ab + ((int64_t)off2 + (y4 / 4) * (w / 4) + x4 / 4) * C + c) =
          pack_bf16x4(make_float4(s4.x * 0.0625f, s4.y * 0.0625f, s4.z * 0.0625f, s4.w * 0.0625f));
      s8.x += s4.x; s8.y += s4.y; s8.z += s4.z; s8.w += s4.w;
    }
    *reinterpret_cast<uint2*>(ab + ((int64_t)off3 + by * bw + bx) * C + c) =
        pack_bf16x4(make_float4(s8.x * 0.015625f, s8.y * 0.015625f, s8.z * 0.015625f, s8.w * 0.015625f));
  }
}

// Source operand: a plain vectorised cast
__global__ void __launch_bounds__(256)
cast_bf16_kernel(const float4* __restrict__ x, const float* __restrict__ bias, uint2* __restrict__ y, int64_t n4, int cq) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(x + i);
    if (bias != nullptr) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + (int)(i % cq));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    y[i] = pack_bf16x4(v);
  }
}

// Source operand for the tiled map layout: the same cast with the pixel rows written in tile order
// (whole C-channel rows move, so loads and stores stay contiguous 4*C / 2*C-byte runs)
__global__ void __launch_bounds__(256)
cast_bf16_tiled_rows_kernel(const float4* __restrict__ x, const float* __restrict__ bias, uint2* __restrict__ y, int64_t pixels,
                            int hw, int w, int cq) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels * cq; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t pix = i / cq;
    const int c = (int)(i - pix * cq);
    const int64_t b = pix / hw;
    const int p = (int)(pix - b * hw);
    const int py = p / w, px = p - py * w;
    float4 v = __ldg(x + i);
    if (bias != nullptr) {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(bias) + c);
      v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
    }
    y[(b * hw + map_offset<true>(0, py, px, w)) * cq + c] = pack_bf16x4(v);
  }
}

// ============================================================================================
// tcgen05 GEMM
// ============================================================================================
constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                            // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;                      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr int kMaxKBlocks = 8;                         // C <= 512
constexpr uint32_t kATileBytes = kBlockM * kBlockK * 2;  // 16 KiB per K block

template <int kW> struct GemmCfg {
  static constexpr int kBlockN = (kW == 128) ? 256 : 128;
  static constexpr uint32_t kBStageBytes = kBlockN * kBlockK * 2;
  static constexpr int kTmemCols = 2 * kBlockN;
};

struct GemmSmemPlan {
  int a_bufs, b_stages;
  uint32_t bytes;
};

static GemmSmemPlan plan_smem(int kblocks, uint32_t b_stage_bytes) {
  const uint32_t budget = 227 * 1024 - 2048;           // alignment slack + barriers
  GemmSmemPlan p;
  const uint32_t a_one = kblocks * kATileBytes;
  p.a_bufs = (2 * a_one + 3 * b_stage_bytes <= budget) ? 2 : 1;
  int stages = (int)((budget - p.a_bufs * a_one) / b_stage_bytes);
  p.b_stages = stages > 8 ? 8 : stages;
  p.bytes = p.a_bufs * a_one + p.b_stages * b_stage_bytes + 2048;
  return p;
}

struct GemmParams {
  int B, kblocks, m_blocks, n_tiles, n_split, a_bufs, b_stages;
  int64_t rows_total;
  int N;               // hw
  float scale;
};

template <int kW>
__global__ void __launch_bounds__(kGemmThreads, 1)
corr_volume_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                   __nv_bfloat16* __restrict__ vol0, __nv_bfloat16* __restrict__ vol1, const GemmParams prm) {
  using Cfg = GemmCfg<kW>;
  constexpr int kBlockN = Cfg::kBlockN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                                   // [a_bufs][kblocks][16 KiB]
  uint8_t* smem_b = smem_a + (size_t)prm.a_bufs * prm.kblocks * kATileBytes;  // [b_stages][kBStageBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)prm.b_stages * Cfg::kBStageBytes);
  uint64_t* a_full = bars;              // [2]
  uint64_t* a_empty = bars + 2;         // [2]
  uint64_t* t_full = bars + 4;          // [2]
  uint64_t* t_empty = bars + 6;         // [2]
  uint64_t* b_full = bars + 8;          // [b_stages]
  uint64_t* b_empty = bars + 8 + 8;     // [b_stages]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);        // one arrive per epilogue warp
    }
    for (int i = 0; i < prm.b_stages; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_unit = prm.n_tiles / prm.n_split;
  const int64_t units = (int64_t)prm.B * prm.m_blocks * prm.n_split;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int split = (int)(u % prm.n_split);
        const int m_blk = (int)((u / prm.n_split) % prm.m_blocks);
        const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
        const int abuf = it % prm.a_bufs;
        const uint32_t aphase = (uint32_t)(it / prm.a_bufs) & 1u;
        mbar_wait(&a_empty[abuf], aphase ^ 1u);
        mbar_expect_tx(&a_full[abuf], (uint32_t)prm.kblocks * kATileBytes);
        for (int kb = 0; kb < prm.kblocks; ++kb)
          tma_load_3d(smem_a + ((size_t)abuf * prm.kblocks + kb) * kATileBytes, &map_a, &a_full[abuf], kb * kBlockK,
                      m_blk * kBlockM, b);
        for (int t = 0; t < tiles_per_unit; ++t) {
          const int nt = split * tiles_per_unit + t;
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            mbar_wait(&b_empty[stage], phase ^ 1u);
            mbar_expect_tx(&b_full[stage], Cfg::kBStageBytes);
            tma_load_3d(smem_b + (size_t)stage * Cfg::kBStageBytes, &map_b, &b_full[stage], kb * kBlockK, nt * kBlockN, b);
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      uint32_t tcount = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++it) {
        const int abuf = it % prm.a_bufs;
        const uint32_t aphase = (uint32_t)(it / prm.a_bufs) & 1u;
        mbar_wait(&a_full[abuf], aphase);
        tcgen05_fence_after();
        for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
          const uint32_t acc = tcount & 1u;
          const uint32_t acc_phase = (tcount >> 1) & 1u;
          mbar_wait(&t_empty[acc], acc_phase ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            mbar_wait(&b_full[stage], phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + ((size_t)abuf * prm.kblocks + kb) * kATileBytes);
            const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * Cfg::kBStageBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t da = umma_desc_sw128(a_addr + k * kUmmaK * 2);
              const uint64_t db = umma_desc_sw128(b_addr + k * kUmmaK * 2);
              tcgen05_mma_bf16(tmem_d, da, db, idesc, (uint32_t)((kb | k) != 0));
            }
            tcgen05_commit(&b_empty[stage]);            // frees the B stage when these MMAs retire
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
          tcgen05_commit(&t_full[acc]);                 // accumulator ready for the epilogue
        }
        tcgen05_commit(&a_empty[abuf]);                 // A buffer reusable
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> global =================
    const int ew = warp - 4;                            // == warp % 4: TMEM lane quarter
    const float scale = prm.scale;
    const float scale4 = prm.scale * 0.25f;
    const int N = prm.N, N4 = prm.N / 4;
    uint32_t tcount = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
      const int split = (int)(u % prm.n_split);
      const int m_blk = (int)((u / prm.n_split) % prm.m_blocks);
      const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
      const int64_t row = (int64_t)m_blk * kBlockM + ew * 32 + lane;
      const bool row_ok = row < prm.rows_total;
      __nv_bfloat16* out0 = vol0 + ((int64_t)b * prm.rows_total + row) * N;
      __nv_bfloat16* out1 = vol1 + ((int64_t)b * prm.rows_total + row) * N4;
      for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
        const int nt = split * tiles_per_unit + t;
        const uint32_t acc = tcount & 1u;
        const uint32_t acc_phase = (tcount >> 1) & 1u;
        mbar_wait(&t_full[acc], acc_phase);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * kBlockN;
        __nv_bfloat16* o0 = out0 + (int64_t)nt * kBlockN;
        __nv_bfloat16* o1 = out1 + (int64_t)nt * (kBlockN / 4);
        if constexpr (kW >= 32) {
          // source rows p (cols [g*2W, g*2W+W)) and p+1 (cols + W): pair 32-column chunks
          constexpr int kGroups = kBlockN / (2 * kW);
#pragma unroll 1
          for (int g = 0; g < kGroups; ++g) {
#pragma unroll 1
            for (int c = 0; c < kW; c += 32) {
              const int col0 = g * 2 * kW + c, col1 = col0 + kW;
              uint32_t v0[32], v1[32];
              tmem_ld_32x32(taddr + col0, v0);
              tmem_ld_32x32(taddr + col1, v1);
              tmem_ld_wait();
              if (row_ok) {
                uint32_t p0[16], p1[16], pl[8];
                float pooled[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float a0 = __uint_as_float(v0[2 * j]), a1 = __uint_as_float(v0[2 * j + 1]);
                  const float b0 = __uint_as_float(v1[2 * j]), b1 = __uint_as_float(v1[2 * j + 1]);
                  p0[j] = pack_bf16(a0 * scale, a1 * scale);
                  p1[j] = pack_bf16(b0 * scale, b1 * scale);
                  pooled[j] = ((a0 + a1) + (b0 + b1)) * scale4;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) pl[j] = pack_bf16(pooled[2 * j], pooled[2 * j + 1]);
                const uint32_t (&q0)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p0[0]);
                const uint32_t (&q1)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p0[8]);
                const uint32_t (&q2)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p1[0]);
                const uint32_t (&q3)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p1[8]);
                st_global_v8(o0 + col0, q0);
                st_global_v8(o0 + col0 + 16, q1);
                st_global_v8(o0 + col1, q2);
                st_global_v8(o0 + col1 + 16, q3);
                st_global_v8(o1 + g * (kW / 2) + c / 2, pl);
              }
            }
          }
        } else {
          // W in {8, 16}: a 32-column chunk holds 32/(2W) complete row pairs
          constexpr int kGroups = 32 / (2 * kW);
          constexpr int kHalf = kW / 2;
#pragma unroll 1
          for (int c = 0; c < kBlockN; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c, v);
            tmem_ld_wait();
            if (row_ok) {
              uint32_t p[16], pl[4];
#pragma unroll
              for (int j = 0; j < 16; ++j)
                p[j] = pack_bf16(__uint_as_float(v[2 * j]) * scale, __uint_as_float(v[2 * j + 1]) * scale);
              float pooled[8];
#pragma unroll
              for (int g = 0; g < kGroups; ++g)
#pragma unroll
                for (int j = 0; j < kHalf; ++j) {
                  const int i0 = g * 2 * kW + 2 * j, i1 = i0 + kW;
                  pooled[g * kHalf + j] = ((__uint_as_float(v[i0]) + __uint_as_float(v[i0 + 1])) +
                                           (__uint_as_float(v[i1]) + __uint_as_float(v[i1 + 1]))) * scale4;
                }
#pragma unroll
              for (int j = 0; j < 4; ++j) pl[j] = pack_bf16(pooled[2 * j], pooled[2 * j + 1]);
              const uint32_t (&q0)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p[0]);
              const uint32_t (&q1)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&p[8]);
              st_global_v8(o0 + c, q0);
              st_global_v8(o0 + c + 16, q1);
              st_global_v4(o1 + c / 4, pl[0], pl[1], pl[2], pl[3]);
            }
          }
        }
        // all of this warp's TMEM reads of the accumulator are complete
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_empty[acc]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}


// ============================================================================================
// v2: 256-row units (two M=128 accumulators share every B tile), A ring released per K block,
//     epilogue staged through swizzled shared memory and written by TMA bulk tensor stores.
//     Used for w in {64, 128}: the production shapes (256x256 and 512x512 frames).
// ============================================================================================
__device__ __forceinline__ void tma_load_3d_mcast(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr uint32_t kStageWarpBytes = 4096 + 4096 + 2048;   // two 32x64 bf16 boxes + one 32x32 box

// One 128-column step of the tiled epilogue (MRFA_MAP_TILED, include/mrfa_b200.h).  The B-operand rows were packed in
// tile order, so 128 consecutive accumulator columns are one level-0 super-tile of 16 x 8 source pixels: four 32-column
// tiles (ty, tx) of 4 rows x 8 columns.  Thread = TMEM lane = volume row.  Per `sub` (= ty) the thread reads tiles (ty,0)
// and (ty,1) -- 64 contiguous columns -- scales, packs to bf16 into the 64-column staging box `sub`, and 2x2-pools them
// into rows 2*ty, 2*ty+1 of the level-1 tile (16 contiguous level-1 values); the pooled sum keeps the order of
// F.avg_pool2d ((a + b) + (c + d), then * 1/4).
template <bool HYBRID>
__device__ __forceinline__ void stage_supertile(uint32_t taddr, float scale, float scale4, uint32_t box0, uint32_t box1,
                                                uint32_t boxl, uint32_t row128, uint32_t row64, uint32_t sw128,
                                                uint32_t sw64) {
#pragma unroll
  for (int sub = 0; sub < 2; ++sub) {
    uint32_t v0[32], v1[32];
    tmem_ld_32x32(taddr + sub * 64, v0);
    tmem_ld_32x32(taddr + sub * 64 + 32, v1);
    tmem_ld_wait();
    uint32_t p0[16], p1[16], pl[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      p0[j] = pack_bf16(__uint_as_float(v0[2 * j]) * scale, __uint_as_float(v0[2 * j + 1]) * scale);
      p1[j] = HYBRID ? pack_bf16_fma(__uint_as_float(v1[2 * j]) * scale, __uint_as_float(v1[2 * j + 1]) * scale)
                     : pack_bf16(__uint_as_float(v1[2 * j]) * scale, __uint_as_float(v1[2 * j + 1]) * scale);
    }
    // tile element (row r, column c) = v[r * 8 + c]; pooled (r2, c2) = rows 2*r2, 2*r2+1 x columns 2*c2, 2*c2+1
    float a[8], b[8];
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
      for (int c2 = 0; c2 < 4; ++c2) {
        const int i0 = r2 * 16 + 2 * c2, i1 = i0 + 8;
        a[r2 * 4 + c2] = ((__uint_as_float(v0[i0]) + __uint_as_float(v0[i0 + 1])) +
                          (__uint_as_float(v0[i1]) + __uint_as_float(v0[i1 + 1]))) * scale4;
        b[r2 * 4 + c2] = ((__uint_as_float(v1[i0]) + __uint_as_float(v1[i0 + 1])) +
                          (__uint_as_float(v1[i1]) + __uint_as_float(v1[i1 + 1]))) * scale4;
      }
    // level-1 tile rows 2*sub + r2: [a(r2, 0..3), b(r2, 0..3)]
#pragma unroll
    for (int r2 = 0; r2 < 2; ++r2) {
      pl[r2 * 4 + 0] = pack_bf16(a[r2 * 4 + 0], a[r2 * 4 + 1]);
      pl[r2 * 4 + 1] = pack_bf16(a[r2 * 4 + 2], a[r2 * 4 + 3]);
      pl[r2 * 4 + 2] = pack_bf16(b[r2 * 4 + 0], b[r2 * 4 + 1]);
      pl[r2 * 4 + 3] = pack_bf16(b[r2 * 4 + 2], b[r2 * 4 + 3]);
    }
    const uint32_t box = sub ? box1 : box0;
    // 16-byte chunk j of a 128-byte row lands at chunk (j ^ (row & 7))  [SWIZZLE_128B]
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      st_shared_v4(box + row128 + (((uint32_t)j) ^ sw128) * 16, p0[4 * j], p0[4 * j + 1], p0[4 * j + 2], p0[4 * j + 3]);
      st_shared_v4(box + row128 + (((uint32_t)(4 + j)) ^ sw128) * 16, p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
    }
    // 64-byte rows: chunk j lands at chunk (j ^ ((row >> 1) & 3))            [SWIZZLE_64B]
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t ch = (uint32_t)(sub * 2 + j) ^ sw64;
      st_shared_v4(boxl + row64 + ch * 16, pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
    }
  }
}

// Half a super-tile (64 accumulator columns = tiles (ty,0), (ty,1)) into ONE staging half-buffer: a 32 x 64 level-0 box
// (SWIZZLE_128B) and a 32 x 16 level-1 box (32-byte rows, SWIZZLE_32B).  Two half-buffers per warp alternate, so the TMA
// engine drains one while the warp fills the other (store mode 2).
template <bool HYBRID>
__device__ __forceinline__ void stage_half(uint32_t taddr, float scale, float scale4, uint32_t box, uint32_t boxl,
                                           uint32_t row128, uint32_t row32, uint32_t sw128, uint32_t sw32) {
  uint32_t v0[32], v1[32];
  tmem_ld_32x32(taddr, v0);
  tmem_ld_32x32(taddr + 32, v1);
  tmem_ld_wait();
  uint32_t p0[16], p1[16], pl[8];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    p0[j] = pack_bf16(__uint_as_float(v0[2 * j]) * scale, __uint_as_float(v0[2 * j + 1]) * scale);
    p1[j] = HYBRID ? pack_bf16_fma(__uint_as_float(v1[2 * j]) * scale, __uint_as_float(v1[2 * j + 1]) * scale)
                   : pack_bf16(__uint_as_float(v1[2 * j]) * scale, __uint_as_float(v1[2 * j + 1]) * scale);
  }
  float a[8], b[8];
#pragma unroll
  for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
    for (int c2 = 0; c2 < 4; ++c2) {
      const int i0 = r2 * 16 + 2 * c2, i1 = i0 + 8;
      a[r2 * 4 + c2] = ((__uint_as_float(v0[i0]) + __uint_as_float(v0[i0 + 1])) +
                        (__uint_as_float(v0[i1]) + __uint_as_float(v0[i1 + 1]))) * scale4;
      b[r2 * 4 + c2] = ((__uint_as_float(v1[i0]) + __uint_as_float(v1[i0 + 1])) +
                        (__uint_as_float(v1[i1]) + __uint_as_float(v1[i1 + 1]))) * scale4;
    }
#pragma unroll
  for (int r2 = 0; r2 < 2; ++r2) {
    pl[r2 * 4 + 0] = pack_bf16(a[r2 * 4 + 0], a[r2 * 4 + 1]);
    pl[r2 * 4 + 1] = pack_bf16(a[r2 * 4 + 2], a[r2 * 4 + 3]);
    pl[r2 * 4 + 2] = pack_bf16(b[r2 * 4 + 0], b[r2 * 4 + 1]);
    pl[r2 * 4 + 3] = pack_bf16(b[r2 * 4 + 2], b[r2 * 4 + 3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    st_shared_v4(box + row128 + (((uint32_t)j) ^ sw128) * 16, p0[4 * j], p0[4 * j + 1], p0[4 * j + 2], p0[4 * j + 3]);
    st_shared_v4(box + row128 + (((uint32_t)(4 + j)) ^ sw128) * 16, p1[4 * j], p1[4 * j + 1], p1[4 * j + 2], p1[4 * j + 3]);
  }
  // 32-byte rows: 16-byte chunk j lands at chunk (j ^ ((row >> 2) & 1))          [SWIZZLE_32B]
#pragma unroll
  for (int j = 0; j < 2; ++j)
    st_shared_v4(boxl + row32 + (((uint32_t)j) ^ sw32) * 16, pl[4 * j], pl[4 * j + 1], pl[4 * j + 2], pl[4 * j + 3]);
}

__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int kW, int kBlockN_, int kMTiles> struct Gemm2Cfg {
  static_assert(kW == 64 || kW == 128, "TMA-store epilogue is specialised for w = 64 / 128 (tiled map layout)");
  static_assert(kBlockN_ % 128 == 0 && kBlockN_ <= 256, "a tile holds whole 16 x 8-pixel super-tiles");
  static constexpr int kBlockN = kBlockN_;
  static constexpr int kSuper = kBlockN / 128;                            // super-tiles (epilogue steps) per tile
  static constexpr int kUnitRows = kBlockM * kMTiles;
  static constexpr uint32_t kBStageBytes = kBlockN * kBlockK * 2;
  static constexpr uint32_t kAKBytes = kUnitRows * kBlockK * 2;          // one K block of the A unit
  static constexpr int kTmemCols = 2 * kMTiles * kBlockN;               // double-buffered accumulators
  static_assert(kTmemCols <= 512, "TMEM has 512 columns");
};

struct Gemm2Params {
  int B, kblocks, m_blocks, n_tiles, n_split, b_stages;
  float scale;
  int store_mode;            // 0 = TMA bulk tensor stores from one staging buffer per warp, 1 = coalesced LSU stores from the
                             // staged boxes, 2 (default) = TMA stores from two alternating half-buffers per warp
  int N;                     // hw
  int64_t rows_total;
  __nv_bfloat16 *vol0, *vol1;
  int cvt_mode;              // 1 (default) = half of the bf16 conversions on the FMA / ALU pipes (pack_bf16_fma), 0 = all on XU
  int debug;   // MRFA_CORR_DEBUG bit mask (timing decomposition only; breaks results): 1 = no TMA stores,
               // 2 = epilogue skips TMEM loads / math / staging, 4 = no wait on staging reuse
};

template <int kW, int kBlockN_, int kMTiles, int kCluster>
__global__ void __launch_bounds__(kGemmThreads, 1)
corr_volume_tma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1,
                       const __grid_constant__ CUtensorMap map_v1h, const Gemm2Params prm) {
  using Cfg = Gemm2Cfg<kW, kBlockN_, kMTiles>;
  constexpr int kBlockN = Cfg::kBlockN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                                 // [kblocks][kUnitRows x 128 B]
  uint8_t* smem_b = smem_a + (size_t)prm.kblocks * Cfg::kAKBytes;         // [b_stages][kBStageBytes]
  uint8_t* smem_st = smem_b + (size_t)prm.b_stages * Cfg::kBStageBytes;   // [4 warps][kStageWarpBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_st + 4 * kStageWarpBytes);
  uint64_t* a_full = bars;               // [8] per K block
  uint64_t* a_empty = bars + 8;          // [8]
  uint64_t* b_full = bars + 16;          // [8] per stage
  uint64_t* b_empty = bars + 24;         // [8]
  uint64_t* t_full = bars + 32;          // [2]
  uint64_t* t_empty = bars + 34;         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  // kCluster == 2: the CTA pair works on adjacent 128-row blocks of the same pair and N tiles in
  // lock step; each CTA loads half of every B tile and multicasts it into both shared memories,
  // so B crosses the L2 -> SM fabric once per 256 rows.
  const uint32_t crank = (kCluster > 1) ? cluster_ctarank() : 0u;
  const int64_t cid = blockIdx.x / kCluster, ncl = gridDim.x / kCluster;
  constexpr uint16_t kMask = (uint16_t)((1u << kCluster) - 1u);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v1h) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], kCluster);   // released by the MMA warp of every CTA that received the tile
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)Cfg::kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // peer barriers are initialised before any multicast can land
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_unit = prm.n_tiles / prm.n_split;
  // a unit = (pair, block of kCluster * kUnitRows rows, N split); prm.m_blocks counts those blocks
  const int64_t units = (int64_t)prm.B * prm.m_blocks * prm.n_split;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int64_t u = cid; u < units; u += ncl, ++it) {
        const int split = (int)(u % prm.n_split);
        const int m_blk = (int)((u / prm.n_split) % prm.m_blocks) * kCluster + (int)crank;
        const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
        const uint32_t aphase = it & 1u;
        for (int kb = 0; kb < prm.kblocks; ++kb) {        // K block kb was released by the previous unit's last tile
          mbar_wait(&a_empty[kb], aphase ^ 1u);
          mbar_expect_tx(&a_full[kb], Cfg::kAKBytes);
          tma_load_3d(smem_a + (size_t)kb * Cfg::kAKBytes, &map_a, &a_full[kb], kb * kBlockK, m_blk * Cfg::kUnitRows, b);
        }
        for (int t = 0; t < tiles_per_unit; ++t) {
          const int nt = split * tiles_per_unit + t;
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            mbar_wait(&b_empty[stage], phase ^ 1u);
            mbar_expect_tx(&b_full[stage], Cfg::kBStageBytes);
            if (kCluster == 1) {
              tma_load_3d(smem_b + (size_t)stage * Cfg::kBStageBytes, &map_b, &b_full[stage], kb * kBlockK, nt * kBlockN, b);
            } else {
              constexpr int kSlice = kBlockN / kCluster;          // B rows this CTA fetches for the cluster
              tma_load_3d_mcast(smem_b + (size_t)stage * Cfg::kBStageBytes + (size_t)crank * kSlice * kBlockK * 2, &map_b,
                                &b_full[stage], kb * kBlockK, nt * kBlockN + (int)crank * kSlice, b, kMask);
            }
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, kBlockN);
      int stage = 0;
      uint32_t phase = 0, it = 0, tcount = 0;
      for (int64_t u = cid; u < units; u += ncl, ++it) {
        const uint32_t aphase = it & 1u;
        for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
          const uint32_t acc = tcount & 1u;
          mbar_wait(&t_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            if (t == 0) mbar_wait(&a_full[kb], aphase);
            mbar_wait(&b_full[stage], phase);
            tcgen05_fence_after();
            const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * Cfg::kBStageBytes);
#pragma unroll
            for (int half = 0; half < kMTiles; ++half) {
              const uint32_t a_addr = smem_u32(smem_a + (size_t)kb * Cfg::kAKBytes + (size_t)half * kATileBytes);
              const uint32_t tmem_d = tmem_base + (acc * kMTiles + half) * kBlockN;
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                tcgen05_mma_bf16(tmem_d, umma_desc_sw128(a_addr + k * kUmmaK * 2), umma_desc_sw128(b_addr + k * kUmmaK * 2),
                                 idesc, (uint32_t)((kb | k) != 0));
            }
            if (kCluster == 1) tcgen05_commit(&b_empty[stage]);
            else tcgen05_commit_mcast(&b_empty[stage], kMask);          // frees the stage in every CTA of the cluster
            if (t == tiles_per_unit - 1) tcgen05_commit(&a_empty[kb]);   // last use of this A K block
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
          tcgen05_commit(&t_full[acc]);
        }
      }
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> swizzled smem -> TMA store =================
    const int ew = warp - 4;
    const float scale = prm.scale, scale4 = prm.scale * 0.25f;
    const uint32_t st_base = smem_u32(smem_st + (size_t)ew * kStageWarpBytes);
    const uint32_t box0 = st_base, box1 = st_base + 4096, boxl = st_base + 8192;
    const uint32_t row128 = (uint32_t)lane * 128, row64 = (uint32_t)lane * 64;
    const uint32_t sw128 = (uint32_t)(lane & 7), sw64 = (uint32_t)((lane >> 1) & 3);
    uint32_t tcount = 0, hcount = 0;
    for (int64_t u = cid; u < units; u += ncl) {
      const int split = (int)(u % prm.n_split);
      const int m_blk = (int)((u / prm.n_split) % prm.m_blocks) * kCluster + (int)crank;
      const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
      for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
        const int nt = split * tiles_per_unit + t;
        const uint32_t acc = tcount & 1u;
        mbar_wait(&t_full[acc], (tcount >> 1) & 1u);
        tcgen05_fence_after();
#pragma unroll 1
        for (int half = 0; half < kMTiles; ++half) {
          const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (acc * kMTiles + half) * kBlockN;
          const int row0 = m_blk * Cfg::kUnitRows + half * kBlockM + ew * 32;
#pragma unroll 1
          for (int gs = 0; gs < Cfg::kSuper; ++gs) {
            if (prm.store_mode == 2) {
              // ---- (c) two half-buffers per warp: fill one while the TMA engine drains the other ----
#pragma unroll 1
              for (int sub = 0; sub < 2; ++sub, ++hcount) {
                const uint32_t hb = st_base + (hcount & 1u) * (kStageWarpBytes / 2);
                if (lane == 0) tma_store_wait_read1();           // the group that last used this half-buffer has been read
                __syncwarp();
                if (prm.cvt_mode) stage_half<true>(taddr + gs * 128 + sub * 64, scale, scale4, hb, hb + 4096, row128, (uint32_t)lane * 32, sw128, (uint32_t)((lane >> 2) & 1));
                else stage_half<false>(taddr + gs * 128 + sub * 64, scale, scale4, hb, hb + 4096, row128, (uint32_t)lane * 32, sw128, (uint32_t)((lane >> 2) & 1));
                if (half == kMTiles - 1 && gs == Cfg::kSuper - 1 && sub == 1) {
                  tcgen05_fence_before();                        // every TMEM read of this accumulator stage is done
                  __syncwarp();
                  if (lane == 0) mbar_arrive(&t_empty[acc]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                  const uint8_t* hp = smem_st + (size_t)ew * kStageWarpBytes + (size_t)(hcount & 1u) * (kStageWarpBytes / 2);
                  tma_store_3d(&map_v0, hp, nt * kBlockN + gs * 128 + sub * 64, row0, b);
                  tma_store_3d(&map_v1h, hp + 4096, nt * (kBlockN / 4) + gs * 32 + sub * 16, row0, b);
                  tma_store_commit();
                }
              }
              continue;
            }
            // the staging boxes are free once the previous bulk stores have read them
            if (prm.store_mode == 0 && lane == 0 && !(prm.debug & 4)) tma_store_wait_read();
            __syncwarp();
            if (!(prm.debug & 2)) {
              if (prm.cvt_mode) stage_supertile<true>(taddr + gs * 128, scale, scale4, box0, box1, boxl, row128, row64, sw128, sw64);
              else stage_supertile<false>(taddr + gs * 128, scale, scale4, box0, box1, boxl, row128, row64, sw128, sw64);
            }
            if (half == kMTiles - 1 && gs == Cfg::kSuper - 1) {
              // every TMEM read of this accumulator stage is done: hand it back to the MMA warp
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&t_empty[acc]);
            }
            if (prm.store_mode == 0) {
              // ---- (a) TMA bulk tensor stores of the three staged boxes ----
              fence_async_smem();
              __syncwarp();
              if (lane == 0 && !(prm.debug & 1)) {
                const int col = nt * kBlockN + gs * 128;
                const int colp = nt * (kBlockN / 4) + gs * 32;
                tma_store_3d(&map_v0, smem_st + (size_t)ew * kStageWarpBytes, col, row0, b);
                tma_store_3d(&map_v0, smem_st + (size_t)ew * kStageWarpBytes + 4096, col + 64, row0, b);
                tma_store_3d(&map_v1, smem_st + (size_t)ew * kStageWarpBytes + 8192, colp, row0, b);
                tma_store_commit();
              }
            } else {
              // ---- (b) read the staged boxes back row-contiguously and store through the LSU:
              //      every warp instruction writes 4 (or 8) whole rows = full 128-byte lines ----
              __syncwarp();
              if (!(prm.debug & 1)) {
                const int col = nt * kBlockN + gs * 128;
                const int colp = nt * (kBlockN / 4) + gs * 32;
                const int64_t rbase = (int64_t)b * prm.rows_total + row0;
                const int rows_ok = (int)min((int64_t)32, prm.rows_total - row0);     // may be <= 0
                {
                  const int ch = lane & 7, rsub = lane >> 3;                           // 8 chunks x 4 rows / instr
#pragma unroll
                  for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + rsub;
                    const uint32_t off = (uint32_t)r * 128 + (uint32_t)((ch ^ (r & 7)) * 16);
                    uint4 x0, x1;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x0.x), "=r"(x0.y), "=r"(x0.z), "=r"(x0.w) : "r"(box0 + off));
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x1.x), "=r"(x1.y), "=r"(x1.z), "=r"(x1.w) : "r"(box1 + off));
                    if (r < rows_ok) {
                      __nv_bfloat16* o = prm.vol0 + (rbase + r) * prm.N + col + ch * 8;
                      st_global_v4(o, x0.x, x0.y, x0.z, x0.w);
                      st_global_v4(o + 64, x1.x, x1.y, x1.z, x1.w);
                    }
                  }
                }
                {
                  const int ch = lane & 3, rsub = lane >> 2;                           // 4 chunks x 8 rows / instr
#pragma unroll
                  for (int it = 0; it < 4; ++it) {
                    const int r = it * 8 + rsub;
                    const uint32_t off = (uint32_t)r * 64 + (uint32_t)((ch ^ ((r >> 1) & 3)) * 16);
                    uint4 x;
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(x.x), "=r"(x.y), "=r"(x.z), "=r"(x.w) : "r"(boxl + off));
                    if (r < rows_ok) st_global_v4(prm.vol1 + (rbase + r) * (prm.N / 4) + colp + ch * 8, x.x, x.y, x.z, x.w);
                  }
                }
              }
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (kCluster > 1) cluster_sync_all();   // no CTA leaves while its peer can still multicast into it
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::kTmemCols)
                 : "memory");
  }
}



// ============================================================================================
// 2-SM variant: the CTA pair of a cluster issues one `tcgen05.mma.cta_group::2` of M = 256 per
// K step.  Each CTA keeps its own 128 A rows stationary and loads only HALF of every B tile
// (its 128 of the 256 source columns); the tensor cores read both halves across the pair, so the
// L2 -> SM operand traffic per output halves and the B ring needs half the shared memory.  The
// leader CTA (cluster rank 0) issues the MMAs; TMA loads of both CTAs signal the leader's
// barriers; tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs; both CTAs
// drain their own TMEM lanes and store their own 128 rows.
// ============================================================================================
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;          // shared::cluster address of the pair's even CTA

__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1,
                                                int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {     // arrive on cluster rank 0's copy
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(0));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <int kW>
__global__ void __launch_bounds__(kGemmThreads, 1)
corr_volume_2sm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                       const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_v1,
                       const Gemm2Params prm) {
  using Cfg = Gemm2Cfg<kW, 256, 1>;
  constexpr int kBlockN = 256;
  constexpr uint32_t kBHalfBytes = (kBlockN / 2) * kBlockK * 2;           // this CTA's half of a B stage
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                                 // [kblocks][128 x 128 B]
  uint8_t* smem_b = smem_a + (size_t)prm.kblocks * kATileBytes;           // [b_stages][128 x 128 B]
  uint8_t* smem_st = smem_b + (size_t)prm.b_stages * kBHalfBytes;         // [4 warps][kStageWarpBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_st + 4 * kStageWarpBytes);
  uint64_t* a_full = bars;               // [8]  (leader's copy is the live one)
  uint64_t* a_empty = bars + 8;          // [8]
  uint64_t* b_full = bars + 16;          // [8]  (leader)
  uint64_t* b_empty = bars + 24;         // [8]
  uint64_t* t_full = bars + 32;          // [2]
  uint64_t* t_empty = bars + 34;         // [2]  (leader; 8 arrivals: 4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 36);

  const int warp = threadIdx.x / 32;
  const int lane = threadIdx.x % 32;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int64_t cid = blockIdx.x / 2, ncl = gridDim.x / 2;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_v1) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_unit = prm.n_tiles / prm.n_split;
  const int64_t units = (int64_t)prm.B * prm.m_blocks * prm.n_split;     // m_blocks counts 256-row blocks

  if (warp == 0) {
    if (lane == 0) {                                   // ===== TMA producer (both CTAs) =====
      int stage = 0;
      uint32_t phase = 0, it = 0;
      for (int64_t u = cid; u < units; u += ncl, ++it) {
        const int split = (int)(u % prm.n_split);
        const int m_blk = (int)((u / prm.n_split) % prm.m_blocks) * 2 + (int)crank;
        const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
        const uint32_t aphase = it & 1u;
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          mbar_wait(&a_empty[kb], aphase ^ 1u);
          if (leader) mbar_expect_tx(&a_full[kb], 2 * kATileBytes);
          tma_load_3d_2sm(smem_a + (size_t)kb * kATileBytes, &map_a, &a_full[kb], kb * kBlockK, m_blk * kBlockM, b);
        }
        for (int t = 0; t < tiles_per_unit; ++t) {
          const int nt = split * tiles_per_unit + t;
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            mbar_wait(&b_empty[stage], phase ^ 1u);
            if (leader) mbar_expect_tx(&b_full[stage], 2 * kBHalfBytes);
            tma_load_3d_2sm(smem_b + (size_t)stage * kBHalfBytes, &map_b, &b_full[stage], kb * kBlockK,
                            nt * kBlockN + (int)crank * (kBlockN / 2), b);
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {                         // ===== MMA issuer (leader CTA only) =====
      constexpr uint32_t idesc = umma_idesc_bf16(256, kBlockN);
      int stage = 0;
      uint32_t phase = 0, it = 0, tcount = 0;
      for (int64_t u = cid; u < units; u += ncl, ++it) {
        const uint32_t aphase = it & 1u;
        for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
          const uint32_t acc = tcount & 1u;
          mbar_wait(&t_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * kBlockN;
          for (int kb = 0; kb < prm.kblocks; ++kb) {
            if (t == 0) mbar_wait(&a_full[kb], aphase);
            mbar_wait(&b_full[stage], phase);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(smem_a + (size_t)kb * kATileBytes);
            const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * kBHalfBytes);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k)
              tcgen05_mma_bf16_2sm(tmem_d, umma_desc_sw128(a_addr + k * kUmmaK * 2), umma_desc_sw128(b_addr + k * kUmmaK * 2),
                                   idesc, (uint32_t)((kb | k) != 0));
            tcgen05_commit_2sm(&b_empty[stage]);
            if (t == tiles_per_unit - 1) tcgen05_commit_2sm(&a_empty[kb]);
            if (++stage == prm.b_stages) { stage = 0; phase ^= 1u; }
          }
          tcgen05_commit_2sm(&t_full[acc]);
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): TMEM -> registers -> swizzled smem -> TMA store =====
    const int ew = warp - 4;
    const float scale = prm.scale, scale4 = prm.scale * 0.25f;
    const uint32_t st_base = smem_u32(smem_st + (size_t)ew * kStageWarpBytes);
    const uint32_t box0 = st_base, box1 = st_base + 4096, boxl = st_base + 8192;
    const uint32_t row128 = (uint32_t)lane * 128, row64 = (uint32_t)lane * 64;
    const uint32_t sw128 = (uint32_t)(lane & 7), sw64 = (uint32_t)((lane >> 1) & 3);
    uint32_t tcount = 0;
    for (int64_t u = cid; u < units; u += ncl) {
      const int split = (int)(u % prm.n_split);
      const int m_blk = (int)((u / prm.n_split) % prm.m_blocks) * 2 + (int)crank;
      const int b = (int)(u / ((int64_t)prm.n_split * prm.m_blocks));
      const int row0 = m_blk * kBlockM + ew * 32;
      for (int t = 0; t < tiles_per_unit; ++t, ++tcount) {
        const int nt = split * tiles_per_unit + t;
        const uint32_t acc = tcount & 1u;
        mbar_wait(&t_full[acc], (tcount >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * kBlockN;
#pragma unroll 1
        for (int gs = 0; gs < Cfg::kSuper; ++gs) {
          if (lane == 0) tma_store_wait_read();
          __syncwarp();
          stage_supertile<true>(taddr + gs * 128, scale, scale4, box0, box1, boxl, row128, row64, sw128, sw64);
          if (gs == Cfg::kSuper - 1) {
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&t_empty[acc]);       // the leader's MMA warp owns the wait
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            const int col = nt * kBlockN + gs * 128;
            const int colp = nt * (kBlockN / 4) + gs * 32;
            tma_store_3d(&map_v0, smem_st + (size_t)ew * kStageWarpBytes, col, row0, b);
            tma_store_3d(&map_v0, smem_st + (size_t)ew * kStageWarpBytes + 4096, col + 64, row0, b);
            tma_store_3d(&map_v1, smem_st + (size_t)ew * kStageWarpBytes + 8192, colp, row0, b);
            tma_store_commit();
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)512) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
static int make_bf16_map(CUtensorMap* map, const void* base, int64_t cols, int64_t rows, int B, int box_cols,
                         int box_rows, CUtensorMapSwizzle swizzle) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MRFA_E_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows * cols * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MRFA_E_DRIVER;
}

static int make_operand_map(CUtensorMap* map, const void* base, int C, int64_t rows, int B, int box_rows) {
  return make_bf16_map(map, base, C, rows, B, kBlockK, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}

template <int kW, int kBlockN_, int kMTiles, int kCluster>
static int launch_corr_volume_tma(const void* a_op, const void* b_op, void* v0, void* v1, int B, int C, int h, int w,
                                  float scale, int num_sms, cudaStream_t st) {
  using Cfg = Gemm2Cfg<kW, kBlockN_, kMTiles>;
  Gemm2Params prm;
  prm.B = B;
  prm.kblocks = C / kBlockK;
  const int N = h * w;
  const int64_t rows_total = mrfa_corr_rows_total(h, w);
  prm.m_blocks = (int)cdiv64(rows_total, Cfg::kUnitRows * kCluster);
  prm.n_tiles = N / Cfg::kBlockN;
  prm.scale = scale;
  const uint32_t fixed = (uint32_t)prm.kblocks * Cfg::kAKBytes + 4 * kStageWarpBytes + 2048;
  const uint32_t budget = 227 * 1024;
  if (fixed + 2 * Cfg::kBStageBytes > budget) return MRFA_E_SHAPE;
  int stages = (int)((budget - fixed) / Cfg::kBStageBytes);
  prm.b_stages = stages > 8 ? 8 : stages;
  const uint32_t smem_bytes = fixed + prm.b_stages * Cfg::kBStageBytes;
  int n_split = 1;
  while ((int64_t)B * prm.m_blocks * n_split < 2ll * num_sms && n_split * 2 <= prm.n_tiles &&
         prm.n_tiles % (n_split * 2) == 0)
    n_split *= 2;
  {
    const char* e = getenv("MRFA_CORR_NSPLIT");
    if (e && atoi(e) > 0 && prm.n_tiles % atoi(e) == 0) n_split = atoi(e);
  }
  prm.n_split = n_split;
  {
    const char* e = getenv("MRFA_CORR_DEBUG");
    prm.debug = e ? atoi(e) : 0;
    const char* cv = getenv("MRFA_CORR_CVT");
    prm.cvt_mode = cv ? atoi(cv) : 1;
    const char* m = getenv("MRFA_CORR_STORE");
    prm.store_mode = m ? atoi(m) : 2;
  }
  prm.N = N;
  prm.rows_total = rows_total;
  prm.vol0 = static_cast<__nv_bfloat16*>(v0);
  prm.vol1 = static_cast<__nv_bfloat16*>(v1);

  CUtensorMap map_a, map_b, map_v0, map_v1;
  int rc = make_operand_map(&map_a, a_op, C, rows_total, B, Cfg::kUnitRows);
  if (rc) return rc;
  rc = make_operand_map(&map_b, b_op, C, N, B, Cfg::kBlockN / kCluster);
  if (rc) return rc;
  rc = make_bf16_map(&map_v0, v0, N, rows_total, B, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_bf16_map(&map_v1, v1, N / 4, rows_total, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  CUtensorMap map_v1h;
  rc = make_bf16_map(&map_v1h, v1, N / 4, rows_total, B, 16, 32, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc) return rc;

  auto kern = corr_volume_tma_kernel<kW, kBlockN_, kMTiles, kCluster>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t units = (int64_t)B * prm.m_blocks * n_split;
  int64_t clusters = num_sms / kCluster;
  if (units < clusters) clusters = units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * kCluster));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_v0, map_v1, map_v1h, prm);
  return (int)e;
}

template <int kW>
static int launch_corr_volume_2sm(const void* a_op, const void* b_op, void* v0, void* v1, int B, int C, int h, int w,
                                  float scale, int num_sms, cudaStream_t st) {
  constexpr int kBlockN = 256;
  Gemm2Params prm = {};
  prm.B = B;
  prm.kblocks = C / kBlockK;
  const int N = h * w;
  const int64_t rows_total = mrfa_corr_rows_total(h, w);
  prm.m_blocks = (int)cdiv64(rows_total, 2 * kBlockM);
  prm.n_tiles = N / kBlockN;
  prm.scale = scale;
  prm.N = N;
  prm.rows_total = rows_total;
  const uint32_t half = (kBlockN / 2) * kBlockK * 2;
  const uint32_t fixed = (uint32_t)prm.kblocks * kATileBytes + 4 * kStageWarpBytes + 2048;
  const uint32_t budget = 227 * 1024;
  if (fixed + 2 * half > budget) return MRFA_E_SHAPE;
  int stages = (int)((budget - fixed) / half);
  prm.b_stages = stages > 8 ? 8 : stages;
  const uint32_t smem_bytes = fixed + prm.b_stages * half;
  int n_split = 1;
  while ((int64_t)B * prm.m_blocks * n_split < 2ll * (num_sms / 2) && n_split * 2 <= prm.n_tiles &&
         prm.n_tiles % (n_split * 2) == 0)
    n_split *= 2;
  prm.n_split = n_split;
  CUtensorMap map_a, map_b, map_v0, map_v1;
  int rc = make_operand_map(&map_a, a_op, C, rows_total, B, kBlockM);
  if (rc) return rc;
  rc = make_operand_map(&map_b, b_op, C, N, B, kBlockN / 2);
  if (rc) return rc;
  rc = make_bf16_map(&map_v0, v0, N, rows_total, B, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_bf16_map(&map_v1, v1, N / 4, rows_total, B, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
  if (rc) return rc;
  auto kern = corr_volume_2sm_kernel<kW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t units = (int64_t)B * prm.m_blocks * n_split;
  int64_t clusters = num_sms / 2;
  if (units < clusters) clusters = units;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, map_a, map_b, map_v0, map_v1, prm);
  return (int)e;
}

template <int kW>
static int launch_corr_volume(const void* a_op, const void* b_op, void* v0, void* v1, int B, int C, int h, int w,
                              float scale, int num_sms, cudaStream_t st) {
  using Cfg = GemmCfg<kW>;
  GemmParams prm;
  prm.B = B;
  prm.kblocks = C / kBlockK;
  prm.N = h * w;
  prm.rows_total = mrfa_corr_rows_total(h, w);
  prm.m_blocks = (int)cdiv64(prm.rows_total, kBlockM);
  prm.n_tiles = prm.N / Cfg::kBlockN;
  prm.scale = scale;
  const GemmSmemPlan plan = plan_smem(prm.kblocks, Cfg::kBStageBytes);
  prm.a_bufs = plan.a_bufs;
  prm.b_stages = plan.b_stages;
  // split the N range of a row block until there are enough units to fill the machine twice
  int n_split = 1;
  while ((int64_t)B * prm.m_blocks * n_split < 2ll * num_sms && n_split * 2 <= prm.n_tiles &&
         prm.n_tiles % (n_split * 2) == 0)
    n_split *= 2;
  prm.n_split = n_split;

  CUtensorMap map_a, map_b;
  int rc = make_operand_map(&map_a, a_op, C, prm.rows_total, B, kBlockM);
  if (rc) return rc;
  rc = make_operand_map(&map_b, b_op, C, prm.N, B, Cfg::kBlockN);
  if (rc) return rc;

  auto kern = corr_volume_kernel<kW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.bytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t units = (int64_t)B * prm.m_blocks * n_split;
  const unsigned grid = (unsigned)(units < num_sms ? units : num_sms);
  kern<<<grid, kGemmThreads, plan.bytes, st>>>(map_a, map_b, static_cast<__nv_bfloat16*>(v0),
                                               static_cast<__nv_bfloat16*>(v1), prm);
  return MRFA_LAUNCH_RESULT();
}

}  // namespace mrfa

using namespace mrfa;

extern "C" int64_t mrfa_corr_rows_total(int h, int w) {
  const int64_t hw = (int64_t)h * w;
  return hw + hw / 4 + hw / 16 + hw / 64;
}

extern "C" int64_t mrfa_corr_row_offset(int h, int w, int pool_log2) {
  const int64_t hw = (int64_t)h * w;
  int64_t off = 0;
  for (int m = 0; m < pool_log2; ++m) off += hw >> (2 * m);
  return off;
}

extern "C" int mrfa_corr_map_layout(int h, int w) {
  // the TMA-store GEMM kernels (w = 64 / 128: 256x256 and 512x512 frames) write tiled maps; the small-shape kernel keeps row-major
  return ((w == 64 || w == 128) && h % 8 == 0) ? MRFA_MAP_TILED : MRFA_MAP_ROWMAJOR;
}

extern "C" int64_t mrfa_corr_map_offset(int map_layout, int level, int y, int x, int W_level) {
  return map_layout == MRFA_MAP_TILED ? map_offset<true>(level, y, x, W_level) : map_offset<false>(level, y, x, W_level);
}

static int corr_pack_impl(const float* q_d, const float* q_bias, const float* k_s, const float* k_bias, void* a_op, void* b_op,
                          int B, int C, int h, int w, int channels_last, mrfa_stream_t stream);

extern "C" int mrfa_corr_pack(const float* q_d, const float* k_s, void* a_op, void* b_op, int B, int C, int h, int w,
                              int channels_last, mrfa_stream_t stream) {
  return corr_pack_impl(q_d, nullptr, k_s, nullptr, a_op, b_op, B, C, h, w, channels_last, stream);
}

extern "C" int mrfa_corr_pack_bias(const float* q_d, const float* q_bias, const float* k_s, const float* k_bias, void* a_op,
                                   void* b_op, int B, int C, int h, int w, mrfa_stream_t stream) {
  if (((reinterpret_cast<uintptr_t>(q_bias) | reinterpret_cast<uintptr_t>(k_bias)) & 15) != 0) return MRFA_E_ALIGN;
  return corr_pack_impl(q_d, q_bias, k_s, k_bias, a_op, b_op, B, C, h, w, 1, stream);
}

static int corr_pack_impl(const float* q_d, const float* q_bias, const float* k_s, const float* k_bias, void* a_op, void* b_op,
                          int B, int C, int h, int w, int channels_last, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(q_d && k_s && a_op && b_op && B >= 0 && C > 0 && h > 0 && w > 0);
  MRFA_CHECK_ARG(channels_last || (q_bias == nullptr && k_bias == nullptr));
  MRFA_CHECK_SHAPE(C % kPackCh == 0 && h % 8 == 0 && w % 8 == 0 && B <= 65535);
  MRFA_CHECK_SHAPE(w <= 32 || w % 32 == 0);
  if (B == 0) return 0;
  const int tiled = mrfa_corr_map_layout(h, w) == MRFA_MAP_TILED;
  if (channels_last) {
    if (((reinterpret_cast<uintptr_t>(q_d) | reinterpret_cast<uintptr_t>(k_s)) & 15) != 0) return MRFA_E_ALIGN;
    const int64_t rows_total = mrfa_corr_rows_total(h, w);
    const int64_t items = (int64_t)(h / 8) * (w / 8) * (C / 4);
    MRFA_CHECK_SHAPE(items < (1ll << 31) && rows_total < (1ll << 31));
    int64_t blocks = cdiv64(items, 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    corr_pack_nhwc_q_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, 0, as_stream(stream)>>>(
        q_d, q_bias, static_cast<__nv_bfloat16*>(a_op), C, h, w, (int)rows_total);
    int rc = MRFA_LAUNCH_RESULT();
    if (rc) return rc;
    const int64_t n4 = (int64_t)B * h * w * C / 4;
    int64_t cblocks = cdiv64(n4, 256);
    if (cblocks > 148 * 32) cblocks = 148 * 32;
    if (tiled)
      cast_bf16_tiled_rows_kernel<<<(unsigned)cblocks, 256, 0, as_stream(stream)>>>(
          reinterpret_cast<const float4*>(k_s), k_bias, static_cast<uint2*>(b_op), (int64_t)B * h * w, h * w, w, C / 4);
    else
      cast_bf16_kernel<<<(unsigned)cblocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(k_s), k_bias,
                                                                        static_cast<uint2*>(b_op), n4, C / 4);
    return MRFA_LAUNCH_RESULT();
  }
  const int tile_w = w < 32 ? w : 32;
  dim3 grid((unsigned)((h / kPackRows) * (w / tile_w)), (unsigned)(2 * C / kPackCh), (unsigned)B);
  const size_t smem = (size_t)kPackCh * (kPackRows * tile_w + 1) * sizeof(float);
  corr_pack_kernel<<<grid, kPackThreads, smem, as_stream(stream)>>>(
      q_d, k_s, static_cast<__nv_bfloat16*>(a_op), static_cast<__nv_bfloat16*>(b_op), C, h, w, tile_w,
      mrfa_corr_rows_total(h, w), tiled);
  return MRFA_LAUNCH_RESULT();
}

// MRFA_CORR_VARIANT (A/B measurements only): 1 = 256-row units x 128-wide tiles, 4 = 2-CTA cluster with
// multicast B tiles; default 0 = 128-row units x 256-wide tiles, no cluster (all within 3 % on B200)
static int corr_variant() {
  static const int v = []() {
    const char* e = getenv("MRFA_CORR_VARIANT");
    return e ? atoi(e) : 0;
  }();
  return v;
}

extern "C" int mrfa_corr_volume(const void* a_op, const void* b_op, void* volume0, void* volume1, int B, int C, int h,
                                int w, float scale, int num_sms, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(a_op && b_op && volume0 && volume1 && B >= 0 && C > 0 && h > 0 && w > 0);
  MRFA_CHECK_SHAPE(C % kBlockK == 0 && C / kBlockK <= kMaxKBlocks && h % 8 == 0);
  if ((reinterpret_cast<uintptr_t>(a_op) | reinterpret_cast<uintptr_t>(b_op) | reinterpret_cast<uintptr_t>(volume0) |
       reinterpret_cast<uintptr_t>(volume1)) & 31)
    return MRFA_E_ALIGN;
  if (B == 0) return 0;
  if (num_sms <= 0) num_sms = 148;
  cudaStream_t st = as_stream(stream);
  const int N = h * w;
  switch (w) {
    case 8: MRFA_CHECK_SHAPE(N % 128 == 0); return launch_corr_volume<8>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
    case 16: MRFA_CHECK_SHAPE(N % 128 == 0); return launch_corr_volume<16>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
    case 32: MRFA_CHECK_SHAPE(N % 128 == 0); return launch_corr_volume<32>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
    case 64:
      MRFA_CHECK_SHAPE(N % 128 == 0);
      // 128-row units, 256-wide tiles (two source row pairs): every B tile (32 KiB per K block) feeds
      // 128x256 outputs and ~96 KiB of B stay in flight; deeper K falls back to 128-wide tiles
      if (corr_variant() == 1) return launch_corr_volume_tma<64, 128, 2, 1>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      if (corr_variant() == 5 && N % 256 == 0) return launch_corr_volume_2sm<64>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      if (corr_variant() == 4 && C <= 256 && N % 256 == 0) return launch_corr_volume_tma<64, 256, 1, 2>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      if (C <= 256 && N % 256 == 0) return launch_corr_volume_tma<64, 256, 1, 1>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      return launch_corr_volume_tma<64, 128, 1, 1>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
    case 128:
      MRFA_CHECK_SHAPE(N % 256 == 0);
      if (corr_variant() == 5) return launch_corr_volume_2sm<128>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      if (corr_variant() == 4) return launch_corr_volume_tma<128, 256, 1, 2>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
      return launch_corr_volume_tma<128, 256, 1, 1>(a_op, b_op, volume0, volume1, B, C, h, w, scale, num_sms, st);
    default: return MRFA_E_SHAPE;
  }
}
