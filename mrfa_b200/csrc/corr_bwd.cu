// Backward of the all-pairs correlation (config 5: the training step differentiates through raft.py:183-185).
//
//   volume0[b,i,j] = scale * sum_c A[b,i,c] * Bm[b,j,c]      A = packed driving operand (all pooled levels),
//   volume1        = 2x2 source pool of volume0              Bm = packed source operand (rows in map order)
//
// The lookups' backward kernels scatter-add fp32 gradients G0 (layout of volume0) and G1 (layout of volume1).  With
// G = scale * (G0 + unpool(G1) / 4):
//   dA[b,i,c] = sum_j G[b,i,j] * Bm[b,j,c]          (rows_total x C, K = hw)
//   dB[b,j,c] = sum_i G[b,i,j] * A[b,i,c]           (hw x C,         K = rows_total)
// -- the two GEMMs the reference gets from autograd through einsum.  Here:
//   corr_bwd_pack_kernel      G0/G1 fp32 -> G (bf16, K = j contiguous) and G^T (bf16, K = i contiguous) in one pass
//   transpose_bf16_kernel     A, Bm -> A^T, Bm^T (the K-major "N x K" operands of the two products)
//   corr_bwd_gemm_kernel      persistent warp-specialised tcgen05 GEMM D = X * Y^T: TMA-fed 4-stage ring for both
//                             operands, fp32 accumulators double-buffered in TMEM, fp32 epilogue
//   corr_bwd_unpack_kernel    dA rows -> d(q_d) (un-pooling the pooled driving rows, raft.py:219), dB rows -> d(k_s)
// Roofline: tensor (2 * 2 * rows_total * hw * C FLOP per pair = 22.8 GFLOP at 256x256); the pack pass is HBM-bound.
#include <cuda.h>
#include "common.cuh"
#include "tcgen05.cuh"
#include "tensormap.cuh"

namespace mrfa {

// ---------------------------------------------------------------------------------------------
// pack: G and G^T in bf16
// ---------------------------------------------------------------------------------------------
// level-1 stored position that level-0 stored position j was pooled into
template <bool TILED>
__device__ __forceinline__ int l1_of_l0(int j, int w) {
  if (TILED) {
    // j = st*128 + ty*64 + tx*32 + r*8 + c  ->  st*32 + (2*ty + r/2)*8 + 4*tx + c/2
    const int st = j >> 7, ty = (j >> 6) & 1, tx = (j >> 5) & 1, r = (j >> 3) & 3, c = j & 7;
    return st * 32 + (2 * ty + (r >> 1)) * 8 + 4 * tx + (c >> 1);
  }
  const int y = j / w, x = j - y * w;
  return (y >> 1) * (w >> 1) + (x >> 1);
}

constexpr int kTP = 32;   // transpose tile

// grid: (N/32, ceil(rows/32), B); block 32 x 8.  Each thread handles 4 rows of a 32 x 32 tile.
template <bool TILED>
__global__ void __launch_bounds__(256)
corr_bwd_pack_kernel(const float* __restrict__ g0, const float* __restrict__ g1, __nv_bfloat16* __restrict__ G,
                     __nv_bfloat16* __restrict__ GT, int rows, int N, int w, float scale, int rows_pad) {
  __shared__ float tile[kTP][kTP + 1];
  const int b = blockIdx.z;
  const int j0 = blockIdx.x * kTP, r0 = blockIdx.y * kTP;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float* g0b = g0 + (int64_t)b * rows * N;
  const float* g1b = g1 + (int64_t)b * rows * (N / 4);
  __nv_bfloat16* Gb = G + (int64_t)b * rows * N;
  __nv_bfloat16* GTb = GT + (int64_t)b * N * rows_pad;
  const int j = j0 + tx;
  const int jl1 = l1_of_l0<TILED>(j, w);
#pragma unroll
  for (int k = 0; k < kTP; k += 8) {
    const int r = r0 + ty + k;
    float v = 0.f;
    if (r < rows) {
      v = scale * fmaf(0.25f, __ldg(g1b + (int64_t)r * (N / 4) + jl1), __ldg(g0b + (int64_t)r * N + j));
      Gb[(int64_t)r * N + j] = __float2bfloat16_rn(v);
    }
    tile[ty + k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTP; k += 8) {
    const int jj = j0 + ty + k, r = r0 + tx;
    if (r < rows) GTb[(int64_t)jj * rows_pad + r] = __float2bfloat16_rn(tile[tx][ty + k]);
  }
}

// out[b, c, r] = in[b, r, c]   (rows x cols -> cols x out_pitch); grid (ceil(cols/32), ceil(rows/32), B), block 32 x 8
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int rows, int cols,
                      int out_pitch) {
  __shared__ __nv_bfloat16 tile[kTP][kTP + 2];
  const int b = blockIdx.z;
  const int c0 = blockIdx.x * kTP, r0 = blockIdx.y * kTP;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const __nv_bfloat16* ib = in + (int64_t)b * rows * cols;
  __nv_bfloat16* ob = out + (int64_t)b * cols * out_pitch;
#pragma unroll
  for (int k = 0; k < kTP; k += 8) {
    const int r = r0 + ty + k, c = c0 + tx;
    tile[ty + k][tx] = (r < rows && c < cols) ? ib[(int64_t)r * cols + c] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kTP; k += 8) {
    const int c = c0 + ty + k, r = r0 + tx;
    if (c < cols && r < rows) ob[(int64_t)c * out_pitch + r] = tile[tx][ty + k];
  }
}

// ---------------------------------------------------------------------------------------------
// D[b, m, n] = sum_k X[b, m, k] * Y[b, n, k]     X (B, M, ldx) / Y (B, Nn, ldy) bf16 K-major, D (B, M, Nn) fp32
// ---------------------------------------------------------------------------------------------
constexpr int kBM = 128, kBK = 64, kUK = 16;
constexpr int kBwdThreads = 256;            // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue
constexpr uint32_t kXTileBytes = kBM * kBK * 2;

struct BwdGemmParams {
  int B, M, Nn, block_n, n_tiles, m_blocks, kblocks, stages;
  uint32_t y_tile_bytes;
  float* D;
};

__global__ void __launch_bounds__(kBwdThreads, 1)
corr_bwd_gemm_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                     const BwdGemmParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t stage_bytes = kXTileBytes + prm.y_tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)prm.stages * stage_bytes);
  uint64_t* full = bars;               // [8]
  uint64_t* empty = bars + 8;          // [8]
  uint64_t* t_full = bars + 16;        // [2]
  uint64_t* t_empty = bars + 18;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t units = (int64_t)prm.B * prm.m_blocks * prm.n_tiles;

  if (warp == 0) {
    if (lane == 0) {                                     // ===== TMA producer =====
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x) {
        const int nt = (int)(u % prm.n_tiles);
        const int mb = (int)((u / prm.n_tiles) % prm.m_blocks);
        const int b = (int)(u / ((int64_t)prm.n_tiles * prm.m_blocks));
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_expect_tx(&full[stage], stage_bytes);
          uint8_t* sx = smem + (size_t)stage * stage_bytes;
          tma_load_3d(sx, &map_x, &full[stage], kb * kBK, mb * kBM, b);
          tma_load_3d(sx + kXTileBytes, &map_y, &full[stage], kb * kBK, nt * prm.block_n, b);
          if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {                                     // ===== MMA issuer =====
      const uint32_t idesc = umma_idesc_bf16(kBM, prm.block_n);
      int stage = 0;
      uint32_t phase = 0, tcount = 0;
      for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++tcount) {
        const uint32_t acc = tcount & 1u;
        mbar_wait(&t_empty[acc], ((tcount >> 1) & 1u) ^ 1u);
        tcgen05_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 256;
        for (int kb = 0; kb < prm.kblocks; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          const uint32_t x_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t y_addr = x_addr + kXTileBytes;
#pragma unroll
          for (int k = 0; k < kBK / kUK; ++k)
            tcgen05_mma_bf16(tmem_d, umma_desc_sw128(x_addr + k * kUK * 2), umma_desc_sw128(y_addr + k * kUK * 2), idesc,
                             (uint32_t)((kb | k) != 0));
          tcgen05_commit(&empty[stage]);
          if (++stage == prm.stages) { stage = 0; phase ^= 1u; }
        }
        tcgen05_commit(&t_full[acc]);
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> fp32 rows (thread = TMEM lane = output row; 32-byte stores) =====
    const int ew = warp - 4;
    uint32_t tcount = 0;
    for (int64_t u = blockIdx.x; u < units; u += gridDim.x, ++tcount) {
      const int nt = (int)(u % prm.n_tiles);
      const int mb = (int)((u / prm.n_tiles) % prm.m_blocks);
      const int b = (int)(u / ((int64_t)prm.n_tiles * prm.m_blocks));
      const uint32_t acc = tcount & 1u;
      mbar_wait(&t_full[acc], (tcount >> 1) & 1u);
      tcgen05_fence_after();
      const int row = mb * kBM + ew * 32 + lane;
      const bool ok = row < prm.M;
      float* drow = prm.D + ((int64_t)b * prm.M + row) * prm.Nn + (int64_t)nt * prm.block_n;
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + acc * 256;
#pragma unroll 1
      for (int c = 0; c < prm.block_n; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        if (ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t (&v8)[8] = *reinterpret_cast<const uint32_t(*)[8]>(&v[8 * q]);
            st_global_v8(drow + c + 8 * q, v8);
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&t_empty[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------
// unpack: operand gradients -> d(q_d), d(k_s) in NHWC memory (B, h, w, C)
// ---------------------------------------------------------------------------------------------
template <bool TILED>
__global__ void __launch_bounds__(256)
corr_bwd_unpack_kernel(const float* __restrict__ dA, const float* __restrict__ dB, float* __restrict__ dq,
                       float* __restrict__ dk, int C, int h, int w, int rows_total) {
  const int cq = C / 4;
  const int hw = h * w;
  const int b = blockIdx.y;
  const int off1 = hw, off2 = hw + hw / 4, off3 = hw + hw / 4 + hw / 16;
  const float* a = dA + (int64_t)b * rows_total * C;
  const float* bm = dB + (int64_t)b * hw * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (int64_t)hw * cq; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / cq);
    const int c = (int)(i - (int64_t)p * cq) * 4;
    const int y = p / w, x = p - y * w;
    // avg_pool2d backward: every pooled driving row spreads its gradient / k^2 over its k x k block
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(a + (int64_t)p * C + c));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(a + ((int64_t)off1 + (y >> 1) * (w >> 1) + (x >> 1)) * C + c));
    const float4 v2 = __ldg(reinterpret_cast<const float4*>(a + ((int64_t)off2 + (y >> 2) * (w >> 2) + (x >> 2)) * C + c));
    const float4 v3 = __ldg(reinterpret_cast<const float4*>(a + ((int64_t)off3 + (y >> 3) * (w >> 3) + (x >> 3)) * C + c));
    float4 r;
    r.x = v0.x + 0.25f * v1.x + 0.0625f * v2.x + 0.015625f * v3.x;
    r.y = v0.y + 0.25f * v1.y + 0.0625f * v2.y + 0.015625f * v3.y;
    r.z = v0.z + 0.25f * v1.z + 0.0625f * v2.z + 0.015625f * v3.z;
    r.w = v0.w + 0.25f * v1.w + 0.0625f * v2.w + 0.015625f * v3.w;
    *reinterpret_cast<float4*>(dq + ((int64_t)b * hw + p) * C + c) = r;
    const int64_t jb = map_offset<TILED>(0, y, x, w);
    *reinterpret_cast<float4*>(dk + ((int64_t)b * hw + p) * C + c) = __ldg(reinterpret_cast<const float4*>(bm + jb * C + c));
  }
}

static int make_kmajor_map(CUtensorMap* map, const void* base, int64_t k_extent, int64_t rows, int64_t pitch, int B,
                           int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return MRFA_E_DRIVER;
  cuuint64_t dims[3] = {(cuuint64_t)k_extent, (cuuint64_t)rows, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)pitch * 2, (cuuint64_t)rows * pitch * 2};
  cuuint32_t box[3] = {(cuuint32_t)kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : MRFA_E_DRIVER;
}

}  // namespace mrfa

using namespace mrfa;

extern "C" int64_t mrfa_corr_bwd_rows_pad(int h, int w) {
  return (mrfa_corr_rows_total(h, w) + 63) / 64 * 64;
}

extern "C" int mrfa_corr_bwd_pack(const float* g0, const float* g1, void* G, void* GT, int B, int h, int w, float scale,
                                  mrfa_stream_t stream) {
  MRFA_CHECK_ARG(g0 && g1 && G && GT && B >= 0 && h > 0 && w > 0);
  const int N = h * w;
  MRFA_CHECK_SHAPE(N % 32 == 0 && h % 8 == 0 && w % 8 == 0 && B <= 65535);
  if (B == 0) return 0;
  const int rows = (int)mrfa_corr_rows_total(h, w);
  const int rows_pad = (int)mrfa_corr_bwd_rows_pad(h, w);
  dim3 grid((unsigned)(N / kTP), (unsigned)cdiv64(rows, kTP), (unsigned)B), block(kTP, 8);
  if (mrfa_corr_map_layout(h, w) == MRFA_MAP_TILED)
    corr_bwd_pack_kernel<true><<<grid, block, 0, as_stream(stream)>>>(g0, g1, static_cast<__nv_bfloat16*>(G),
                                                                     static_cast<__nv_bfloat16*>(GT), rows, N, w, scale, rows_pad);
  else
    corr_bwd_pack_kernel<false><<<grid, block, 0, as_stream(stream)>>>(g0, g1, static_cast<__nv_bfloat16*>(G),
                                                                      static_cast<__nv_bfloat16*>(GT), rows, N, w, scale, rows_pad);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_transpose_bf16(const void* in, void* out, int B, int rows, int cols, int out_pitch,
                                   mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && out && B >= 0 && rows > 0 && cols > 0 && out_pitch >= rows);
  MRFA_CHECK_SHAPE(B <= 65535);
  if (B == 0) return 0;
  dim3 grid((unsigned)cdiv64(cols, kTP), (unsigned)cdiv64(rows, kTP), (unsigned)B), block(kTP, 8);
  transpose_bf16_kernel<<<grid, block, 0, as_stream(stream)>>>(static_cast<const __nv_bfloat16*>(in),
                                                               static_cast<__nv_bfloat16*>(out), rows, cols, out_pitch);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_corr_bwd_gemm(const void* X, const void* Y, float* D, int B, int M, int Nn, int K, int64_t ldx,
                                  int64_t ldy, int num_sms, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(X && Y && D && B >= 0 && M > 0 && Nn > 0 && K > 0 && ldx >= K && ldy >= K);
  MRFA_CHECK_SHAPE(Nn % 64 == 0 && (Nn <= 256 || Nn % 256 == 0) && ldx % 8 == 0 && ldy % 8 == 0);
  if ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) & 15) return MRFA_E_ALIGN;
  if (reinterpret_cast<uintptr_t>(D) & 31) return MRFA_E_ALIGN;
  if (B == 0) return 0;
  if (num_sms <= 0) num_sms = 148;
  BwdGemmParams prm;
  prm.B = B; prm.M = M; prm.Nn = Nn;
  prm.block_n = Nn <= 256 ? Nn : 256;
  prm.n_tiles = Nn / prm.block_n;
  prm.m_blocks = (int)cdiv64(M, kBM);
  prm.kblocks = (int)cdiv64(K, kBK);
  prm.y_tile_bytes = (uint32_t)prm.block_n * kBK * 2;
  const uint32_t stage_bytes = kXTileBytes + prm.y_tile_bytes;
  int stages = (int)((227 * 1024 - 2048) / stage_bytes);
  prm.stages = stages > 8 ? 8 : stages;
  prm.D = D;
  const uint32_t smem_bytes = prm.stages * stage_bytes + 2048;
  CUtensorMap map_x, map_y;
  int rc = make_kmajor_map(&map_x, X, K, M, ldx, B, kBM);
  if (rc) return rc;
  rc = make_kmajor_map(&map_y, Y, K, Nn, ldy, B, prm.block_n);
  if (rc) return rc;
  cudaError_t e = cudaFuncSetAttribute(corr_bwd_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t units = (int64_t)B * prm.m_blocks * prm.n_tiles;
  const unsigned grid = (unsigned)(units < num_sms ? units : num_sms);
  corr_bwd_gemm_kernel<<<grid, kBwdThreads, smem_bytes, as_stream(stream)>>>(map_x, map_y, prm);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_corr_bwd_unpack(const float* dA, const float* dB, float* d_q, float* d_k, int B, int C, int h, int w,
                                    mrfa_stream_t stream) {
  MRFA_CHECK_ARG(dA && dB && d_q && d_k && B >= 0 && C > 0 && h > 0 && w > 0);
  MRFA_CHECK_SHAPE(C % 4 == 0 && h % 8 == 0 && w % 8 == 0 && B <= 65535);
  if (((reinterpret_cast<uintptr_t>(dA) | reinterpret_cast<uintptr_t>(dB) | reinterpret_cast<uintptr_t>(d_q) |
        reinterpret_cast<uintptr_t>(d_k)) & 15) != 0)
    return MRFA_E_ALIGN;
  if (B == 0) return 0;
  const int64_t items = (int64_t)h * w * (C / 4);
  int64_t blocks = cdiv64(items, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  dim3 grid((unsigned)blocks, (unsigned)B);
  const int rows = (int)mrfa_corr_rows_total(h, w);
  if (mrfa_corr_map_layout(h, w) == MRFA_MAP_TILED)
    corr_bwd_unpack_kernel<true><<<grid, 256, 0, as_stream(stream)>>>(dA, dB, d_q, d_k, C, h, w, rows);
  else
    corr_bwd_unpack_kernel<false><<<grid, 256, 0, as_stream(stream)>>>(dA, dB, d_q, d_k, C, h, w, rows);
  return MRFA_LAUNCH_RESULT();
}
