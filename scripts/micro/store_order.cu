// Micro-benchmark 3: does the ORDER in which the CTAs walk the correlation volume change the achieved HBM write rate?
// Same bytes, same TMA box shapes (32 rows x 128 B level 0, 32 rows x 32 B level 1, two staging half-buffers per warp) as the
// epilogue of corr_volume_tma_kernel; optionally a fifth warp streams operand tiles out of L2 with bulk copies at the rate the
// GEMM main loop would (so the L2 -> SM fabric and the DRAM read/write turnarounds are loaded like in the real kernel).
//   order 0  A-stationary: a CTA owns a 128-row block of one pair and walks its 16 N tiles (each volume row receives 512 B
//            per tile time; the 8 KB of a row are written over the whole unit, ~50 us)
//   order 1  B-stationary: a CTA owns one 256-column N tile of one pair and walks the 43 row blocks; the 16 CTAs of a pair run
//            in lock step, so the 8 KB of a row are written within one tile time by 16 CTAs
//   order 2  like 1, but the CTAs of a pair are 16 apart in blockIdx (neighbouring SMs work on different pairs)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_order store_order.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kHalf = 4096 + 1024;      // 32 x 128 B level-0 box + 32 x 32 B level-1 box
constexpr int kRows = 5440, kPairs = 64, kNTiles = 16, kMBlocks = 43;
constexpr int kChunk = 32768;           // operand bytes per bulk load (one K block of a B tile)

__global__ void __launch_bounds__(160, 1)
store_order(const __grid_constant__ CUtensorMap m0, const __grid_constant__ CUtensorMap m1, const uint8_t* operands, int order,
            int load_chunks_per_tile) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem + 4 * 2 * kHalf + 1024;          // 3 x 32 KB operand ring
  __shared__ uint64_t full[3];
  __shared__ volatile int tiles_done;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(&full[i], 1);
    tiles_done = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int units = (order == 0) ? kPairs * kMBlocks : kPairs * kNTiles;
  const int tiles_per_unit = (order == 0) ? kNTiles : kMBlocks;

  if (warp == 4) {
    // operand stream: load_chunks_per_tile bulk copies of 32 KB per tile out of the pair's (L2-resident) operand block
    if (lane == 0 && load_chunks_per_tile > 0) {
      uint32_t n = 0;
      int tile = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        const int pair = (order == 0) ? u / kMBlocks : ((order == 1) ? u / kNTiles : u % kPairs);
        const uint8_t* src = operands + (size_t)pair * (2 << 20);
        for (int t = 0; t < tiles_per_unit; ++t, ++tile) {
          while (tile > tiles_done + 2) __nanosleep(64);
          for (int c = 0; c < load_chunks_per_tile; ++c, ++n) {
            const uint32_t s = n % 3;
            if (n >= 3) mbar_wait(&full[s], ((n / 3) - 1) & 1u);          // the copy that last used this stage has landed
            mbar_expect_tx(&full[s], kChunk);
            bulk_load(smem_u32(ring + s * kChunk), src + (size_t)((t * load_chunks_per_tile + c) % 64) * kChunk, kChunk, &full[s]);
          }
        }
      }
      for (uint32_t k = (n > 3 ? n - 3 : 0); k < n; ++k) mbar_wait(&full[k % 3], (k / 3) & 1u);
    }
    return;
  }
  const uint32_t st_base = smem_u32(smem) + warp * 2 * kHalf;
  uint32_t hcount = 0;
  int tile = 0;
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    int pair, m_fixed = 0, n_fixed = 0;
    if (order == 0) { pair = u / kMBlocks; m_fixed = u % kMBlocks; }
    else if (order == 1) { pair = u / kNTiles; n_fixed = u % kNTiles; }
    else { pair = u % kPairs; n_fixed = u / kPairs; }
    for (int t = 0; t < tiles_per_unit; ++t, ++tile) {
      const int mb = (order == 0) ? m_fixed : t, nt = (order == 0) ? t : n_fixed;
      const int row0 = mb * 128 + warp * 32;
      for (int sub = 0; sub < 4; ++sub, ++hcount) {                 // four 64-column pieces of the 256-column tile
        const uint32_t hb = st_base + (hcount & 1u) * kHalf;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) sts_v4(hb + lane * 128 + ((uint32_t)j ^ (lane & 7)) * 16, u, t, j, sub);
#pragma unroll
        for (int j = 0; j < 2; ++j) sts_v4(hb + 4096 + lane * 32 + ((uint32_t)j ^ ((lane >> 2) & 1)) * 16, u, t, j, sub);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          tma_store_3d(&m0, hb, nt * 256 + sub * 64, row0, pair);
          tma_store_3d(&m1, hb + 4096, nt * 64 + sub * 16, row0, pair);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (warp == 0 && lane == 0) tiles_done = tile + 1;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  uint8_t *v0, *v1, *flush, *ops;
  const size_t b0 = (size_t)kPairs * kRows * 8192, b1 = (size_t)kPairs * kRows * 2048;
  cudaMalloc(&v0, b0);
  cudaMalloc(&v1, b1);
  cudaMalloc(&flush, 256 << 20);
  cudaMalloc(&ops, (size_t)kPairs * (2 << 20));
  cudaMemset(ops, 1, (size_t)kPairs * (2 << 20));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fp;
  CUtensorMap m0, m1;
  {
    cuuint32_t es[3] = {1, 1, 1};
    cuuint64_t dims[3] = {4096, (cuuint64_t)kRows, (cuuint64_t)kPairs};
    cuuint64_t str[2] = {8192, (cuuint64_t)kRows * 8192};
    cuuint32_t box[3] = {64, 32, 1};
    enc(&m0, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v0, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dims1[3] = {1024, (cuuint64_t)kRows, (cuuint64_t)kPairs};
    cuuint64_t str1[2] = {2048, (cuuint64_t)kRows * 2048};
    cuuint32_t box1[3] = {16, 32, 1};
    enc(&m1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, v1, dims1, str1, box1, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int smem = 4 * 2 * kHalf + 1024 + 3 * kChunk + 1024;
  cudaFuncSetAttribute(store_order, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const char* names[3] = {"A-stationary (row block, walk N)", "B-stationary (N tile, walk rows)", "B-stationary, pairs interleaved"};
  {  // reference: plain memset of the same bytes
    float best = 1e9f;
    for (int i = 0; i < 5; ++i) {
      cudaMemsetAsync(flush, 0, 256 << 20);
      cudaEventRecord(e0);
      cudaMemsetAsync(v0, 0, b0);
      cudaMemsetAsync(v1, 0, b1);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (i >= 1 && ms < best) best = ms;
    }
    printf("%-40s loads/tile %d  %8.4f ms  %7.1f GB/s\n", "cudaMemset of both levels", 0, best, (double)(b0 + b1) / best / 1e6);
  }
  for (int chunks : {0, 2, 4}) {
    for (int order = 0; order < 3; ++order) {
      float best = 1e9f;
      for (int i = 0; i < 6; ++i) {
        cudaMemsetAsync(flush, 0, 256 << 20);
        cudaEventRecord(e0);
        store_order<<<148, 160, smem>>>(m0, m1, ops, order, chunks);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (i >= 2 && ms < best) best = ms;
      }
      // order 0 writes rows past 5440 (clipped by the tensor map), the others none: count the real bytes
      printf("%-40s loads/tile %d  %8.4f ms  %7.1f GB/s written  %s\n", names[order], chunks, best, (double)(b0 + b1) / best / 1e6,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
