// Micro-benchmark: does a store cache policy change the achievable write bandwidth of
// (a) a linear stream and (b) the correlation epilogue's strided-row pattern?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

template <int POLICY> __device__ __forceinline__ void st16(void* p, uint4 v, uint64_t pol) {
  if (POLICY == 0) asm volatile("st.global.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POLICY == 1) asm volatile("st.global.cs.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POLICY == 2) asm volatile("st.global.wt.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  if (POLICY == 3) asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
  if (POLICY == 4) asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <int POLICY>
__global__ void linear(uint4* out, int64_t n) {
  uint64_t pol = 0;
  if (POLICY == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    st16<POLICY>(out + i, make_uint4(1, 2, 3, 4), pol);
}

// CTA = 128-row block walking 16 column tiles of 512 B; warp = 32 rows; instruction = 4 rows x 128 B
template <int POLICY>
__global__ void pattern(uint8_t* out, int64_t pitch, int64_t rows) {
  uint64_t pol = 0;
  if (POLICY == 3) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  for (int64_t blk = blockIdx.x; blk < rows / 128; blk += gridDim.x)
    for (int t = 0; t < 16; ++t)
      for (int bx = 0; bx < 4; ++bx) {
        uint8_t* base = out + (blk * 128 + warp * 32) * pitch + t * 512 + bx * 128;
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + lane / 8;
          st16<POLICY>(base + r * pitch + (lane % 8) * 16, make_uint4(blk, t, bx, r), pol);
        }
      }
}

int main() {
  const int64_t rows = 64 * 5440, pitch = 8192;
  uint8_t *buf, *flush;
  cudaMalloc(&buf, rows * pitch);
  cudaMalloc(&flush, 256 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](const char* name, auto fn) {
    float best = 1e9f;
    for (int i = 0; i < 6; ++i) {
      cudaMemsetAsync(flush, 0, 256 << 20);
      cudaEventRecord(e0);
      fn();
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (i >= 2 && ms < best) best = ms;
    }
    printf("%-40s %8.4f ms  %7.1f GB/s  %s\n", name, best, (double)rows * pitch / best / 1e6, cudaGetErrorString(cudaGetLastError()));
  };
  const int64_t n = rows * pitch / 16;
  run("linear default", [&] { linear<0><<<148 * 8, 256>>>((uint4*)buf, n); });
  run("linear .cs", [&] { linear<1><<<148 * 8, 256>>>((uint4*)buf, n); });
  run("linear .wt", [&] { linear<2><<<148 * 8, 256>>>((uint4*)buf, n); });
  run("linear L2::evict_first", [&] { linear<3><<<148 * 8, 256>>>((uint4*)buf, n); });
  run("linear L1::no_allocate", [&] { linear<4><<<148 * 8, 256>>>((uint4*)buf, n); });
  run("linear default 148x32 blocks", [&] { linear<0><<<148 * 32, 256>>>((uint4*)buf, n); });
  run("pattern default", [&] { pattern<0><<<148, 128>>>(buf, pitch, rows); });
  run("pattern .cs", [&] { pattern<1><<<148, 128>>>(buf, pitch, rows); });
  run("pattern .wt", [&] { pattern<2><<<148, 128>>>(buf, pitch, rows); });
  run("pattern L2::evict_first", [&] { pattern<3><<<148, 128>>>(buf, pitch, rows); });
  run("pattern default 296 CTAs", [&] { pattern<0><<<296, 128>>>(buf, pitch, rows); });
  return 0;
}
