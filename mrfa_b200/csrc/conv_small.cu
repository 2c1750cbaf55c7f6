// 7x7 / stride 1 / pad 3 convolutions with 2-3 input channels, fused with bias (+ folded
// BatchNorm) and ReLU: BasicMotionEncoder.convf1 (2 -> 128, raft.py:56,63, at every refinement
// level up to 256x256) and the generator's `first` SameBlock2d (3 -> 64, generator.py:13 /
// util.py:199-214).  These are callers on either side of the hot path: the
// library convolution serves them with a legacy indexed kernel (2.3 ms and 1.6 ms per batch of
// 64 at 256x256, 3-10x their output-write time).
//
// Implicit GEMM on the tcgen05 tensor cores, kind::tf32:
//   M = 128 consecutive pixels of one image row, N = Cout, K = 49 * Cin (zero-padded to a
//   multiple of 32 = one 128-byte swizzle row of fp32).
//   producers (warps 0-7)  stage the 7 x 134 x Cin input patch in shared memory, then write the
//                          im2col A tile straight into the K-major SWIZZLE_128B layout the MMA
//                          reads (bank-conflict-free: a warp store covers 8 rows x 4 words);
//   MMA       (warp 12)    one thread issues K/8 tcgen05.mma per tile into a double-buffered
//                          TMEM accumulator; weights stay resident in shared memory;
//   epilogue  (warps 8-11) TMEM -> registers -> ReLU -> swizzled staging -> TMA bulk stores, NHWC
//                          (tiles enumerate pixels in NHWC order: pixel = tile * 128 + row).  The bias rides
//                          in two spare K columns (TF32 hi + lo against a column of ones).
// HBM-bound on the output write: algorithmic bytes per pixel = 4 * (Cin + Cout).
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "tcgen05.cuh"
#include "tensormap.cuh"

namespace mrfa {

constexpr int kProducerWarps = 8;           // warps 0-7 producers, 8-11 epilogue, 12 MMA
constexpr int kProducers = kProducerWarps * 32;
constexpr int kMmaWarp = kProducerWarps + 4;
constexpr int kConvThreads = (kMmaWarp + 1) * 32;
constexpr int kTileM = 128;
constexpr int kPatchW = kTileM + 6;
constexpr uint32_t kKBlockBytes = kTileM * 128;       // one 32-wide K block of the A tile

template <int CIN, int COUT> struct ConvSmallCfg {
  static constexpr int kK = 49 * CIN;
  static constexpr int kKBlocks = (kK + 31) / 32;
  static constexpr int kKP = kKBlocks * 32;
  static constexpr uint32_t kABytes = kKBlocks * kKBlockBytes;               // per stage
  static constexpr uint32_t kBBlockBytes = COUT * 128;
  static constexpr uint32_t kBBytes = kKBlocks * kBBlockBytes;
  static constexpr int kBiasK = ((kK + 3) / 4) * 4;                           // K columns kBiasK, kBiasK+1 carry the bias
  static constexpr int kPatchElems = 7 * kPatchW * CIN;
  static constexpr uint32_t kPatchBytes = (((kPatchElems + kTileM * CIN) * 4 + 15) / 16) * 16;   // + zero tail for K padding
  static_assert(kBiasK + 2 <= kKP, "no free K columns for the bias");
  static constexpr uint32_t kTmemCols = 2 * COUT;                             // power of two >= 32
  // epilogue through TMA bulk stores: per warp and 32-channel chunk one (kStageRows pixels x 32 channels) fp32 box
  static constexpr int kStageRows = COUT == 128 ? 32 : 16;                     // rows per staged box (what fits in smem)
  static constexpr uint32_t kStageWarpBytes = kStageRows * 128;
  static constexpr uint32_t kStageBytes = 4 * kStageWarpBytes;
  static constexpr uint32_t kSmemBytes = 1024 + 2 * kABytes + kBBytes + kStageBytes + kPatchBytes + kKP * 4 + 128;
};

struct ConvSmallParams {
  const float* x;
  int64_t sn, sy, sx, sc;       // element strides of x
  const float* w_packed;        // (COUT, kKP), k = (ky*7 + kx)*CIN + c, zero padded
  const float* bias;            // (COUT) or null
  float* y;                     // (B, H, W, COUT) NHWC
  int H, W;
  int64_t tiles;                // B * H * (W / 128)
  int relu;
  int debug;                    // MRFA_CONV_DEBUG bits (timing experiments only): 1 no stores, 2 no im2col, 4 no patch loads
};

// byte offset of element (row, k) inside a K-major SWIZZLE_128B tile whose K blocks are
// `block_bytes` apart: 128-byte rows, 8-row groups of 1024 bytes, 16-byte chunk j of row r at
// chunk position j ^ (r & 7)
__device__ __forceinline__ uint32_t sw128_offset(int row, int k, uint32_t block_bytes) {
  const int kb = k >> 5, j = (k >> 2) & 7, wi = k & 3;
  return (uint32_t)kb * block_bytes + (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u +
         (uint32_t)((j ^ (row & 7)) << 4) + (uint32_t)wi * 4u;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

template <int CIN, int COUT>
__global__ void __launch_bounds__(kConvThreads, 1)
conv7x7_small_kernel(const ConvSmallParams prm, const __grid_constant__ CUtensorMap map_y) {
  using Cfg = ConvSmallCfg<CIN, COUT>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;                                   // [2][kKBlocks][128 x 128 B]
  uint8_t* smem_b = smem + 2 * Cfg::kABytes;                // [kKBlocks][COUT x 128 B]
  uint8_t* smem_st = smem_b + Cfg::kBBytes;                 // [4 warps][kStageRows x 128 B] (TMA-store staging)
  float* patch = reinterpret_cast<float*>(smem_st + Cfg::kStageBytes);
  int* koff = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(patch) + Cfg::kPatchBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(koff + Cfg::kKP);
  uint64_t* a_full = bars;          // [2] producers -> MMA        (one arrival per producer thread)
  uint64_t* a_empty = bars + 2;     // [2] MMA commit -> producers
  uint64_t* t_full = bars + 4;      // [2] MMA commit -> epilogue
  uint64_t* t_empty = bars + 6;     // [2] epilogue -> MMA          (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  // ---- one-time setup: barriers, TMEM, resident weights, im2col offset table, zeroed A pad ----
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], kProducerWarps);
      mbar_init(&a_empty[i], 1);
      mbar_init(&t_full[i], 1);
      mbar_init(&t_empty[i], 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  for (int i = threadIdx.x; i < COUT * Cfg::kKP; i += kConvThreads) {
    const int n = i / Cfg::kKP, k = i - n * Cfg::kKP;
    float wv = to_tf32(__ldg(prm.w_packed + i));
    if (prm.bias != nullptr && (k == Cfg::kBiasK || k == Cfg::kBiasK + 1)) {
      // the bias rides in two spare K columns as a TF32 hi + lo pair (A holds 1.0 there): fp32-accurate
      const float bv = __ldg(prm.bias + n), hi = to_tf32(bv);
      wv = k == Cfg::kBiasK ? hi : to_tf32(bv - hi);
    }
    *reinterpret_cast<float*>(smem_b + sw128_offset(n, k, Cfg::kBBlockBytes)) = wv;
  }
  for (int k = threadIdx.x; k < Cfg::kKP; k += kConvThreads) {
    const int tap = k / CIN, c = k - tap * CIN, ky = tap / 7, kx = tap - ky * 7;
    koff[k] = k < Cfg::kK ? (ky * kPatchW + kx) * CIN + c : -1;
  }
  for (uint32_t i = threadIdx.x; i < 2 * Cfg::kABytes / 16; i += kConvThreads)
    reinterpret_cast<uint4*>(smem_a)[i] = make_uint4(0, 0, 0, 0);
  for (int i = threadIdx.x; i < kTileM * CIN; i += kConvThreads) patch[Cfg::kPatchElems + i] = 0.f;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * kTileM * 2; i += kConvThreads) {         // A[stage][row][kBiasK + {0,1}] = 1
    const int st = i / (2 * kTileM), row = (i >> 1) % kTileM, k = Cfg::kBiasK + (i & 1);
    *reinterpret_cast<float*>(smem_a + (size_t)st * Cfg::kABytes + sw128_offset(row, k, kKBlockBytes)) = 1.f;
  }
  fence_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tiles_per_row = prm.W / kTileM;

  if (warp < kProducerWarps) {
    // ================= producers: patch load + im2col into the swizzled A tile =================
    const int tid = threadIdx.x;                     // 0..kProducers-1
    const int r8 = lane & 7, wi = lane >> 3;         // store mapping: 8 rows x 4 words per warp instruction
    const bool chan_inner = prm.sc == 1;             // NHWC input: channels fastest in memory
    constexpr int kPatchElems = Cfg::kPatchElems;
    constexpr int kLoads = (kPatchElems + kProducers - 1) / kProducers;
    constexpr int kRowGroups = kTileM / 8 / kProducerWarps;
    // patch slot -> (global offset from the tile origin, row, column) is the same for every tile
    int goff[kLoads];                                // (fits: offsets stay inside one image)
    int soff[kLoads];                                // smem index, or -1 past the end
    int rp[kLoads];                                  // (row - 3) << 16 | (column - 3) & 0xffff
#pragma unroll
    for (int j = 0; j < kLoads; ++j) {
      const int i = tid + j * kProducers;
      int r, px, c;
      if (chan_inner) {
        c = i % CIN; px = (i / CIN) % kPatchW; r = i / (CIN * kPatchW);
      } else {
        px = i % kPatchW; c = (i / kPatchW) % CIN; r = i / (CIN * kPatchW);
      }
      goff[j] = (int)((r - 3) * prm.sy + (px - 3) * prm.sx + c * prm.sc);
      soff[j] = i < kPatchElems ? (r * kPatchW + px) * CIN + c : -1;
      rp[j] = ((r - 3) << 16) | ((px - 3) & 0xffff);
    }
    // im2col: lane (r8, cg) writes whole 16-byte chunks cg, cg+4, ... of row r8 of each 8-row group, so a
    // warp store covers 8 rows x 4 chunks (conflict-free under the 128-byte swizzle).  Source offsets are
    // relative to the row's patch origin; K padding reads the zero tail behind the patch.
    const int cg = wi;
    constexpr int kChunks = Cfg::kBiasK / 4;
    constexpr int kIters = (kChunks + 3) / 4;
    int ksrc[kIters][4];
#pragma unroll
    for (int i = 0; i < kIters; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = (cg + 4 * i) * 4 + e;
        ksrc[i][e] = (k < Cfg::kK ? koff[k] : Cfg::kPatchElems) * 4;
      }
    const uint32_t patch_u32 = smem_u32(patch);

    float pv[kLoads];
    auto load_patch = [&](int64_t t) {
      const int xt = (int)(t % tiles_per_row);
      const int y = (int)((t / tiles_per_row) % prm.H);
      const int64_t b = t / ((int64_t)tiles_per_row * prm.H);
      const int x0 = xt * kTileM;
      const float* org = prm.x + b * prm.sn + (int64_t)y * prm.sy + (int64_t)x0 * prm.sx;
#pragma unroll
      for (int j = 0; j < kLoads; ++j) {
        const int yy = y + (rp[j] >> 16), xx = x0 + (int)(short)(rp[j] & 0xffff);
        const bool ok = soff[j] >= 0 && yy >= 0 && yy < prm.H && xx >= 0 && xx < prm.W;
        pv[j] = (ok && !(prm.debug & 4)) ? __ldg(org + goff[j]) : 0.f;      // zero outside the image = the convolution's padding
      }
    };

    uint32_t it = 0;
    if ((int64_t)blockIdx.x < prm.tiles) load_patch(blockIdx.x);
    for (int64_t t = blockIdx.x; t < prm.tiles; t += gridDim.x, ++it) {
      const int s = it & 1;
      // (1) park the prefetched patch in shared memory, then prefetch the next tile's while building
#pragma unroll
      for (int j = 0; j < kLoads; ++j)
        if (soff[j] >= 0) sts_f32(patch_u32 + soff[j] * 4, to_tf32(pv[j]));
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");
      if (t + gridDim.x < prm.tiles) load_patch(t + gridDim.x);
      // (2) wait until the MMAs that read this A stage two tiles ago have retired
      mbar_wait(&a_empty[s], ((it >> 1) & 1u) ^ 1u);
      const uint32_t a_u32 = smem_u32(smem_a + (size_t)s * Cfg::kABytes);
#pragma unroll
      for (int g = 0; g < kRowGroups; ++g) {
        if (prm.debug & 2) break;
        const int m = (warp * kRowGroups + g) * 8 + r8;
        const uint32_t prow = patch_u32 + m * CIN * 4;
        const uint32_t arow = a_u32 + (uint32_t)(m >> 3) * 1024u + (uint32_t)r8 * 128u;
#pragma unroll
        for (int i = 0; i < kIters; ++i) {
          const int c4 = cg + 4 * i;                 // chunk index along K
          if (c4 < kChunks) {
            const float v0 = lds_f32(prow + ksrc[i][0]), v1 = lds_f32(prow + ksrc[i][1]);
            const float v2 = lds_f32(prow + ksrc[i][2]), v3 = lds_f32(prow + ksrc[i][3]);
            st_shared_v4(arow + (uint32_t)(c4 >> 3) * kKBlockBytes + (uint32_t)(((c4 & 7) ^ r8) << 4),
                         __float_as_uint(v0), __float_as_uint(v1), __float_as_uint(v2), __float_as_uint(v3));
          }
        }
      }
      fence_async_smem();                            // generic-proxy writes -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      asm volatile("bar.sync 1, %0;" ::"n"(kProducers) : "memory");   // the patch may now be overwritten
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(kTileM, COUT);
      uint32_t it = 0;
      for (int64_t t = blockIdx.x; t < prm.tiles; t += gridDim.x, ++it) {
        const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
        mbar_wait(&t_empty[s], ph ^ 1u);
        mbar_wait(&a_full[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + (size_t)s * Cfg::kABytes);
        const uint32_t b_addr = smem_u32(smem_b);
        const uint32_t d = tmem_base + s * COUT;
#pragma unroll
        for (int kb = 0; kb < Cfg::kKBlocks; ++kb) {
          const uint64_t da = umma_desc_sw128(a_addr + kb * kKBlockBytes);
          const uint64_t db = umma_desc_sw128(b_addr + kb * Cfg::kBBlockBytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)                 // 8 tf32 = 32 bytes per MMA: +2 in the (>>4) address field
            tcgen05_mma_tf32(d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        }
        tcgen05_commit(&a_empty[s]);
        tcgen05_commit(&t_full[s]);
      }
    }
  } else {
    // ================= epilogue: TMEM -> registers -> bias / ReLU -> NHWC =================
    const int ew = warp - kProducerWarps;            // == warp % 4: the TMEM lane quarter this warp may read
    uint32_t it = 0;
    for (int64_t t = blockIdx.x; t < prm.tiles; t += gridDim.x, ++it) {
      const uint32_t s = it & 1u, ph = (it >> 1) & 1u;
      mbar_wait(&t_full[s], ph);
      tcgen05_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + s * COUT;
#pragma unroll 1
      for (int ch = 0; ch < COUT / 32; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + ch * 32, v);
        tmem_ld_wait();
        if (ch == COUT / 32 - 1) {
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&t_empty[s]);
        }
        if (prm.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));   // bias came with the MMA
        }
        {
          // stage kStageRows pixels x 32 channels (128-byte rows, SWIZZLE_128B) and hand the box to the TMA engine
          const uint32_t st = smem_u32(smem_st + ew * Cfg::kStageWarpBytes);
#pragma unroll
          for (int h = 0; h < 32; h += Cfg::kStageRows) {
            if (lane == 0) tma_store_wait_read();        // the previous box has left shared memory
            __syncwarp();
            if (lane >= h && lane < h + Cfg::kStageRows) {
              const uint32_t row = (uint32_t)(lane - h);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                st_shared_v4(st + row * 128u + (uint32_t)((j ^ (row & 7)) << 4), v[4 * j], v[4 * j + 1], v[4 * j + 2],
                             v[4 * j + 3]);
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0 && !(prm.debug & 1)) {
              tma_store_3d(&map_y, smem_st + ew * Cfg::kStageWarpBytes, ch * 32, (int)(t * kTileM + ew * 32 + h), 0);
              tma_store_commit();
            }
          }
        }
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace mrfa

using namespace mrfa;

template <int CIN, int COUT>
static int launch_conv_small(const ConvSmallParams& prm, int sm_count, cudaStream_t st) {
  using Cfg = ConvSmallCfg<CIN, COUT>;
  auto kern = conv7x7_small_kernel<CIN, COUT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  const int64_t grid = prm.tiles < sm_count ? prm.tiles : sm_count;
  CUtensorMap map_y;
  memset(&map_y, 0, sizeof(map_y));
  {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return MRFA_E_DRIVER;
    const cuuint64_t pixels = (cuuint64_t)prm.tiles * kTileM;
    cuuint64_t dims[3] = {(cuuint64_t)COUT, pixels, 1};
    cuuint64_t strides[2] = {(cuuint64_t)COUT * 4, pixels * COUT * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)Cfg::kStageRows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map_y, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, prm.y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return MRFA_E_DRIVER;
  }
  kern<<<(unsigned)grid, kConvThreads, Cfg::kSmemBytes, st>>>(prm, map_y);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_conv7x7_small_kpad(int Cin) { return Cin > 0 ? ((49 * Cin + 31) / 32) * 32 : 0; }

extern "C" int mrfa_conv7x7_small(const float* x, mrfa_grid_strides_t xs, const float* w_packed, const float* bias,
                                  float* y, int B, int Cin, int Cout, int H, int W, int relu, int sm_count,
                                  mrfa_stream_t stream) {
  MRFA_CHECK_ARG(x && w_packed && y && B >= 0 && H > 0 && W > 0 && sm_count > 0);
  MRFA_CHECK_SHAPE(W % kTileM == 0 && (int64_t)B * H * W < ((int64_t)1 << 31));
  MRFA_CHECK_SHAPE((Cin == 2 && Cout == 128) || (Cin == 3 && Cout == 64));
  if ((reinterpret_cast<uintptr_t>(y) & 31) != 0) return MRFA_E_ALIGN;
  if (B == 0) return 0;
  ConvSmallParams prm;
  prm.x = x; prm.sn = xs.sn; prm.sy = xs.sy; prm.sx = xs.sx; prm.sc = xs.sc;
  prm.w_packed = w_packed; prm.bias = bias; prm.y = y; prm.H = H; prm.W = W;
  prm.tiles = (int64_t)B * H * (W / kTileM);
  prm.relu = relu;
  const char* dbg = getenv("MRFA_CONV_DEBUG");
  prm.debug = dbg ? atoi(dbg) : 0;
  if (Cin == 2) return launch_conv_small<2, 128>(prm, sm_count, as_stream(stream));
  return launch_conv_small<3, 64>(prm, sm_count, as_stream(stream));
}
