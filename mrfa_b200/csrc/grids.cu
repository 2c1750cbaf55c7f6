// Coordinate grids, key-point heat-maps and the prior->flow conversion (K8/K9, a10-a12, a16).
// Tiny elementwise kernels: one thread per output element, coalesced stores.  They exist as
// standalone entry points for API parity; inside the fused kernels the grids are recomputed
// from the thread index and never materialised.
#include "common.cuh"

namespace mrfa {

__global__ void coords_grid_kernel(float* __restrict__ out, int batch, int ht, int wd) {
  const int64_t total = (int64_t)batch * 2 * ht * wd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % wd);
    const int y = (int)((i / wd) % ht);
    const int c = (int)((i / ((int64_t)wd * ht)) % 2);
    out[i] = (float)(c == 0 ? x : y);
  }
}

__global__ void make_coordinate_grid_kernel(float* __restrict__ out, int h, int w) {
  const int64_t total = (int64_t)h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % w), y = (int)(i / w);
    reinterpret_cast<float2*>(out)[i] = make_float2(norm_coord(x, w), norm_coord(y, h));
  }
}

__global__ void kp2gaussian_kernel(const float* __restrict__ kp, const float* __restrict__ add, int add_period,
                                   float* __restrict__ out, int P, int h, int w, float variance) {
  const int64_t hw = (int64_t)h * w, total = (int64_t)P * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / hw);
    const int r = (int)(i - (int64_t)p * hw);
    const int y = r / w, x = r - y * w;
    const float dx = __fsub_rn(norm_coord(x, w), __ldg(kp + 2 * p));
    const float dy = __fsub_rn(norm_coord(y, h), __ldg(kp + 2 * p + 1));
    const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    float v = expf(__fdiv_rn(__fmul_rn(-0.5f, s), variance));
    if (add != nullptr) v = __fadd_rn(v, __ldg(add + (int64_t)(p % add_period) * hw + r));
    out[i] = v;
  }
}

// w % 4 == 0, 16-byte aligned maps: one thread per 4 consecutive pixels of a row (32-bit index arithmetic, the row
// coordinate and the key-point loaded once, one vector load of `add`, one vector store)
__global__ void __launch_bounds__(256)
kp2gaussian_v4_kernel(const float* __restrict__ kp, const float4* __restrict__ add, int add_period, float4* __restrict__ out,
                      uint32_t total4, uint32_t hw4, uint32_t w4, int h, int w, float variance) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += gridDim.x * blockDim.x) {
    const uint32_t p = i / hw4, r4 = i - p * hw4;
    const uint32_t y = r4 / w4, x0 = (r4 - y * w4) * 4;
    const float kx = __ldg(kp + 2 * p), ky = __ldg(kp + 2 * p + 1);
    const float dy = __fsub_rn(norm_coord((int)y, h), ky);
    const float dy2 = __fmul_rn(dy, dy);
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float dx = __fsub_rn(norm_coord((int)x0 + j, w), kx);
      v[j] = expf(__fdiv_rn(__fmul_rn(-0.5f, __fadd_rn(__fmul_rn(dx, dx), dy2)), variance));
    }
    if (add != nullptr) {
      const float4 a = __ldg(add + (p % (uint32_t)add_period) * hw4 + r4);
      v[0] = __fadd_rn(v[0], a.x); v[1] = __fadd_rn(v[1], a.y); v[2] = __fadd_rn(v[2], a.z); v[3] = __fadd_rn(v[3], a.w);
    }
    out[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

__global__ void prior_to_flow_kernel(const float* __restrict__ deformation, float* __restrict__ flow, int B, int h,
                                     int w, float hm1) {
  const int64_t hw = (int64_t)h * w, total = (int64_t)B * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / hw);
    const int r = (int)(i - (int64_t)b * hw);
    const int y = r / w, x = r - y * w;
    const float2 d = __ldg(reinterpret_cast<const float2*>(deformation) + i);
    // (h-1)*(d+1)/2.0 - id  : multiply, divide, subtract in that order (raft.py:190)
    flow[((int64_t)b * 2 + 0) * hw + r] = __fsub_rn(__fdiv_rn(__fmul_rn(hm1, __fadd_rn(d.x, 1.f)), 2.f), (float)x);
    flow[((int64_t)b * 2 + 1) * hw + r] = __fsub_rn(__fdiv_rn(__fmul_rn(hm1, __fadd_rn(d.y, 1.f)), 2.f), (float)y);
  }
}

static inline unsigned blocks_for(int64_t total) {
  int64_t b = cdiv64(total, 256);
  return (unsigned)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace mrfa

using namespace mrfa;

extern "C" int mrfa_coords_grid(float* out, int batch, int ht, int wd, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(out && batch >= 0 && ht > 0 && wd > 0);
  if (batch == 0) return 0;
  coords_grid_kernel<<<blocks_for((int64_t)batch * 2 * ht * wd), 256, 0, as_stream(stream)>>>(out, batch, ht, wd);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_make_coordinate_grid(float* out, int h, int w, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(out && h > 0 && w > 0);
  make_coordinate_grid_kernel<<<blocks_for((int64_t)h * w), 256, 0, as_stream(stream)>>>(out, h, w);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_kp2gaussian(const float* kp, const float* add, int add_period, float* out, int P, int h, int w,
                                float variance, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(kp && out && P >= 0 && h > 0 && w > 0);
  MRFA_CHECK_ARG(add == nullptr || add_period > 0);
  if (P == 0) return 0;
  const int64_t total = (int64_t)P * h * w;
  if (w % 4 == 0 && total < ((int64_t)1 << 32) &&
      ((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(add)) & 15) == 0) {
    kp2gaussian_v4_kernel<<<blocks_for(total / 4), 256, 0, as_stream(stream)>>>(
        kp, reinterpret_cast<const float4*>(add), add_period, reinterpret_cast<float4*>(out), (uint32_t)(total / 4),
        (uint32_t)(h * w / 4), (uint32_t)(w / 4), h, w, variance);
    return MRFA_LAUNCH_RESULT();
  }
  kp2gaussian_kernel<<<blocks_for((int64_t)P * h * w), 256, 0, as_stream(stream)>>>(kp, add, add_period, out, P, h, w, variance);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_prior_to_flow(const float* deformation, float* flow, int B, int h, int w, float hm1,
                                  mrfa_stream_t stream) {
  MRFA_CHECK_ARG(deformation && flow && B >= 0 && h > 0 && w > 0);
  if (B == 0) return 0;
  prior_to_flow_kernel<<<blocks_for((int64_t)B * h * w), 256, 0, as_stream(stream)>>>(deformation, flow, B, h, w, hm1);
  return MRFA_LAUNCH_RESULT();
}
