// Fused per-channel / per-pixel elementwise passes around the warps (SURVEY.md section 8(f) row
// N2: the decoder's occlusion blending `warp*occ + out*(1-occ)` over the largest feature maps,
// and the pre-activation BatchNorm(eval)+ReLU / bias(+residual) passes between cuDNN
// convolutions).  Pure streaming kernels: one float4 per thread per iteration, grid-stride,
// HBM-bound; algorithmic bytes = (inputs + output) * 4.
#include "common.cuh"

namespace mrfa {

__device__ __forceinline__ float act_fn(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return 1.f / (1.f + expf(-v));
  return v;
}

// NHWC: element i has channel i % C; C % 4 == 0 so a float4 never straddles a pixel
__global__ void __launch_bounds__(256)
channel_affine_nhwc_kernel(const float4* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                           const float4* __restrict__ residual, float4* __restrict__ y, int64_t n4, int C, int act) {
  const int cq = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cq) * 4;
    float4 v = __ldg(x + i);
    if (scale != nullptr) {
      const float4 s = __ldg(reinterpret_cast<const float4*>(scale + c));
      v.x *= s.x; v.y *= s.y; v.z *= s.z; v.w *= s.w;
    }
    if (shift != nullptr) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(shift + c));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    if (residual != nullptr) {
      const float4 r = __ldg(residual + i);
      v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
    }
    v.x = act_fn(v.x, act); v.y = act_fn(v.y, act); v.z = act_fn(v.z, act); v.w = act_fn(v.w, act);
    y[i] = v;
  }
}

// NCHW (or any C): scalar, channel = (i / HW) % C
__global__ void __launch_bounds__(256)
channel_affine_nchw_kernel(const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                           const float* __restrict__ residual, float* __restrict__ y, int64_t n, int C, int HW, int act) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    float v = __ldg(x + i);
    if (scale != nullptr) v *= __ldg(scale + c);
    if (shift != nullptr) v += __ldg(shift + c);
    if (residual != nullptr) v += __ldg(residual + i);
    y[i] = act_fn(v, act);
  }
}

// y = a * o + b * (1 - o)   (b == nullptr: y = a * o), o per pixel (broadcast over channels)
__global__ void __launch_bounds__(256)
occlusion_blend_nhwc_kernel(const float4* __restrict__ a, const float4* __restrict__ b, const float* __restrict__ occ,
                            float4* __restrict__ y, int64_t n4, int C) {
  const int cq = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float o = __ldg(occ + i / cq);
    float4 v = __ldg(a + i);
    v.x *= o; v.y *= o; v.z *= o; v.w *= o;
    if (b != nullptr) {
      const float4 w = __ldg(b + i);
      const float q = 1.f - o;
      v.x = fmaf(w.x, q, v.x); v.y = fmaf(w.y, q, v.y); v.z = fmaf(w.z, q, v.z); v.w = fmaf(w.w, q, v.w);
    }
    y[i] = v;
  }
}

__global__ void __launch_bounds__(256)
occlusion_blend_nchw_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ occ,
                            float* __restrict__ y, int64_t n, int C, int HW) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t img = i / ((int64_t)C * HW);
    const float o = __ldg(occ + img * HW + i % HW);
    float v = __ldg(a + i) * o;
    if (b != nullptr) v = fmaf(__ldg(b + i), 1.f - o, v);
    y[i] = v;
  }
}


// Tail of the generator (generator.py:61-63) behind the space-to-depth final convolution: pixel shuffle + bias + sigmoid +
// the last occlusion blend in ONE pass.  conv (B, H/r, W/r, C*r*r) is the NHWC output of the 3x3 convolution that stands in
// for `final` (blocks.py::_final_s2d, no bias), channel c*r*r + (Y%r)*r + X%r of block (Y/r, X/r) is pixel (Y, X) of plane c;
// y[b,c,Y,X] = a[b,c,Y,X] * occ[b,Y,X] + sigmoid(conv + bias[c]) * (1 - occ[b,Y,X]), a / y NCHW planes.  Replaces the bias
// pass, the pixel-shuffle copy, the sigmoid pass and the blend (4 launches, 140 us at B = 64) by one.
__global__ void __launch_bounds__(256)
final_blend_s2d_kernel(const float* __restrict__ conv, const float* __restrict__ bias, const float* __restrict__ a,
                       const float* __restrict__ occ, float* __restrict__ y, int64_t pixels, int C, int H, int W, int r,
                       FastDiv fw, FastDiv fh, FastDiv fr) {
  const int HW = H * W, Wb = W / r, rr = r * r;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = (uint32_t)i;
    const uint32_t row = fast_div(u, fw);                       // b * H + Y
    const int X = (int)(u - row * (uint32_t)W);
    const uint32_t b = fast_div(row, fh);
    const int Y = (int)(row - b * (uint32_t)H);
    const int Yb = (int)fast_div((uint32_t)Y, fr), Xb = (int)fast_div((uint32_t)X, fr);
    const float* cv = conv + (((int64_t)b * (H / r) + Yb) * Wb + Xb) * ((int64_t)C * rr) + (Y - Yb * r) * r + (X - Xb * r);
    const float o = __ldg(occ + i), q = 1.f - o;
    const int64_t p0 = (int64_t)b * C * HW + (int64_t)Y * W + X;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(cv + c * rr) + __ldg(bias + c);
      const float sg = 1.f / (1.f + expf(-v));
      y[p0 + (int64_t)c * HW] = fmaf(sg, q, __ldg(a + p0 + (int64_t)c * HW) * o);
    }
  }
}

// r == 4 (the production block size): one thread per run of four pixels of a row = the four consecutive floats of one
// channel group in `conv`; every access is a 16-byte vector.
__global__ void __launch_bounds__(256)
final_blend_s2d_r4_kernel(const float* __restrict__ conv, const float* __restrict__ bias, const float4* __restrict__ a,
                          const float4* __restrict__ occ, float4* __restrict__ y, int64_t runs, int C, int H, int W,
                          FastDiv fwb, FastDiv fh) {
  const int Wb = W >> 2, HW4 = (H * W) >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < runs; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = (uint32_t)i;
    const uint32_t row = fast_div(u, fwb);                      // b * H + Y
    const int Xb = (int)(u - row * (uint32_t)Wb);
    const uint32_t b = fast_div(row, fh);
    const int Y = (int)(row - b * (uint32_t)H);
    const float* cv = conv + (((int64_t)b * (H >> 2) + (Y >> 2)) * Wb + Xb) * ((int64_t)C * 16) + (Y & 3) * 4;
    const float4 o = __ldg(occ + i);
    const int64_t p0 = (int64_t)b * C * HW4 + (int64_t)Y * Wb + Xb;
    for (int c = 0; c < C; ++c) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cv + c * 16));
      const float4 w = __ldg(a + p0 + (int64_t)c * HW4);
      const float bc = __ldg(bias + c);
      float4 out;
      out.x = fmaf(1.f / (1.f + expf(-(v.x + bc))), 1.f - o.x, w.x * o.x);
      out.y = fmaf(1.f / (1.f + expf(-(v.y + bc))), 1.f - o.y, w.y * o.y);
      out.z = fmaf(1.f / (1.f + expf(-(v.z + bc))), 1.f - o.z, w.z * o.z);
      out.w = fmaf(1.f / (1.f + expf(-(v.w + bc))), 1.f - o.w, w.w * o.w);
      y[p0 + (int64_t)c * HW4] = out;
    }
  }
}

// Bilinear resize with align_corners=True (F.interpolate as used at raft.py:243) fused with an
// optional activation: SURVEY.md 8(f) row N1.  A 1x1 convolution commutes with this resize, so
// the decoder applies convc1 at the basic resolution and lets this kernel produce
// relu(upsample(.)) directly -- the (B,98,R,R) upsampled correlation features are never written.
// Same interpolation arithmetic as ATen upsample_bilinear2d (scale = (in-1)/(out-1) in fp32).
__device__ __forceinline__ void resize_axis(int o, float scale, int in, int& i0, int& i1, float& l1) {
  const float src = scale * (float)o;
  i0 = min((int)src, in - 1);
  i1 = min(i0 + 1, in - 1);
  l1 = src - (float)i0;
}

// One thread produces kResizeRun consecutive output pixels of one row for one channel quad and keeps the four
// source taps in registers, reloading a column only when the left tap index advances: for the x2 / x4
// up-samplings of the decoder that is ~0.5-1 instead of 4 vector loads per output (the plain form is bound by
// L1/L2 load traffic at 4x the output bytes).  Lanes run along the channel quads, so every load and store of
// a warp is one contiguous pixel.
constexpr int kResizeRun = 8;

// The two source rows are blended first (one fused multiply-add per component when a column enters the window), so an
// output costs two operations per component instead of seven: hx * (hy v00 + ly v10) + lx * (hy v01 + ly v11) is the
// same bilinear form as ATen's hy (hx v00 + lx v01) + ly (hx v10 + lx v11), rounded in a different order (~1e-7
// relative; the kernel is issue-bound, not HBM-bound, with the ATen order).
__device__ __forceinline__ float4 resize_vblend(const float* r0, const float* r1, float hy, float ly) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(r0)), b = __ldg(reinterpret_cast<const float4*>(r1));
  return make_float4(fmaf(ly, b.x, hy * a.x), fmaf(ly, b.y, hy * a.y), fmaf(ly, b.z, hy * a.z), fmaf(ly, b.w, hy * a.w));
}

template <int ACT>
__global__ void __launch_bounds__(256)
resize_bilinear_nhwc_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W, int Ho, int Wo,
                            float sy, float sx, int64_t total_runs, IndexSplit sp, const float* __restrict__ bias) {
  const int cq = C / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_runs; i += (int64_t)gridDim.x * blockDim.x) {
    int c4, run, oy;
    int64_t n, unused;
    split_index(i, sp, c4, run, oy, n, unused);        // (channel quad, run of the row, output row, sample)
    const int c = c4 * 4;
    int y0, y1;
    float ly;
    resize_axis(oy, sy, H, y0, y1, ly);
    const float hy = 1.f - ly;
    const float* r0 = x + (n * H + y0) * (int64_t)W * C + c;
    const float* r1 = x + (n * H + y1) * (int64_t)W * C + c;
    float4* dst = reinterpret_cast<float4*>(y + ((n * Ho + oy) * (int64_t)Wo + run * kResizeRun) * C + c);
    int cx0 = -1, cx1 = -1;
    float4 vl = make_float4(0, 0, 0, 0), vr = vl;      // vertically blended left / right source columns
    // a per-channel bias commutes with the interpolation (the four weights sum to one): added after it, before the activation
    const float4 bq = bias != nullptr ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < kResizeRun; ++j) {
      const int ox = run * kResizeRun + j;
      if (ox >= Wo) break;
      int x0, x1;
      float lx;
      resize_axis(ox, sx, W, x0, x1, lx);
      if (x0 != cx0) {
        vl = (x0 == cx1) ? vr : resize_vblend(r0 + (int64_t)x0 * C, r1 + (int64_t)x0 * C, hy, ly);   // slide or fetch
        cx0 = x0;
      }
      if (x1 != cx1) {
        vr = (x1 == x0) ? vl : resize_vblend(r0 + (int64_t)x1 * C, r1 + (int64_t)x1 * C, hy, ly);
        cx1 = x1;
      }
      const float hx = 1.f - lx;
      float4 r = make_float4(fmaf(lx, vr.x, hx * vl.x) + bq.x, fmaf(lx, vr.y, hx * vl.y) + bq.y, fmaf(lx, vr.z, hx * vl.z) + bq.z,
                             fmaf(lx, vr.w, hx * vl.w) + bq.w);
      if (ACT != 0) { r.x = act_fn(r.x, ACT); r.y = act_fn(r.y, ACT); r.z = act_fn(r.z, ACT); r.w = act_fn(r.w, ACT); }
      dst[(int64_t)j * cq] = r;
    }
  }
}

// NHWC with any channel count (flows / occlusion maps coming out of channels-last convolutions)
__global__ void __launch_bounds__(256)
resize_bilinear_nhwc_scalar_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int H, int W, int Ho,
                                   int Wo, float sy, float sx, int64_t total, int act, const float* __restrict__ bias) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t pix = i / C;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int64_t n = pix / ((int64_t)Wo * Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_axis(oy, sy, H, y0, y1, ly);
    resize_axis(ox, sx, W, x0, x1, lx);
    const float* base = x + n * H * W * C + c;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v = hy * (hx * __ldg(base + ((int64_t)y0 * W + x0) * C) + lx * __ldg(base + ((int64_t)y0 * W + x1) * C)) +
                    ly * (hx * __ldg(base + ((int64_t)y1 * W + x0) * C) + lx * __ldg(base + ((int64_t)y1 * W + x1) * C));
    y[i] = act_fn(bias != nullptr ? v + __ldg(bias + c) : v, act);
  }
}

__global__ void __launch_bounds__(256)
resize_bilinear_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int Ho, int Wo, float sy,
                            float sx, int64_t total, int act, const float* __restrict__ bias, int C) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wo);
    const int oy = (int)((i / Wo) % Ho);
    const int64_t plane = i / ((int64_t)Wo * Ho);
    int y0, y1, x0, x1;
    float ly, lx;
    resize_axis(oy, sy, H, y0, y1, ly);
    resize_axis(ox, sx, W, x0, x1, lx);
    const float* p = x + plane * H * W;
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float v = hy * (hx * __ldg(p + y0 * W + x0) + lx * __ldg(p + y0 * W + x1)) +
                    ly * (hx * __ldg(p + y1 * W + x0) + lx * __ldg(p + y1 * W + x1));
    y[i] = act_fn(bias != nullptr ? v + __ldg(bias + (int)(plane % C)) : v, act);
  }
}


// Planar maps whose output rows are whole float4s (the 1-2 channel flow / occlusion maps resized to every decoder
// resolution): one thread per 4 consecutive outputs of a row -- the row taps and the plane decomposition are evaluated
// once per 4 outputs and the store is one 16-byte vector.
__global__ void __launch_bounds__(256)
resize_bilinear_nchw_v4_kernel(const float* __restrict__ x, float4* __restrict__ y, int H, int W, int Ho, int Wo4, float sy,
                               float sx, int64_t total4, int act, int64_t y_row_pitch4) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % Wo4);
    const int oy = (int)((i / Wo4) % Ho);
    const int64_t plane = i / ((int64_t)Wo4 * Ho);
    int y0, y1;
    float ly;
    resize_axis(oy, sy, H, y0, y1, ly);
    const float hy = 1.f - ly;
    const float* r0 = x + (plane * H + y0) * W;
    const float* r1 = x + (plane * H + y1) * W;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int x0, x1;
      float lx;
      resize_axis(4 * q + j, sx, W, x0, x1, lx);
      const float hx = 1.f - lx;
      o[j] = act_fn(hy * (hx * __ldg(r0 + x0) + lx * __ldg(r0 + x1)) + ly * (hx * __ldg(r1 + x0) + lx * __ldg(r1 + x1)), act);
    }
    // y_row_pitch4: float4s per output row of the destination (== Wo4 for a dense map; wider when the map is one column
    // block of a strip, mrfa_resize_bilinear_strip)
    y[(plane * Ho + oy) * y_row_pitch4 + q] = make_float4(o[0], o[1], o[2], o[3]);
  }
}


// Flow / occlusion carry to the next (2x finer) level (raft.py:276-295).
// The reference issues ~10 align_corners bilinear resizes of 1-2 channel maps plus as many
// elementwise kernels per level; here one thread per fine pixel evaluates the whole update:
//   d_f   = 2 * up(d_flow[:, 0:2])                 flow = d_f + up(init_flow) / scale
//   d_o   = up(d_flow[:, 2:3])                     occ  = d_o + up(prior_occ)
//   with a previous accumulated update (level > 0):
//   up_f  = 2 * up(d_f_pre), up_o = up(d_occ_pre)  flow += up_f, occ += up_o,
//   d_f_acc = d_f + up_f, d_occ_acc = d_o + up_o   (else d_f_acc = d_f, d_occ_acc = d_o)
// Every product / sum is rounded separately, in the order of the reference expression.
struct CarryTaps {
  int y0, y1, x0, x1;
  float ly, lx;
};

__device__ __forceinline__ float carry_interp(const float* __restrict__ p, int sy, int sx, const CarryTaps& t) {
  const float hy = 1.f - t.ly, hx = 1.f - t.lx;
  return hy * (hx * __ldg(p + t.y0 * sy + t.x0 * sx) + t.lx * __ldg(p + t.y0 * sy + t.x1 * sx)) +
         t.ly * (hx * __ldg(p + t.y1 * sy + t.x0 * sx) + t.lx * __ldg(p + t.y1 * sy + t.x1 * sx));
}

__global__ void __launch_bounds__(256)
flow_carry_kernel(const float* __restrict__ d_flow, mrfa_grid_strides_t ds, const float* __restrict__ init_flow,
                  const float* __restrict__ prior_occ, const float* __restrict__ d_f_pre,
                  const float* __restrict__ d_occ_pre, float* __restrict__ flow, float* __restrict__ occ,
                  float* __restrict__ d_f_acc, float* __restrict__ d_occ_acc, int R, int h, float scale, int cl,
                  float s_r, float s_h, int64_t total, IndexSplit sp) {
  const int Ro = 2 * R;
  const int dsy = (int)ds.sy, dsx = (int)ds.sx;          // strides inside one sample: checked to fit 32 bits at launch
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int zero, ox, oy;
    int64_t n, same;
    split_index(i, sp, zero, ox, oy, n, same);           // sp.c divides by 1: i -> (ox, oy, n)
    CarryTaps tr, th;
    resize_axis(oy, s_r, R, tr.y0, tr.y1, tr.ly);
    resize_axis(ox, s_r, R, tr.x0, tr.x1, tr.lx);
    resize_axis(oy, s_h, h, th.y0, th.y1, th.ly);
    resize_axis(ox, s_h, h, th.x0, th.x1, th.lx);
    const float* df = d_flow + n * ds.sn;
    const float dfx = __fmul_rn(carry_interp(df, dsy, dsx, tr), 2.f);
    const float dfy = __fmul_rn(carry_interp(df + ds.sc, dsy, dsx, tr), 2.f);
    const float dov = carry_interp(df + 2 * ds.sc, dsy, dsx, tr);
    const float* fi = init_flow + n * 2 * h * h;
    float fx = __fadd_rn(dfx, __fdiv_rn(carry_interp(fi, h, 1, th), scale));
    float fy = __fadd_rn(dfy, __fdiv_rn(carry_interp(fi + (int64_t)h * h, h, 1, th), scale));
    float oc = __fadd_rn(dov, carry_interp(prior_occ + n * h * h, h, 1, th));
    float ax = dfx, ay = dfy, ao = dov;
    if (d_f_pre != nullptr) {
      // the accumulated update of the previous level lives at R x R in the caller's memory format
      const float* pp = d_f_pre + n * 2 * R * R;
      const int psy = cl ? 2 * R : R, psx = cl ? 2 : 1, psc = cl ? 1 : R * R;
      const float ux = __fmul_rn(carry_interp(pp, psy, psx, tr), 2.f);
      const float uy = __fmul_rn(carry_interp(pp + psc, psy, psx, tr), 2.f);
      const float uo = carry_interp(d_occ_pre + n * R * R, R, 1, tr);
      fx = __fadd_rn(fx, ux); fy = __fadd_rn(fy, uy); oc = __fadd_rn(oc, uo);
      ax = __fadd_rn(dfx, ux); ay = __fadd_rn(dfy, uy); ao = __fadd_rn(dov, uo);
    }
    const int64_t pix = (int64_t)oy * Ro + ox, plane = (int64_t)Ro * Ro;
    if (cl) {
      *reinterpret_cast<float2*>(flow + (n * plane + pix) * 2) = make_float2(fx, fy);
      *reinterpret_cast<float2*>(d_f_acc + (n * plane + pix) * 2) = make_float2(ax, ay);
    } else {
      flow[n * 2 * plane + pix] = fx; flow[n * 2 * plane + plane + pix] = fy;
      d_f_acc[n * 2 * plane + pix] = ax; d_f_acc[n * 2 * plane + plane + pix] = ay;
    }
    occ[n * plane + pix] = oc;
    d_occ_acc[n * plane + pix] = ao;
  }
}


// Per-level flow / occlusion update (raft.py:256-262): flow_w = flow + d_flow[:, 0:2]; occlusion += d_flow[:, 2:3];
// sigmoid(occlusion) for the decoder -- three strided elementwise passes and a sigmoid in the reference's op order,
// one thread per pixel here.  flow / flow_w are planar (B,2,R,R) or pixel-interleaved (channels_last).
__global__ void __launch_bounds__(256)
flow_update_kernel(const float* __restrict__ flow, const float* __restrict__ occ, const float* __restrict__ d_flow,
                   mrfa_grid_strides_t ds, float* __restrict__ flow_w, float* __restrict__ occ_new,
                   float* __restrict__ occ_sig, int64_t total, int cl, IndexSplit sp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int zero, x, y;
    int64_t n, same;
    split_index(i, sp, zero, x, y, n, same);
    const int64_t plane = (int64_t)sp.w.d * sp.h.d, pix = i - n * plane;
    const float* d = d_flow + n * ds.sn + (int64_t)y * ds.sy + (int64_t)x * ds.sx;
    const float dx = __ldg(d), dy = __ldg(d + ds.sc), dz = __ldg(d + 2 * ds.sc);
    if (cl) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + i);
      reinterpret_cast<float2*>(flow_w)[i] = make_float2(__fadd_rn(f.x, dx), __fadd_rn(f.y, dy));
    } else {
      flow_w[n * 2 * plane + pix] = __fadd_rn(__ldg(flow + n * 2 * plane + pix), dx);
      flow_w[n * 2 * plane + plane + pix] = __fadd_rn(__ldg(flow + n * 2 * plane + plane + pix), dy);
    }
    const float o = __fadd_rn(__ldg(occ + i), dz);
    occ_new[i] = o;
    occ_sig[i] = act_fn(o, 2);
  }
}


// AntiAliasInterpolation2d (util.py:282-326): zero-padded depthwise
// KxK Gaussian followed by nearest sub-sampling by `stride`.  The reference convolves at full
// resolution and throws 15/16 of the result away; this evaluates only the kept pixels.
// NCHW in / out; out[b,c,y,x] = sum_ij w[c,i,j] * in[b,c, y*stride + i - ka, x*stride + j - ka].
__global__ void __launch_bounds__(256)
antialias_down_kernel(const float* __restrict__ in, const float* __restrict__ weight, float* __restrict__ out, int C,
                      int H, int W, int Ho, int Wo, int K, int ka, int stride, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % Wo);
    const int y = (int)((i / Wo) % Ho);
    const int64_t plane = i / ((int64_t)Wo * Ho);
    const int c = (int)(plane % C);
    const float* src = in + plane * H * W;
    const float* wc = weight + (int64_t)c * K * K;
    const int y0 = y * stride - ka, x0 = x * stride - ka;
    float acc = 0.f;
    for (int a = 0; a < K; ++a) {
      const int yy = y0 + a;
      if (yy < 0 || yy >= H) continue;
      for (int b = 0; b < K; ++b) {
        const int xx = x0 + b;
        if (xx >= 0 && xx < W) acc = fmaf(__ldg(src + (int64_t)yy * W + xx), __ldg(wc + a * K + b), acc);
      }
    }
    out[i] = acc;
  }
}


// The configuration the path runs (scale 0.25: stride 4, K = 13, ka = 6): a block produces a 16 x 16 output tile of one
// plane from a 76 x 76 input patch staged in shared memory (zero-filled outside the image, so skipped taps of the
// generic kernel become exact +0 terms of the same fused-multiply-add chain).  The patch is stored de-interleaved by
// x mod 4 with a row pitch of 20 words: lane tx reads column 4 tx + b at plane b & 3, word tx + (b >> 2) -- consecutive
// lanes, consecutive banks; the two output rows of a warp sit 16 banks apart.
constexpr int kAaTile = 16, kAaStride = 4, kAaK = 13, kAaKa = 6;
constexpr int kAaPatch = kAaTile * kAaStride + kAaK - 1;          // 76 input rows / columns per tile
constexpr int kAaPitch = 20;                                      // words per de-interleaved row (19 used)

__global__ void __launch_bounds__(256)
antialias_down_s4k13_kernel(const float* __restrict__ in, const float* __restrict__ weight, float* __restrict__ out, int C,
                            int H, int W, int Ho, int Wo) {
  __shared__ float patch[4][kAaPatch][kAaPitch];
  __shared__ float wk[kAaK * kAaK];
  const int tiles_x = (Wo + kAaTile - 1) / kAaTile;
  const int tile_x = blockIdx.x % tiles_x, tile_y = blockIdx.x / tiles_x;
  const int64_t plane = blockIdx.y;
  const float* src = in + plane * H * W;
  const int gy0 = tile_y * kAaTile * kAaStride - kAaKa, gx0 = tile_x * kAaTile * kAaStride - kAaKa;
  for (int e = threadIdx.x; e < kAaPatch * kAaPatch; e += 256) {
    const int py = e / kAaPatch, px = e - py * kAaPatch;
    const int yy = gy0 + py, xx = gx0 + px;
    patch[px & 3][py][px >> 2] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (int64_t)yy * W + xx) : 0.f;
  }
  if (threadIdx.x < kAaK * kAaK) wk[threadIdx.x] = __ldg(weight + (int64_t)(plane % C) * kAaK * kAaK + threadIdx.x);
  __syncthreads();
  const int tx = threadIdx.x % kAaTile, ty = threadIdx.x / kAaTile;
  const int ox = tile_x * kAaTile + tx, oy = tile_y * kAaTile + ty;
  float acc = 0.f;
#pragma unroll 1
  for (int a = 0; a < kAaK; ++a) {
    const int py = ty * kAaStride + a;
#pragma unroll
    for (int b = 0; b < kAaK; ++b) acc = fmaf(patch[b & 3][py][tx + (b >> 2)], wk[a * kAaK + b], acc);
  }
  if (ox < Wo && oy < Ho) out[(plane * Ho + oy) * Wo + ox] = acc;
}


// Occlusion blend whose second operand is a "sub-pixel" up-convolution result: a 3x3 conv on a
// x2 nearest-upsampled map equals four 2x2 convs on the low-resolution map (one per output
// parity), so UpBlock2d (util.py:160-177) runs as ONE 2x2 conv with 4*C outputs on the padded
// low-res input (16/36 of the FLOPs, no upsampled tensor) and this kernel reads phase (a,b) of
// full-res pixel (Y,X) at b2[n, Y/2 + a, X/2 + b, (2a+b)*C + c] with a = Y&1, b = X&1.
// y = a * occ + b2_shuffled * (1 - occ);  a, y: (N, 2H, 2W, C) NHWC;  b2: (N, H+1, W+1, 4C) NHWC.
// out_block r > 1 writes y in r x r space-to-depth order -- pixel (Y,X) at
// [n, Y/r, X/r, ((Y%r)*r + X%r)*C + c], i.e. an (N, r*r*C, 2H/r, 2W/r) NHWC tensor -- so the generator's final
// 7x7 convolution (generator.py:32,61) can run as a 3x3 convolution with r*r*3 outputs.
__global__ void __launch_bounds__(256)
occlusion_blend_subpixel_kernel(const float4* __restrict__ a, const float* __restrict__ b2, const float* __restrict__ occ,
                                float4* __restrict__ y, int64_t n4, int C, int H, int W, int r, int64_t ostride, IndexSplit sp,
                                FastDiv fr) {
  const int cq = C / 4, W2 = 2 * W, H2 = 2 * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4, X, Y;
    int64_t n, pix;
    split_index(i, sp, c4, X, Y, n, pix);
    const int c = c4 * 4;
    const int pa = Y & 1, pb = X & 1;
    const float o = __ldg(occ + pix);
    const float4 w = __ldg(reinterpret_cast<const float4*>(
        b2 + ((n * (H + 1) + (Y >> 1) + pa) * (W + 1) + (X >> 1) + pb) * (4 * (int64_t)C) + (2 * pa + pb) * C + c));
    float4 v = __ldg(a + i);
    const float q = 1.f - o;
    v.x = fmaf(w.x, q, v.x * o); v.y = fmaf(w.y, q, v.y * o); v.z = fmaf(w.z, q, v.z * o); v.w = fmaf(w.w, q, v.w * o);
    if (r == 1) {
      y[(pix * ostride + c) / 4] = v;
    } else {
      const int Yb = (int)fast_div((uint32_t)Y, fr), Xb = (int)fast_div((uint32_t)X, fr);
      const int64_t blk = (n * (H2 / r) + Yb) * (W2 / r) + Xb;
      y[(blk * (r * r) + (Y - Yb * r) * r + (X - Xb * r)) * cq + c4] = v;
    }
  }
}


// Hourglass decoder step (util.py:239-263: out = cat([up_block(out), skip], 1)) with the up-block run as
// the sub-pixel 2x2 convolution above: this kernel de-interleaves the four phases of b2 into
// channels [0, C) of the full-resolution map and copies the skip connection into [C, C+Cs) --
// the nearest-upsampled tensor, the 3x3 convolution on it and the separate cat pass disappear.
// b2 (N, H+1, W+1, 4C) NHWC; skip (N, Cs, 2H, 2W) with element strides; y (N, 2H, 2W, C+Cs) NHWC.
__global__ void __launch_bounds__(256)
subpixel_shuffle_cat_v4_kernel(const float* __restrict__ b2, const float4* __restrict__ skip, float4* __restrict__ y,
                               int64_t n4, int C, int Cs, int H, int W, IndexSplit sp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4, X, Y;
    int64_t n, pix;
    split_index(i, sp, c4, X, Y, n, pix);
    const int c = c4 * 4;
    float4 v;
    if (c < C) {
      const int pa = Y & 1, pb = X & 1;
      v = __ldg(reinterpret_cast<const float4*>(
          b2 + ((n * (H + 1) + (Y >> 1) + pa) * (W + 1) + (X >> 1) + pb) * (4 * (int64_t)C) + (2 * pa + pb) * C + c));
    } else {
      v = __ldg(skip + (pix * Cs + (c - C)) / 4);
    }
    y[i] = v;
  }
}

// Any channel counts / any skip strides (the hourglass inputs: 10, 13 or 44 channels, NCHW): a warp owns one output pixel
// per step, so the pixel decomposition is evaluated once per pixel instead of once per element, lanes run along the
// channels (contiguous 4C-channel phase block of b2, contiguous output pixel), and the strided reads of a planar skip map
// fall into sectors that the neighbouring warps of the block (the next pixels of the row) reuse out of L1.
__global__ void __launch_bounds__(256)
subpixel_shuffle_cat_kernel(const float* __restrict__ b2, const float* __restrict__ skip, mrfa_grid_strides_t ss,
                            float* __restrict__ y, int64_t pixels, int C, int Cs, int H, int W, IndexSplit sp) {
  const int ct = C + Cs;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t pix = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); pix < pixels; pix += nwarps) {
    int zero, X, Y;
    int64_t n, same;
    split_index(pix, sp, zero, X, Y, n, same);          // sp.c divides by 1: pix -> (X, Y, n)
    const int pa = Y & 1, pb = X & 1;
    const float* src = b2 + ((n * (H + 1) + (Y >> 1) + pa) * (W + 1) + (X >> 1) + pb) * (4 * (int64_t)C) + (2 * pa + pb) * C;
    const float* sk = skip + n * ss.sn + (int64_t)Y * ss.sy + (int64_t)X * ss.sx;
    float* dst = y + pix * ct;
    for (int c = lane; c < C; c += 32) dst[c] = __ldg(src + c);
    for (int c = lane; c < Cs; c += 32) dst[C + c] = __ldg(sk + (int64_t)c * ss.sc);
  }
}


// Backward of the 2x2 average pool in NHWC (training path): each output gradient is spread as g/4 over its
// window; one float4 load and four float4 stores per thread.  (ATen's NHWC avg_pool2d_backward takes 0.23 ms
// per launch at batch 16, 4.6 ms per training step.)
__global__ void __launch_bounds__(256)
avg_pool2x2_nhwc_bwd_kernel(const float4* __restrict__ gy, float* __restrict__ gx, int C, int H, int W, int64_t n4) {
  const int cq = C / 4, Wo = W / 2, Ho = H / 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cq) * 4;
    const int64_t pix = i / cq;
    const int ox = (int)(pix % Wo);
    const int oy = (int)((pix / Wo) % Ho);
    const int64_t n = pix / ((int64_t)Wo * Ho);
    float4 g = __ldg(gy + i);
    g.x *= 0.25f; g.y *= 0.25f; g.z *= 0.25f; g.w *= 0.25f;
    float* p = gx + ((n * H + 2 * oy) * W + 2 * ox) * (int64_t)C + c;
    *reinterpret_cast<float4*>(p) = g;
    *reinterpret_cast<float4*>(p + C) = g;
    *reinterpret_cast<float4*>(p + (int64_t)W * C) = g;
    *reinterpret_cast<float4*>(p + (int64_t)W * C + C) = g;
  }
}


// cat([a, b], dim=1) of two NHWC maps (the update block's cat([cor, flo]) and cat([motion, context]),
// raft.py:66, :83): one float4 per thread, every warp instruction a contiguous run of one pixel row.
// (ATen's CatArrayBatchedCopy reaches ~4 TB/s on these 2-4 GB copies.)
__global__ void __launch_bounds__(256)
cat2_nhwc_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, int64_t n4, int Ca,
                 int Cb) {
  const int qa = Ca / 4, qt = (Ca + Cb) / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(i % qt);
    const int64_t pix = i / qt;
    y[i] = q < qa ? __ldg(a + pix * qa + q) : __ldg(b + pix * (qt - qa) + (q - qa));
  }
}


// 2x2 average pooling in NHWC (DownBlock2d, util.py:190-196): one float4 per thread, 4 loads +
// 1 store, window summed row-major then divided by 4 like ATen.  (The stock NHWC avg_pool2d
// kernel runs at ~0.8 TB/s on the 2 GB encoder maps.)
__global__ void __launch_bounds__(256)
avg_pool2x2_nhwc_kernel(const float* __restrict__ x, float4* __restrict__ y, int C, int H, int W, int64_t n4, IndexSplit sp) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int c4, ox, oy;
    int64_t n, pix;
    split_index(i, sp, c4, ox, oy, n, pix);
    const int c = c4 * 4;
    const float* p = x + ((n * H + 2 * oy) * W + 2 * ox) * (int64_t)C + c;
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p + C));
    const float4 d = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * C));
    const float4 e = __ldg(reinterpret_cast<const float4*>(p + (int64_t)W * C + C));
    float4 r;
    r.x = (((a.x + b.x) + d.x) + e.x) * 0.25f; r.y = (((a.y + b.y) + d.y) + e.y) * 0.25f;
    r.z = (((a.z + b.z) + d.z) + e.z) * 0.25f; r.w = (((a.w + b.w) + d.w) + e.w) * 0.25f;
    y[i] = r;
  }
}

static inline unsigned stream_blocks(int64_t items) {
  int64_t b = cdiv64(items, 256);
  return (unsigned)(b < 1 ? 1 : (b > 148 * 32 ? 148 * 32 : b));
}

}  // namespace mrfa

using namespace mrfa;

extern "C" int mrfa_channel_affine(const float* x, const float* scale, const float* shift, const float* residual,
                                   float* y, int64_t pixels, int C, int HW, int channels_last, int act,
                                   mrfa_stream_t stream) {
  MRFA_CHECK_ARG(x && y && pixels >= 0 && C > 0 && HW > 0 && act >= 0 && act <= 2);
  if (pixels == 0) return 0;
  const int64_t n = pixels * C;
  if (channels_last) {
    MRFA_CHECK_SHAPE(C % 4 == 0);
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(residual) |
          reinterpret_cast<uintptr_t>(scale) | reinterpret_cast<uintptr_t>(shift)) & 15) != 0)
      return MRFA_E_ALIGN;
    channel_affine_nhwc_kernel<<<stream_blocks(n / 4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(x), scale, shift, reinterpret_cast<const float4*>(residual),
        reinterpret_cast<float4*>(y), n / 4, C, act);
  } else {
    channel_affine_nchw_kernel<<<stream_blocks(n), 256, 0, as_stream(stream)>>>(x, scale, shift, residual, y, n, C, HW, act);
  }
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_occlusion_blend(const float* a, const float* b, const float* occ, float* y, int64_t pixels, int C,
                                    int HW, int channels_last, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(a && occ && y && pixels >= 0 && C > 0 && HW > 0);
  if (pixels == 0) return 0;
  const int64_t n = pixels * C;
  if (channels_last) {
    MRFA_CHECK_SHAPE(C % 4 == 0);
    if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
      return MRFA_E_ALIGN;
    occlusion_blend_nhwc_kernel<<<stream_blocks(n / 4), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b), occ, reinterpret_cast<float4*>(y), n / 4, C);
  } else {
    occlusion_blend_nchw_kernel<<<stream_blocks(n), 256, 0, as_stream(stream)>>>(a, b, occ, y, n, C, HW);
  }
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_final_blend_s2d(const float* conv, const float* bias, const float* a, const float* occ, float* y, int B,
                                    int C, int H, int W, int r, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(conv && bias && a && occ && y && B >= 0 && C > 0 && H > 0 && W > 0 && r >= 1);
  MRFA_CHECK_SHAPE(H % r == 0 && W % r == 0 && (int64_t)B * H * W < ((int64_t)1 << 31));
  if (B == 0) return 0;
  const int64_t pixels = (int64_t)B * H * W;
  if (r == 4 && ((reinterpret_cast<uintptr_t>(conv) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(occ) |
                  reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    final_blend_s2d_r4_kernel<<<stream_blocks(pixels / 4), 256, 0, as_stream(stream)>>>(
        conv, bias, reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(occ), reinterpret_cast<float4*>(y),
        pixels / 4, C, H, W, make_fastdiv((uint32_t)(W / 4)), make_fastdiv((uint32_t)H));
    return MRFA_LAUNCH_RESULT();
  }
  final_blend_s2d_kernel<<<stream_blocks(pixels), 256, 0, as_stream(stream)>>>(
      conv, bias, a, occ, y, pixels, C, H, W, r, make_fastdiv((uint32_t)W), make_fastdiv((uint32_t)H), make_fastdiv((uint32_t)r));
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_resize_bilinear(const float* x, const float* bias, float* y, int N, int C, int H, int W, int Ho, int Wo,
                                    int channels_last, int act, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(x && y && N >= 0 && C > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && act >= 0 && act <= 2);
  if (N == 0) return 0;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total = (int64_t)N * C * Ho * Wo;
  if (channels_last && C % 4 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(bias)) & 15) == 0) {
    const int64_t runs = (int64_t)N * Ho * ((Wo + kResizeRun - 1) / kResizeRun) * (C / 4);
    const IndexSplit sp = make_index_split(C / 4, (Wo + kResizeRun - 1) / kResizeRun, Ho);
    const unsigned g = stream_blocks(runs);
    cudaStream_t st = as_stream(stream);
    if (act == 0) resize_bilinear_nhwc_kernel<0><<<g, 256, 0, st>>>(x, y, C, H, W, Ho, Wo, sy, sx, runs, sp, bias);
    else if (act == 1) resize_bilinear_nhwc_kernel<1><<<g, 256, 0, st>>>(x, y, C, H, W, Ho, Wo, sy, sx, runs, sp, bias);
    else resize_bilinear_nhwc_kernel<2><<<g, 256, 0, st>>>(x, y, C, H, W, Ho, Wo, sy, sx, runs, sp, bias);
  } else if (channels_last) {
    resize_bilinear_nhwc_scalar_kernel<<<stream_blocks(total), 256, 0, as_stream(stream)>>>(x, y, C, H, W, Ho, Wo, sy, sx,
                                                                                            total, act, bias);
  } else if (bias == nullptr && Wo % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    resize_bilinear_nchw_v4_kernel<<<stream_blocks(total / 4), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<float4*>(y), H, W,
                                                                                           Ho, Wo / 4, sy, sx, total / 4, act, Wo / 4);
  } else {
    resize_bilinear_nchw_kernel<<<stream_blocks(total), 256, 0, as_stream(stream)>>>(x, y, H, W, Ho, Wo, sy, sx, total, act, bias, C);
  }
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_resize_bilinear_strip(const float* x, float* y, int64_t planes, int H, int W, int Ho, int Wo,
                                         int64_t y_row_pitch, int64_t y_col_offset, int act, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(x && y && planes >= 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && act >= 0 && act <= 2);
  MRFA_CHECK_ARG(y_col_offset >= 0 && y_row_pitch >= y_col_offset + Wo);
  MRFA_CHECK_SHAPE(Wo % 4 == 0 && y_row_pitch % 4 == 0 && y_col_offset % 4 == 0);
  if ((reinterpret_cast<uintptr_t>(y) & 15) != 0) return MRFA_E_ALIGN;
  if (planes == 0) return 0;
  const float sy = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
  const float sx = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
  const int64_t total4 = planes * Ho * (Wo / 4);
  resize_bilinear_nchw_v4_kernel<<<stream_blocks(total4), 256, 0, as_stream(stream)>>>(
      x, reinterpret_cast<float4*>(y + y_col_offset), H, W, Ho, Wo / 4, sy, sx, total4, act, y_row_pitch / 4);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_antialias_down(const float* in, const float* weight, float* out, int N, int C, int H, int W, int K,
                                   int ka, int stride, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(in && weight && out && N >= 0 && C > 0 && H > 0 && W > 0 && K > 0 && ka >= 0 && stride > 0);
  if (N == 0) return 0;
  const int Ho = H / stride, Wo = W / stride;       // floor(H * scale) of F.interpolate(scale_factor = 1/stride)
  MRFA_CHECK_SHAPE(Ho > 0 && Wo > 0);
  const int64_t total = (int64_t)N * C * Ho * Wo;
  if (stride == kAaStride && K == kAaK && ka == kAaKa && (int64_t)N * C <= 65535) {
    dim3 grid((unsigned)(((Wo + kAaTile - 1) / kAaTile) * ((Ho + kAaTile - 1) / kAaTile)), (unsigned)(N * C));
    antialias_down_s4k13_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, weight, out, C, H, W, Ho, Wo);
    return MRFA_LAUNCH_RESULT();
  }
  antialias_down_kernel<<<stream_blocks(total), 256, 0, as_stream(stream)>>>(in, weight, out, C, H, W, Ho, Wo, K, ka, stride, total);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_occlusion_blend_subpixel(const float* a, const float* b2, const float* occ, float* y, int N, int C,
                                             int H, int W, int out_block, int64_t out_pixel_stride, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(a && b2 && occ && y && N >= 0 && C > 0 && H > 0 && W > 0 && out_block >= 1);
  MRFA_CHECK_ARG(out_pixel_stride == 0 || (out_block == 1 && out_pixel_stride >= C && out_pixel_stride % 4 == 0));
  MRFA_CHECK_SHAPE(C % 4 == 0 && (2 * H) % out_block == 0 && (2 * W) % out_block == 0);
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
    return MRFA_E_ALIGN;
  if (N == 0) return 0;
  const int64_t n4 = (int64_t)N * 4 * H * W * C / 4;
  occlusion_blend_subpixel_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(a), b2, occ, reinterpret_cast<float4*>(y), n4, C, H, W, out_block,
      out_pixel_stride > 0 ? out_pixel_stride : (int64_t)C, make_index_split(C / 4, 2 * W, 2 * H), make_fastdiv((uint32_t)out_block));
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_avg_pool2x2_nhwc(const float* x, float* y, int N, int C, int H, int W, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(x && y && N >= 0 && C > 0 && H >= 2 && W >= 2);
  MRFA_CHECK_SHAPE(C % 4 == 0);
  if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) != 0) return MRFA_E_ALIGN;
  if (N == 0) return 0;
  const int64_t n4 = (int64_t)N * (H / 2) * (W / 2) * C / 4;
  avg_pool2x2_nhwc_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(x, reinterpret_cast<float4*>(y), C, H, W, n4,
                                                                            make_index_split(C / 4, W / 2, H / 2));
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_flow_carry(const float* d_flow, mrfa_grid_strides_t d_strides, const float* init_flow,
                               const float* prior_occ, const float* d_f_pre, const float* d_occ_pre, float* flow,
                               float* occ, float* d_f_acc, float* d_occ_acc, int B, int R, int h, float scale,
                               int channels_last, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(d_flow && init_flow && prior_occ && flow && occ && d_f_acc && d_occ_acc);
  MRFA_CHECK_ARG((d_f_pre == nullptr) == (d_occ_pre == nullptr));
  MRFA_CHECK_ARG(B >= 0 && R > 0 && h > 0 && scale != 0.f);
  MRFA_CHECK_SHAPE(R <= 16384 && h <= 16384 && d_strides.sy * R < ((int64_t)1 << 31) && d_strides.sx * R < ((int64_t)1 << 31));
  if (B == 0) return 0;
  if (channels_last && ((reinterpret_cast<uintptr_t>(flow) | reinterpret_cast<uintptr_t>(d_f_acc)) & 7) != 0)
    return MRFA_E_ALIGN;
  const int Ro = 2 * R;
  const float s_r = (float)(R - 1) / (float)(Ro - 1);
  const float s_h = (float)(h - 1) / (float)(Ro - 1);
  const int64_t total = (int64_t)B * Ro * Ro;
  flow_carry_kernel<<<stream_blocks(total), 256, 0, as_stream(stream)>>>(d_flow, d_strides, init_flow, prior_occ, d_f_pre,
                                                                         d_occ_pre, flow, occ, d_f_acc, d_occ_acc, R, h,
                                                                         scale, channels_last, s_r, s_h, total,
                                                                         make_index_split(1, Ro, Ro));
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_flow_update(const float* flow, const float* occ, const float* d_flow, mrfa_grid_strides_t d_strides,
                                float* flow_w, float* occ_new, float* occ_sig, int B, int H, int W, int channels_last,
                                mrfa_stream_t stream) {
  MRFA_CHECK_ARG(flow && occ && d_flow && flow_w && occ_new && occ_sig && B >= 0 && H > 0 && W > 0);
  if (B == 0) return 0;
  if (channels_last && ((reinterpret_cast<uintptr_t>(flow) | reinterpret_cast<uintptr_t>(flow_w)) & 7) != 0) return MRFA_E_ALIGN;
  const int64_t total = (int64_t)B * H * W;
  flow_update_kernel<<<stream_blocks(total), 256, 0, as_stream(stream)>>>(flow, occ, d_flow, d_strides, flow_w, occ_new, occ_sig,
                                                                          total, channels_last, make_index_split(1, W, H));
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_subpixel_shuffle_cat(const float* b2, const float* skip, mrfa_grid_strides_t skip_strides, float* y,
                                         int N, int C, int Cs, int H, int W, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(b2 && y && N >= 0 && C > 0 && Cs >= 0 && H > 0 && W > 0 && (skip || Cs == 0));
  if (N == 0) return 0;
  const int64_t total = (int64_t)N * 4 * H * W * (C + Cs);
  const bool dense_nhwc = skip_strides.sc == 1 && skip_strides.sx == Cs && skip_strides.sy == (int64_t)Cs * 2 * W &&
                          skip_strides.sn == (int64_t)Cs * 4 * H * W;
  if (C % 4 == 0 && Cs % 4 == 0 && (Cs == 0 || dense_nhwc) &&
      ((reinterpret_cast<uintptr_t>(b2) | reinterpret_cast<uintptr_t>(skip) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    subpixel_shuffle_cat_v4_kernel<<<stream_blocks(total / 4), 256, 0, as_stream(stream)>>>(
        b2, reinterpret_cast<const float4*>(skip), reinterpret_cast<float4*>(y), total / 4, C, Cs, H, W,
        make_index_split((C + Cs) / 4, 2 * W, 2 * H));
  } else {
    const int64_t pixels = (int64_t)N * 4 * H * W;
    subpixel_shuffle_cat_kernel<<<stream_blocks(pixels * 32), 256, 0, as_stream(stream)>>>(
        b2, skip, skip_strides, y, pixels, C, Cs, H, W, make_index_split(1, 2 * W, 2 * H));
  }
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_cat2_nhwc(const float* a, const float* b, float* y, int64_t pixels, int Ca, int Cb, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(a && b && y && pixels >= 0 && Ca > 0 && Cb > 0);
  MRFA_CHECK_SHAPE(Ca % 4 == 0 && Cb % 4 == 0);
  if (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(y)) & 15) != 0)
    return MRFA_E_ALIGN;
  if (pixels == 0) return 0;
  const int64_t n4 = pixels * (Ca + Cb) / 4;
  cat2_nhwc_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(a),
                                                                     reinterpret_cast<const float4*>(b),
                                                                     reinterpret_cast<float4*>(y), n4, Ca, Cb);
  return MRFA_LAUNCH_RESULT();
}

extern "C" int mrfa_avg_pool2x2_nhwc_bwd(const float* grad_y, float* grad_x, int N, int C, int H, int W, mrfa_stream_t stream) {
  MRFA_CHECK_ARG(grad_y && grad_x && N >= 0 && C > 0 && H >= 2 && W >= 2);
  MRFA_CHECK_SHAPE(C % 4 == 0 && H % 2 == 0 && W % 2 == 0);       // odd sizes would leave an unwritten border
  if (((reinterpret_cast<uintptr_t>(grad_y) | reinterpret_cast<uintptr_t>(grad_x)) & 15) != 0) return MRFA_E_ALIGN;
  if (N == 0) return 0;
  const int64_t n4 = (int64_t)N * (H / 2) * (W / 2) * C / 4;
  avg_pool2x2_nhwc_bwd_kernel<<<stream_blocks(n4), 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(grad_y), grad_x,
                                                                                C, H, W, n4);
  return MRFA_LAUNCH_RESULT();
}
