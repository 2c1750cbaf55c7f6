/*
 * mrfa_b200.h -- C ABI of libmrfa_b200.so: hand-written sm_100a kernels for the MRFA
 * per-frame-pair motion-refinement hot path.
 *
 * The reference (JialeTao/MRFA) is pure Python: its "FFI" for this path is the Python symbol
 * table of modules/util.py, modules/raft.py and modules/dense_motion.py.  Each entry point
 * below names the reference lines it replaces.  The Python binding that sits above this ABI
 * is mrfa_b200/_lib.py (ctypes) + mrfa_b200/ops.py (torch.library custom ops).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless stated otherwise; the caller owns all memory
 *     (outputs and workspaces are allocated by the caller and passed in); the library never
 *     allocates, frees or retains pointers and keeps no mutable global state;
 *   - every launch goes to `stream` (a cudaStream_t passed as void*) on the current device;
 *     no call synchronises;
 *   - return value: 0 = success; < 0 = argument error (MRFA_E_*); > 0 = a cudaError_t;
 *   - tensors are dense row-major ("contiguous") unless explicit element strides are passed;
 *   - all floating point is IEEE fp32 (no fast-math); the correlation volume is bf16.
 */
#ifndef MRFA_B200_H_
#define MRFA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MRFA_B200_ABI_VERSION 8

#define MRFA_E_BADARG   (-1)   /* null pointer, non-positive extent, unsupported enum  */
#define MRFA_E_SHAPE    (-2)   /* shape outside what the kernel is specialised for     */
#define MRFA_E_ALIGN    (-3)   /* pointer not aligned as documented                    */
#define MRFA_E_DRIVER   (-4)   /* cuTensorMapEncodeTiled entry point unavailable       */

typedef void* mrfa_stream_t;   /* cudaStream_t */

int         mrfa_abi_version(void);
const char* mrfa_error_string(int code);

/* ---------------------------------------------------------------------------------------
 * Sampling conventions for the bilinear warps
 * ------------------------------------------------------------------------------------- */
enum {
  MRFA_COORD_NORM_ACF = 0, /* normalised grid, align_corners=False: bare F.grid_sample(x, grid)
                              raft.py:166,168,271  dense_motion.py:83                         */
  MRFA_COORD_NORM_ACT = 1, /* normalised grid, align_corners=True: dense_motion.py:241        */
  MRFA_COORD_PIXEL    = 2  /* pixel coordinates: util.bilinear_sampler util.py:26-38 (replays
                              the 2*x/(W-1)-1 normalisation and the un-normalisation in fp32)  */
};
enum { MRFA_PAD_ZEROS = 0, MRFA_PAD_REFLECTION = 1 /* model.py:48 */ };

/* grid element (n,y,x,c) lives at grid[n*sn + y*sy + x*sx + c*sc]: lets the kernels consume
 * both (N,Ho,Wo,2) grids and channel-planar flow fields (B,2,R,R) viewed through permute.   */
typedef struct {
  int64_t sn, sy, sx, sc;
} mrfa_grid_strides_t;

/* a6/a8/a9/a17: out[n,c,y,x] = bilinear(in[n / in_batch_div, c], grid[n,y,x]).
 * in (N/in_batch_div, C, H, W); out (N, C, Ho, Wo).  If add_identity != 0 the pixel index
 * (x, y) is added to the grid value first (fuses `flow + coords_grid`, raft.py:247,260,302).
 * in_batch_div > 1 reads one input for several grids (fuses the `.repeat` of
 * dense_motion.py:80-81 / :238-239).
 * channels_last != 0: `in` and `out` are NHWC in memory (torch.channels_last; the layout the
 * sm_100 tensor-core convolutions produce), C % 4 == 0, 16-byte aligned: a tap is C contiguous
 * floats and every access is a 16-byte vector.  channels_last == 0: NCHW.                    */
int mrfa_grid_sample_fwd(const float* in, const float* grid, mrfa_grid_strides_t gs, float* out,
                         int N, int C, int H, int W, int Ho, int Wo, int in_batch_div,
                         int coord_mode, int padding_mode, int add_identity, int channels_last,
                         mrfa_stream_t stream);

/* Backward of the above.  grad_in must be ZERO-FILLED by the caller (scatter-add target);
 * grad_grid is written densely as (N,Ho,Wo,2).  Either may be NULL to skip it.               */
int mrfa_grid_sample_bwd(const float* grad_out, const float* in, const float* grid, mrfa_grid_strides_t gs,
                         float* grad_in, float* grad_grid,
                         int N, int C, int H, int W, int Ho, int Wo, int in_batch_div,
                         int coord_mode, int padding_mode, int add_identity, int channels_last,
                         mrfa_stream_t stream);

/* Two warps of the same feature map in one pass: refined (pixel flow + identity, raft.py:247)
 * and coarse/prior (normalised grid, align_corners=False, raft.py:271).  flow (B,2,Ho,Wo)
 * planar; prior_grid (B,Ho,Wo,2).  Reads `in` once from HBM.  coarse_pixel_stride (NHWC only; 0 = C): element
 * stride between pixels of out_coarse, so the coarse warp can be written straight into a channel slice of the
 * buffer the decoder would otherwise build with cat([y, warp_c]) (generator.py:51,60).                  */
int mrfa_dual_warp_fwd(const float* in, const float* flow, const float* prior_grid,
                       float* out_refined, float* out_coarse,
                       int N, int C, int H, int W, int channels_last, int64_t coarse_pixel_stride,
                       mrfa_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Coordinate grids and key-point heat-maps
 * ------------------------------------------------------------------------------------- */
/* a10 util.coords_grid util.py:53-56 -> (B,2,ht,wd), channel 0 = x, 1 = y (bit-exact).       */
int mrfa_coords_grid(float* out, int batch, int ht, int wd, mrfa_stream_t stream);
/* a11 util.make_coordinate_grid util.py:90-108 -> (h,w,2), x = 2*(j/(w-1))-1 (bit-exact).    */
int mrfa_make_coordinate_grid(float* out, int h, int w, mrfa_stream_t stream);
/* a12 util.kp2gaussian util.py:59-87: out[p,y,x] = exp((-0.5*|g-kp[p]|^2)/variance) (+ add).
 * kp (P,2); out (P,h,w).  `add` is NULL or a (add_period,h,w) tensor added with index
 * p % add_period (fuses `+ self.pos_embedding`, raft.py:177-178).                            */
int mrfa_kp2gaussian(const float* kp, const float* add, int add_period, float* out,
                     int P, int h, int w, float variance, mrfa_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Prior dense motion (a12-a15): heat-map difference, K+1 sparse motions, K+1 warped sources
 * ------------------------------------------------------------------------------------- */
/* DenseMotionNetwork.create_heatmap_representations / create_sparse_motions /
 * create_deformed_source_image dense_motion.py:36-85 in ONE pass over the (K+1) x h x w plane.
 *   kp_d, kp_s (B,K,2); jac_d, jac_s (B,K,2,2) or both NULL; bg_param (B,3,3) or NULL;
 *   source (B,C,h,w);
 *   motions  (B,K+1,h,w,2)            <- sparse motions, channel 0 = background
 *   hg_input (B,(K+1)*(C+1),h,w)      <- per k: [heat-map, deformed source (C planes)], the
 *                                         exact layout the hourglass consumes (:118-120)
 * variance is kp_variance (0.01).  The deformed source is sampled with align_corners=False.  */
int mrfa_dense_motion_prior(const float* kp_d, const float* kp_s, const float* jac_d, const float* jac_s,
                            const float* bg_param, const float* source,
                            float* motions, float* hg_input,
                            int B, int K, int C, int h, int w, float variance, mrfa_stream_t stream);

/* TPS.__init__(mode='kp') util.py:355-383: per (b,g) solve the (n+3)x(n+3) system, n == 5.
 * kp_1, kp_2 (B*G,5,2) -> theta (B*G,2,3), control_params (B*G,5,2).                        */
int mrfa_tps_solve(const float* kp_1, const float* kp_2, float* theta, float* control_params,
                   int BG, mrfa_stream_t stream);

/* TPSDenseMotionNetwork.create_heatmap_representations / create_transformations /
 * create_deformed_source_image dense_motion.py:200-243 in one pass.
 *   kp_d, kp_s (B,G*5,2); theta (B,G,2,3); control_params (B,G,5,2) from mrfa_tps_solve
 *   (control points are kp_d); bg_param (B,3,3) or NULL; source (B,C,h,w);
 *   motions  (B,G+1,h,w,2);
 *   hg_input (B, (G*5+1) + (G+1)*C, h, w): heat-maps first, then deformed sources (:274-275),
 *   sampled with align_corners=True.                                                         */
int mrfa_tps_motion_prior(const float* kp_d, const float* kp_s, const float* theta, const float* control_params,
                          const float* bg_param, const float* source,
                          float* motions, float* hg_input,
                          int B, int G, int C, int h, int w, float variance, mrfa_stream_t stream);

/* Backward of the prior-motion synthesis (the reference back-propagates dense_motion.py:36-85 / :200-243 and
 * util.py:59-87, :355-410 through autograd over (B,K,h,w)-sized eager ops; training, train.py:64-70).
 * All gradient outputs and `workspace` are caller-owned; `workspace` (the number of floats
 * mrfa_*_bwd_workspace returns), grad_kp_d, grad_kp_s and grad_source must be ZERO on entry (the kernels
 * accumulate with red.global.add); grad_jac_d / grad_jac_s / grad_bg are overwritten.
 *   grad_motions (same shape as motions) or NULL; grad_hg (same shape as hg_input);
 *   grad_bg NULL unless bg_param is given; grad_source NULL when the source needs no gradient.        */
int64_t mrfa_dense_motion_prior_bwd_workspace(int B, int K);
int mrfa_dense_motion_prior_bwd(const float* grad_motions, const float* grad_hg, const float* kp_d,
                                const float* kp_s, const float* jac_d, const float* jac_s,
                                const float* bg_param, const float* source, float* workspace,
                                float* grad_kp_d, float* grad_kp_s, float* grad_jac_d, float* grad_jac_s,
                                float* grad_bg, float* grad_source, int B, int K, int C, int h, int w,
                                float variance, mrfa_stream_t stream);
/* d kp of mrfa_kp2gaussian: grad (P,h,w) -> grad_kp (P,2), zero on entry.                      */
int mrfa_kp2gaussian_bwd(const float* grad, const float* kp, float* grad_kp, int P, int h, int w,
                         float variance, mrfa_stream_t stream);
/* mrfa_tps_solve + mrfa_tps_motion_prior together: the motion / heat-map / warp gradients are reduced per (b,g),
 * then the adjoint of the 8x8 solve (the same fp64 elimination on a new right-hand side; L is symmetric) carries
 * d theta / d control_params back to the key-points.  theta, control_params: the forward's outputs.   */
int64_t mrfa_tps_motion_prior_bwd_workspace(int B, int G);
int mrfa_tps_motion_prior_bwd(const float* grad_motions, const float* grad_hg, const float* kp_d,
                              const float* kp_s, const float* theta, const float* control_params,
                              const float* bg_param, const float* source, float* workspace,
                              float* grad_kp_d, float* grad_kp_s, float* grad_bg, float* grad_source,
                              int B, int G, int C, int h, int w, float variance, mrfa_stream_t stream);

/* a16 raft.py:189-190: flow[b,c,y,x] = (h-1)*(deformation[b,y,x,c]+1)/2 - (c ? y : x).
 * deformation (B,h,w,2) -> flow (B,2,h,w); `hm1` is self.h - 1 for both axes.                */
int mrfa_prior_to_flow(const float* deformation, float* flow, int B, int h, int w, float hm1,
                       mrfa_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * All-pairs structure correlation (a1-a4): pack -> tcgen05 GEMM with fused scale + pyramid
 * ------------------------------------------------------------------------------------- */
/* Row layout of the packed driving operand / of the volume: the driving plane at the basic
 * resolution (h x w rows), followed by its 2x, 4x, 8x average-pooled versions.  Pooling the
 * driving dims of the volume (raft.py:219) commutes with the contraction, so pooled query
 * rows give the pooled volume rows directly.
 *   rows_total = hw + hw/4 + hw/16 + hw/64 ; row offset of pool level l (k=2^l) is
 *   sum_{m<l} hw/4^m.                                                                        */
int64_t mrfa_corr_rows_total(int h, int w);
int64_t mrfa_corr_row_offset(int h, int w, int pool_log2);

/* Map layouts: where element (y, x) of one query's H x W source map sits inside its volume row.
 *   MRFA_MAP_ROWMAJOR  y*W + x  -- the reference's (B*Q,1,H,W) `corr` (raft.py:208) and every fp32 map.
 *   MRFA_MAP_TILED     bf16 pyramid maps only: 64-byte tiles of 4 rows x 8 columns, so the (2r+2)^2 lookup
 *                      footprint (raft.py:23-48) touches 4-6 DRAM granules instead of 8 scattered row pieces.
 *        level 0 (H x W, H % 8 == 0, W % 16 == 0): 2 x 2 tiles form a 16 x 8-pixel super-tile
 *            off0(y,x) = ((y>>3)*(W>>4) + (x>>4))*128 + ((y>>2)&1)*64 + ((x>>3)&1)*32 + (y&3)*8 + (x&7)
 *        level 1 (H/2 x W/2): plain tiles, tile t of level 1 = the 2x2 pool of super-tile t of level 0
 *            off1(y,x) = ((y>>2)*((W/2)>>3) + (x>>3))*32 + (y&3)*8 + (x&7)
 * mrfa_corr_map_layout(h, w) is the layout mrfa_corr_pack / mrfa_corr_volume produce for an h x w plane
 * (TILED for w in {64, 128}: the 256x256 and 512x512 configurations; ROWMAJOR for the small shapes).   */
enum { MRFA_MAP_ROWMAJOR = 0, MRFA_MAP_TILED = 1 };
int     mrfa_corr_map_layout(int h, int w);
int64_t mrfa_corr_map_offset(int map_layout, int level, int y, int x, int W_level);

/* fp32 NCHW features -> bf16 K-major GEMM operands (fuses the rearranges raft.py:183-184, the
 * fp32->bf16 cast and the driving-side average pooling raft.py:219).
 *   q_d, k_s (B,C,h,w) fp32;  a_op (B, rows_total, C) bf16;  b_op (B, h*w, C) bf16: source pixel (y,x)
 *   is row mrfa_corr_map_offset(mrfa_corr_map_layout(h,w), 0, y, x, w) -- permuting the B-operand rows is what
 *   makes the GEMM write every volume row directly in the map layout.
 * C % 64 == 0, h and w multiples of 8.  channels_last != 0: q_d / k_s are NHWC in memory (then
 * the rearrange is the identity and the kernel is a vectorised cast + pool).                 */
int mrfa_corr_pack(const float* q_d, const float* k_s, void* a_op, void* b_op,
                   int B, int C, int h, int w, int channels_last, mrfa_stream_t stream);

/* mrfa_corr_pack for NHWC inputs with the per-channel biases of the two 1x1 head convolutions that produce q_d / k_s
 * (raft.py:180-181 kp_img_head, kp_head) added on the fly: packs (q_d + q_bias) and (k_s + k_bias), so the heads can run
 * as plain GEMMs without a separate bias pass over the (B,C,h,w) maps.  q_bias / k_bias (C) fp32 or NULL, 16-byte aligned. */
int mrfa_corr_pack_bias(const float* q_d, const float* q_bias, const float* k_s, const float* k_bias, void* a_op, void* b_op,
                        int B, int C, int h, int w, mrfa_stream_t stream);

/* volume0[b,i,j]  = scale * sum_c a_op[b,i,c] * b_op[b,j,c]            (B, rows_total, h*w)  bf16
 * volume1[b,i,j'] = mean over the 2x2 source block j' of the above      (B, rows_total, hw/4) bf16
 * (raft.py:185 einsum * self.scale; CorrBlock.__init__ raft.py:20 avg_pool2d level 1).
 * tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in TMEM), operands by TMA, 128 x (2w | 128)
 * tiles so the source-side 2x2 pool is tile-local.  Requirements: C % 64 == 0, C <= 512,
 * w in {8,16,32,64,128}, (h*w) % 128 == 0, all pointers 16-byte aligned (cudaMalloc).
 * num_sms: SM count of the device (grid size of the persistent kernel); <= 0 -> 148.        */
int mrfa_corr_volume(const void* a_op, const void* b_op, void* volume0, void* volume1,
                     int B, int C, int h, int w, float scale, int num_sms, mrfa_stream_t stream);

/* Backward of mrfa_corr_pack + mrfa_corr_volume (the two GEMMs autograd derives from the einsum at raft.py:185, plus the
 * avg_pool2d backward of raft.py:20 and :219), for the training-step configuration.
 *   g0 (B, rows_total, hw) / g1 (B, rows_total, hw/4): fp32 gradients w.r.t. volume0 / volume1 in their map layout
 *   (what mrfa_corr_lookup_bwd scatter-adds).  With G = scale * (g0 + unpool(g1)/4):
 *     mrfa_corr_bwd_pack    G  (B, rows_total, hw) bf16 and GT (B, hw, rows_pad) bf16 = G transposed, rows_pad =
 *                           mrfa_corr_bwd_rows_pad(h, w) (pitch rounded up to 64; the pad columns are never read)
 *     mrfa_transpose_bf16   out[b, c, r] = in[b, r, c], out row pitch `out_pitch` >= rows  (operand transposes)
 *     mrfa_corr_bwd_gemm    D[b,m,n] = sum_k X[b,m,k] * Y[b,n,k]; X (B, M, ldx), Y (B, Nn, ldy) bf16 with K contiguous,
 *                           D (B, M, Nn) fp32; tcgen05.mma kind::f16, TMA operands, TMEM accumulators.
 *                           Nn % 64 == 0 and (Nn <= 256 or Nn % 256 == 0); ldx, ldy multiples of 8; X, Y 16-byte aligned.
 *                           dA = gemm(G, Bm^T) (M = rows_total, K = hw), dB = gemm(GT, A^T) (M = hw, K = rows_total).
 *     mrfa_corr_bwd_unpack  d_q[b,y,x,c] = dA[b, y*w+x, c] + sum_l dA[b, off_l + pooled(y,x), c] / 4^l ;
 *                           d_k[b,y,x,c] = dB[b, map_offset(y,x), c]; both (B, h, w, C) fp32 (NHWC memory).       */
int64_t mrfa_corr_bwd_rows_pad(int h, int w);
int mrfa_corr_bwd_pack(const float* g0, const float* g1, void* G, void* GT, int B, int h, int w, float scale,
                       mrfa_stream_t stream);
int mrfa_transpose_bf16(const void* in, void* out, int B, int rows, int cols, int out_pitch, mrfa_stream_t stream);
int mrfa_corr_bwd_gemm(const void* X, const void* Y, float* D, int B, int M, int Nn, int K, int64_t ldx, int64_t ldy,
                       int num_sms, mrfa_stream_t stream);
int mrfa_corr_bwd_unpack(const float* dA, const float* dB, float* d_q, float* d_k, int B, int C, int h, int w,
                         mrfa_stream_t stream);

/* Generic CorrBlock.__init__ level build for the drop-in API (raft.py:19-21):
 * out (P,1,H/2,W/2) = avg_pool2d(in (P,1,H,W), 2, 2), fp32.                                  */
int mrfa_avg_pool2x2(const float* in, float* out, int64_t P, int H, int W, mrfa_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Pyramid lookup (a5 CorrBlock.__call__ raft.py:23-48)
 * ------------------------------------------------------------------------------------- */
/* out[b, lvl*(2r+1)^2 + a*(2r+1) + b', q] = bilinear(level_lvl[map(b,q)], (x/2^lvl + a - r,
 * y/2^lvl + b' - r)), zeros padding, align_corners=True pixel convention (util.py:26-38).
 *   coords (B,2,Q) planar (channel 0 = x) in level-0 pixels, Q = h1*w1 queries per sample;
 *   the map of query (b,q) is level0 + (b*map_batch_stride + row_offset + q) * H*W, i.e.
 *   reference `corr` (B*Q,1,H,W) has map_batch_stride = Q, row_offset = 0; a pyramid volume
 *   from mrfa_corr_volume has map_batch_stride = rows_total, row_offset = pooled-level offset.
 *   level1 is the 2x2-pooled map (H/2 x W/2) with the same indexing.  map_layout: MRFA_MAP_* of both levels
 *   (TILED needs bf16 maps, H % 8 == 0, W % 16 == 0, 16-byte aligned levels).
 *   elem_bf16: 0 -> fp32 maps, 1 -> bf16 maps.   out (B, 2*(2r+1)^2, Q) fp32, or (B, Q, 2*(2r+1)^2)
 *   in memory when out_channels_last != 0 (feeds the 1x1 convc1 without a layout change).
 *   radius <= 4.                                                                              */
int mrfa_corr_lookup_fwd(const void* level0, const void* level1, int elem_bf16,
                         const float* coords, float* out,
                         int B, int Q, int H, int W, int64_t map_batch_stride, int64_t row_offset,
                         int radius, int map_layout, int out_channels_last, mrfa_stream_t stream);

/* Backward: grad_level0/1 (fp32, same indexing as the maps, ZERO-FILLED by the caller, may be
 * NULL) receive the scatter-add; grad_coords (B,2,Q) is written.                             */
int mrfa_corr_lookup_bwd(const float* grad_out, const void* level0, const void* level1, int elem_bf16,
                         const float* coords, float* grad_level0, float* grad_level1, float* grad_coords,
                         int B, int Q, int H, int W, int64_t map_batch_stride, int64_t row_offset,
                         int radius, int map_layout, mrfa_stream_t stream);

/* ---------------------------------------------------------------------------------------
 * Fused elementwise passes between the warps and the cuDNN convolutions
 * ------------------------------------------------------------------------------------- */
/* y = act(x * scale[c] + shift[c] + residual); scale / shift / residual may be NULL.
 * Replaces eval-mode BatchNorm2d + ReLU (util.py:126-127, :150-151), conv bias adds and the
 * residual add of ResBlock2d (util.py:156).  x, residual, y: (N,C,H,W) NCHW (channels_last=0)
 * or NHWC (channels_last=1, C % 4 == 0, 16-byte aligned); pixels = N*H*W; HW = H*W.
 * act: 0 identity, 1 ReLU, 2 sigmoid.  y may alias x.                                        */
int mrfa_channel_affine(const float* x, const float* scale, const float* shift, const float* residual,
                        float* y, int64_t pixels, int C, int HW, int channels_last, int act,
                        mrfa_stream_t stream);

/* generator.py:61-63 behind the space-to-depth final convolution (mrfa_occlusion_blend_subpixel with out_block r feeds a 3x3
 * convolution with C*r*r outputs): pixel shuffle + bias + sigmoid + the last occlusion blend in one pass.
 *   conv (B, H/r, W/r, C*r*r) NHWC, channel c*r*r + (Y%r)*r + X%r of block (Y/r, X/r) = pixel (Y,X) of plane c, no bias;
 *   bias (C); a, y (B,C,H,W) NCHW; occ (B,1,H,W):  y = a * occ + sigmoid(conv + bias[c]) * (1 - occ).              */
int mrfa_final_blend_s2d(const float* conv, const float* bias, const float* a, const float* occ, float* y, int B,
                         int C, int H, int W, int r, mrfa_stream_t stream);

/* Decoder occlusion blending generator.py:47,57: y = a * occ + b * (1 - occ) (b NULL: y = a * occ),
 * occ (N,1,H,W) broadcast over channels.  Same layouts as above.  y may alias a or b.        */
int mrfa_occlusion_blend(const float* a, const float* b, const float* occ, float* y,
                         int64_t pixels, int C, int HW, int channels_last, mrfa_stream_t stream);

/* Same blend with b given as a sub-pixel up-convolution result (UpBlock2d util.py:160-177 as one 2x2
 * conv with 4*C phase-major outputs on the padded low-resolution map): a, y (N,C,2H,2W) NHWC;
 * b2 (N,4C,H+1,W+1) NHWC, phase (Y&1, X&1) of pixel (Y,X) at b2[n, Y/2+(Y&1), X/2+(X&1), (2(Y&1)+(X&1))*C+c];
 * occ (N,1,2H,2W).  C % 4 == 0.  out_block r = 1: y plain NHWC; r > 1 (dividing 2H and 2W): y in r x r
 * space-to-depth order, pixel (Y,X) at y[n, Y/r, X/r, ((Y%r)*r + X%r)*C + c] = an (N, r*r*C, 2H/r, 2W/r) NHWC
 * tensor -- the layout in which the generator's final 7x7 convolution (generator.py:32,61) is a 3x3 one.
 * out_pixel_stride (out_block 1 only; 0 = C): element stride between pixels of y (a channel slice of a wider
 * NHWC buffer, see mrfa_dual_warp_fwd).                                                         */
int mrfa_occlusion_blend_subpixel(const float* a, const float* b2, const float* occ, float* y,
                                  int N, int C, int H, int W, int out_block, int64_t out_pixel_stride,
                                  mrfa_stream_t stream);

/* F.interpolate(x, size=(Ho,Wo), mode='bilinear', align_corners=True) raft.py:243 (and :205,228,
 * 266,...) fused with an optional activation (act as above).  SURVEY.md 8(f) N1: used as
 * relu(upsample(convc1(corr))) == relu(convc1(upsample(corr))) (raft.py:241-243, :61).
 * x (N,C,H,W) -> y (N,C,Ho,Wo); NCHW or NHWC memory (any C; vectorised when C % 4 == 0).
 * bias (C) or NULL: added after the interpolation and before the activation -- the bias of the 1x1 convolution that
 * was commuted below the resize (the bilinear weights sum to one, so resize(conv + b) == resize(conv) + b).      */
int mrfa_resize_bilinear(const float* x, const float* bias, float* y, int N, int C, int H, int W, int Ho, int Wo,
                         int channels_last, int act, mrfa_stream_t stream);

/* torch.cat([F.interpolate(o, size=(Ho,Wo), mode='bilinear', align_corners=True) for o in maps], dim=3) one map at a
 * time (raft.py:304-306, the occlusion strip returned next to the prediction): x (planes, H, W) planar -> columns
 * [y_col_offset, y_col_offset + Wo) of y (planes, Ho, y_row_pitch).  Wo, y_row_pitch, y_col_offset multiples of 4,
 * y 16-byte aligned.  act as above.                                                                                */
int mrfa_resize_bilinear_strip(const float* x, float* y, int64_t planes, int H, int W, int Ho, int Wo,
                               int64_t y_row_pitch, int64_t y_col_offset, int act, mrfa_stream_t stream);

/* Backward of mrfa_avg_pool2x2_nhwc: grad_y (N,C,H/2,W/2) -> grad_x (N,C,H,W), both NHWC; H, W even, C % 4 == 0. */
int mrfa_avg_pool2x2_nhwc_bwd(const float* grad_y, float* grad_x, int N, int C, int H, int W, mrfa_stream_t stream);

/* cat([a, b], dim=1) of two NHWC maps over `pixels` = N*H*W pixels (raft.py:66 cat([cor, flo]), :83
 * cat([motion_feature, context])): a (.., Ca), b (.., Cb) -> y (.., Ca+Cb); Ca % 4 == Cb % 4 == 0, 16-byte aligned. */
int mrfa_cat2_nhwc(const float* a, const float* b, float* y, int64_t pixels, int Ca, int Cb, mrfa_stream_t stream);

/* Hourglass decoder step `out = cat([up_block(out), skip], dim=1)` (util.py:239-263) with the up-block evaluated as
 * the sub-pixel 2x2 convolution (see mrfa_occlusion_blend_subpixel): b2 (N,4C,H+1,W+1) NHWC phase-major;
 * skip (N,Cs,2H,2W) with element strides {sn, sy, sx, sc}; y (N,C+Cs,2H,2W) NHWC = [shuffle(b2), skip].         */
int mrfa_subpixel_shuffle_cat(const float* b2, const float* skip, mrfa_grid_strides_t skip_strides, float* y,
                              int N, int C, int Cs, int H, int W, mrfa_stream_t stream);

/* 7x7 / stride 1 / pad 3 convolution with 2-3 input channels, + bias (BatchNorm pre-folded by the caller) and
 * optional ReLU, as a TF32 implicit GEMM on the tcgen05 tensor cores: BasicMotionEncoder.convf1 (2 -> 128,
 * raft.py:56,63) and the generator's `first` block (3 -> 64, generator.py:13).  Supported (Cin, Cout): (2,128), (3,64);
 * W % 128 == 0.  x (B,Cin,H,W) with element strides xs {sn, sy, sx, sc}; w_packed (Cout, KP) with
 * KP = mrfa_conv7x7_small_kpad(Cin), element [o][(ky*7+kx)*Cin + c] = weight[o][c][ky][kx], zero padded;
 * bias (Cout) or NULL; y (B,H,W,Cout) NHWC, 32-byte aligned.  sm_count sizes the persistent grid.            */
int mrfa_conv7x7_small_kpad(int Cin);
int mrfa_conv7x7_small(const float* x, mrfa_grid_strides_t xs, const float* w_packed, const float* bias, float* y,
                       int B, int Cin, int Cout, int H, int W, int relu, int sm_count, mrfa_stream_t stream);

/* Random affine + thin-plate warp of the identity grid -- the training-only equivariance warps.
 * metric MRFA_TPS_L1   : Transform.warp_coordinates model.py:50-70 (d = |dx|+|dy|, U = d^2 log(d + 1e-6));
 * metric MRFA_TPS_L2SQ : TPS mode 'random' util.py:412-423 (r2 = dx^2+dy^2, U = r2 log(r2 + 1e-9)).
 * theta (B,2,3); control_points (P,2) and control_params (B,P), or control_params NULL for the affine-only
 * Transform; grid (B,h,w,2) receives theta . p + sum_k U_k * params[b,k] (the same term on x and y).     */
enum { MRFA_TPS_L1 = 0, MRFA_TPS_L2SQ = 1 };
int mrfa_random_warp_grid(const float* theta, const float* control_points, const float* control_params, float* grid,
                          int B, int P, int h, int w, int metric, mrfa_stream_t stream);

/* Per-level update of the refinement loop (raft.py:256-262) as one pass:
 *   flow_w = flow + d_flow[:,0:2];  occ_new = occ + d_flow[:,2:3];  occ_sig = sigmoid(occ_new)
 * flow, flow_w (B,2,H,W) planar or (channels_last) pixel-interleaved; occ, occ_new, occ_sig (B,1,H,W) dense;
 * d_flow (B,>=3,H,W) with element strides d_strides {sn, sy, sx, sc}.                                            */
int mrfa_flow_update(const float* flow, const float* occ, const float* d_flow, mrfa_grid_strides_t d_strides,
                     float* flow_w, float* occ_new, float* occ_sig, int B, int H, int W, int channels_last,
                     mrfa_stream_t stream);

/* Flow / occlusion carry to the next, 2x finer refinement level (raft.py:276-295) as one pass:
 *   d_f = 2*up(d_flow[:,0:2]);  flow = d_f + up(init_flow)/scale;  d_o = up(d_flow[:,2]);  occ = d_o + up(prior_occ)
 *   if d_f_pre: up_f = 2*up(d_f_pre), up_o = up(d_occ_pre); flow += up_f; occ += up_o; d_f_acc = d_f + up_f; d_occ_acc = d_o + up_o
 *   else d_f_acc = d_f, d_occ_acc = d_o.            up(.) = bilinear, align_corners=True, to (2R, 2R).
 * d_flow (B,>=3,R,R) with element strides d_strides {sn, sy, sx, sc}; init_flow (B,2,h,h) and prior_occ (B,1,h,h)
 * NCHW-contiguous; d_f_pre (B,2,R,R) / d_occ_pre (B,1,R,R) or both NULL.  Outputs flow, d_f_acc (B,2,2R,2R) and
 * occ, d_occ_acc (B,1,2R,2R); `channels_last` gives the memory format of d_f_pre, flow and d_f_acc.            */
int mrfa_flow_carry(const float* d_flow, mrfa_grid_strides_t d_strides, const float* init_flow, const float* prior_occ,
                    const float* d_f_pre, const float* d_occ_pre, float* flow, float* occ, float* d_f_acc,
                    float* d_occ_acc, int B, int R, int h, float scale, int channels_last, mrfa_stream_t stream);

/* F.avg_pool2d(x, (2,2)) of DownBlock2d (util.py:190-196) in NHWC memory: x (N,C,H,W) -> y (N,C,H/2,W/2),
 * C % 4 == 0, 16-byte aligned.                                                               */
int mrfa_avg_pool2x2_nhwc(const float* x, float* y, int N, int C, int H, int W, mrfa_stream_t stream);

/* AntiAliasInterpolation2d.forward util.py:318-326 for scale = 1/stride:
 * F.pad(ka) -> depthwise KxK conv with weight (C,1,K,K) -> nearest sub-sampling, evaluated only at
 * the kept pixels.  in (N,C,H,W) NCHW -> out (N,C,H/stride,W/stride).  K odd (ka == kb).      */
int mrfa_antialias_down(const float* in, const float* weight, float* out, int N, int C, int H, int W, int K,
                        int ka, int stride, mrfa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MRFA_B200_H_ */
