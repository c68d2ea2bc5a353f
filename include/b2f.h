/* libb2f_cuda.so -- C ABI of the B200-native Back2Future hot path.
 *
 * Drop-in boundary for the reference's Torch7 nn.Module / nn.Criterion surface on this path
 * (SURVEY.md section 8b).  Plain C symbols, raw DEVICE pointers + explicit sizes/strides + an
 * explicit stream in, int status out.  No torch / THC / Lua types.  All tensors are fp32.
 *
 * Conventions
 *   - Every entry returns B2F_OK (0), a negative B2F_E* code for a rejected argument, or a
 *     positive cudaError_t.  Nothing throws, exits or longjmps across the ABI; the host shim
 *     turns a non-zero status into error()/an exception using b2f_last_error().
 *   - The caller owns every buffer.  The library owns only a small per-thread scratch for loss
 *     partials (released by b2f_release_scratch()) and a per-thread cache of TMA descriptors.
 *   - Re-entrant; the only mutable state is thread-local.  Work is enqueued on `stream` of the
 *     calling thread's current device and is asynchronous, except that criterion entries
 *     synchronise the stream when `loss_host` is non-NULL (the reference's criterions return
 *     a Lua number, i.e. they synchronise too).
 *   - Layouts follow the reference: feature maps and criterion tensors are BDHW (B,C,h,w)
 *     contiguous; sampler images are BHWD (B,H,W,C) contiguous; sampler grids are (B,Hg,Wg,2)
 *     with channel 0 = x offset, channel 1 = y offset, in pixels.
 *
 * Citations are relative to the reference repository root.
 */
#ifndef B2F_H_
#define B2F_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2F_ABI_VERSION 1

#if defined(__GNUC__)
#define B2F_API __attribute__((visibility("default")))
#else
#define B2F_API
#endif

typedef struct CUstream_st* b2f_stream_t; /* == cudaStream_t */

enum {
  B2F_OK = 0,
  B2F_EINVAL = -1,      /* bad size / flag / NULL pointer                                   */
  B2F_EUNSUPPORTED = -2, /* valid in the reference but not implemented (see message)         */
  B2F_ENOMEM = -3,      /* scratch allocation failed                                        */
  B2F_EALIGN = -4       /* pointer not 4-byte aligned                                       */
};

/* penalty functions, criterions/penalty/{quadratic,L1,Lorentzian}_function.lua */
enum {
  B2F_PENALTY_QUADRATIC = 0,  /* x^2, 2x                                                     */
  B2F_PENALTY_L1 = 1,         /* sqrt(x^2+1e-6), x/sqrt(x^2+1e-6)  (exponent is always 0.5)  */
  B2F_PENALTY_LORENTZIAN = 2  /* log(1+x^2/(2 eps^2)), 2x/(x^2+2 eps^2); eps = penalty_eps   */
};

/* ---- library ------------------------------------------------------------------------- */
B2F_API int b2f_abi_version(void);
B2F_API const char* b2f_last_error(void);          /* thread-local, never NULL                      */
B2F_API const char* b2f_status_string(int status);
B2F_API int b2f_release_scratch(void);             /* frees this thread's scratch on the current device */
/* Reserve `bytes` of zeroed loss scratch on the current device for criterion calls that are issued while their
 * stream is CAPTURING into a CUDA graph: every captured call takes its own slice (8 bytes per thread block + 16;
 * 64 KiB covers a 5-level training step), so a replay never shares a ticket counter with eager calls.  Call it
 * before the capture -- allocation is illegal during capture.  Released by b2f_release_scratch(); nothing the
 * library hands to a kernel is freed before that call, so a captured graph stays valid until then.        */
B2F_API int b2f_reserve_scratch(size_t bytes);
/* Test / measurement hook selecting the cost-volume kernel family (thread-local; returns the
 * previous mode): 0 = automatic (default), 1 = generic direct kernels, 2/3/4 = the tiled TMA
 * kernels whenever their preconditions hold (F=2, win=9, W%4==0, 16-byte aligned), ignoring the
 * grid-size heuristics, with 32- (2, 3) or 16-column (4) forward tiles; 5 = 16-column tiles with
 * the channel range split over at least two work items per tile (the small-level path);
 * 6 / 7 = force / forbid the persistent software-pipelined forward (FFMA2 + TMA-store epilogue).
 * 13 / 14 / 15 = backward with nine slab buffers and one CTA per SM / with the double-buffered ring
 * and two CTAs per SM / with 64-column tiles and a 3-deep ring.
 * 16 = the forward on the tensor cores (tcgen05 / TMEM, three-pass TF32 split; costvol_tc.cu) whenever F=2,
 * win=9, W%4==0 and the maps are 16-byte aligned; it is parity-green and measured level with the FFMA2
 * kernel, so the automatic mode does not select it.
 * 8, 9, 10 are measurement aids that produce WRONG results: the tiled kernels without their
 * arithmetic (8), without their stores (9), or with neither (10) -- they time the TMA feed; 18 / 19 are the
 * same for the tensor-core forward (no global stores / no TMEM loads).                              */
B2F_API int b2f_debug_costvol_path(int mode);
/* Stream-ordered zero-fill of a device buffer (what the sampler's Lua wrapper does with
 * gradInput:zero() before the native call, BilinearSamplerBHWD.lua:99-102).                  */
B2F_API int b2f_zero_async(void* ptr, size_t bytes, b2f_stream_t stream);
/* Stream-ordered device-to-device copy of `rows` rows of `row_elems` floats with independent row strides (in
 * elements): what nn.Narrow(2, a, 3) followed by a :contiguous() / nn.JoinTable(2) copy is in the reference's
 * graph (pwc.lua:141-146, 267) when a channel slice of one buffer has to become a dense tensor (or the reverse).  */
B2F_API int b2f_copy2d_async(float* dst, int64_t dst_row_stride, const float* src, int64_t src_row_stride,
                             int64_t row_elems, int64_t rows, b2f_stream_t stream);
/* Number of kernels (not memsets/copies) this thread has launched through the library since
 * the last reset; used by bench.py for its `gpu_launches` claim.                            */
B2F_API int64_t b2f_launch_count(int reset);

/* ---- nn.CostVolMulti  (models/CostVolMulti.lua) ----------------------------------------
 * frames[0] is the reference map, frames[1..F-1] the other maps; each (B,C,H,W) contiguous.
 * out (B, win*win, H, W): element stride between batch items is out_batch_stride (0 means
 * win*win*H*W), which lets both directions write straight into a 162-channel buffer.
 * Channel index is x-major: i = (qx+n)*win + (qy+n), n=(win-1)/2; source pixel is
 * p - s*(f)*q with s=+1 for fwd!=0, -1 for fwd==0; normaliser 1/(C*(F-1)) everywhere.
 * Replaces CostVolMulti:updateOutput (CostVolMulti.lua:49-109).  `frames` is a HOST array
 * of device pointers.  win must be odd.                                                   */
B2F_API int b2f_costvol_forward(const float* const* frames, int F, int B, int C, int H, int W,
                        int win, int fwd, float* out, int64_t out_batch_stride,
                        b2f_stream_t stream);

/* Replaces CostVolMulti:updateGradInput (CostVolMulti.lua:111-181).  gradOut is (B,win*win,H,W)
 * with batch stride gradOut_batch_stride elements (0 = contiguous): in the model it is a
 * narrow of the 162-channel JoinTable gradient (models/pwc.lua:267).  gradFrames[f] are
 * overwritten (the reference zero-fills then accumulates); gradFrames[f] may be NULL for
 * f >= 1 to skip that frame's gradient, gradFrames[0] may be NULL to skip the ref gradient. */
B2F_API int b2f_costvol_backward(const float* const* frames, int F, int B, int C, int H, int W,
                         int win, int fwd, const float* gradOut, int64_t gradOut_batch_stride,
                         float* const* gradFrames, b2f_stream_t stream);

/* ---- nn.BilinearSamplerBHWD  (extras/stnbhwd/BilinearSamplerBHWD.cu) --------------------
 * Replaces cunn_BilinearSamplerBHWD_updateOutput (.cu:118-158 -> kernel :41-115):
 * out[b,y,x,:] = bilinear(img[b], clamp(x+grid.x,0,W-1), clamp(y+grid.y,0,H-1)); a tap at
 * index W (or H) reads 0.  img (B,H,W,C), grid (B,Hg,Wg,2), out (B,Hg,Wg,C).               */
B2F_API int b2f_warp_bhwd_forward(const float* img, const float* grid, float* out,
                          int B, int H, int W, int C, int Hg, int Wg, b2f_stream_t stream);

/* Replaces cunn_BilinearSamplerBHWD_updateGradInput (.cu:313-365 -> kernel :161-307) and, with
 * gradImg == NULL, ..._updateGradInputOnlyGrid (.cu:368-419).  Exactly like the reference's
 * native entry, gradImg is ACCUMULATED into (atomic adds; the Lua wrapper zero-fills it first,
 * BilinearSamplerBHWD.lua:99-102) and gradGrid is overwritten.  No clamp derivative.        */
B2F_API int b2f_warp_bhwd_backward(const float* img, const float* grid, const float* gradOut,
                           float* gradImg, float* gradGrid,
                           int B, int H, int W, int C, int Hg, int Wg, b2f_stream_t stream);

/* ---- criterions: one fused pass each, loss + gradients ---------------------------------
 * Loss delivery (all criterion entries): the scalar is accumulated in double; it is written
 * to *loss_dev (device double, may be NULL) asynchronously, and if loss_host != NULL the
 * stream is synchronised and the value stored there.  Gradient outputs may be NULL to skip
 * them (forward only).  size_average mirrors the Lua field `sizeAverage`.                  */

typedef struct b2f_ob_params {
  int gradient_terms;     /* 0 = OBCC (OBCCriterion.lua), 1 = OBGCC (OBGCCriterion.lua)      */
  int penalty;            /* B2F_PENALTY_*  (field `p`)                                      */
  float penalty_eps;      /* Lorentzian eps (0.05 default)                                   */
  float penalty_out;      /* field `penalty_out`                                             */
  float alpha, beta, gamma; /* OBGCC weights; alpha is backward-only, as in the reference    */
  float pwc_flow_scaling; /* field `pwc_flow_scaling` (train.lua:425)                        */
  int past_flow;          /* field `past_flow`: past frame uses `bflow` for its mask         */
  int grad_check;         /* field `gradCheck`: non-zero skips the out-of-image mask         */
  int size_average;       /* field `sizeAverage`                                             */
} b2f_ob_params;

/* Occlusion-aware photometric criterion for F = 3 frames (two warped frames), the only
 * configuration the model produces; other F return B2F_EUNSUPPORTED.
 * flow,bflow,occ (B,2,h,w); warp_past, warp_future, target (B,C,h,w).  bflow may be NULL
 * unless past_flow.  Outputs: grad_occ (B,2,h,w), grad_warp_past/future (B,C,h,w).
 * Replaces OBCCriterion:updateOutput+updateGradInput (OBCCriterion.lua:36-240) and
 * OBGCCriterion (OBGCCriterion.lua:39-300).  No flow gradient exists in the reference.     */
B2F_API int b2f_ob_criterion(const b2f_ob_params* prm,
                     const float* flow, const float* bflow, const float* occ,
                     const float* warp_past, const float* warp_future, const float* target,
                     int B, int C, int h, int w,
                     float* grad_occ, float* grad_warp_past, float* grad_warp_future,
                     double* loss_dev, double* loss_host, b2f_stream_t stream);

typedef struct b2f_smooth_params {
  int order;              /* 1 = SmoothnessCriterion, 2 = SecondOrderSmoothnessCriterion      */
  int penalty;            /* B2F_PENALTY_*                                                    */
  float penalty_eps;
  float cs;               /* field `cs` (20)                                                  */
  int size_average;
  int alias_weights;      /* order 1 only: 1 = reproduce the Torch7 view-resize aliasing of the
                             edge weights (SmoothnessCriterion.lua:49-59; parity
                             default), 0 = the intended weights                              */
} b2f_smooth_params;

/* input (B,Cin,h,w) -- flow or occlusion map; target (B,Ct,h,w).  grad (B,Cin,h,w) or NULL.
 * Replaces SmoothnessCriterion.lua:28-106 / SecondOrderSmoothnessCriterion.lua:28-104.       */
B2F_API int b2f_smoothness_criterion(const b2f_smooth_params* prm, const float* input,
                             const float* target, int B, int Cin, int Ct, int h, int w,
                             float* grad, double* loss_dev, double* loss_host,
                             b2f_stream_t stream);

/* ConstVelCriterion.lua:29-74.  f,b (B,C,h,w) future / past flow; grads like inputs or NULL. */
B2F_API int b2f_constvel_criterion(const float* f, const float* b, int B, int C, int h, int w,
                           int size_average, float* grad_f, float* grad_b,
                           double* loss_dev, double* loss_host, b2f_stream_t stream);

/* OcclusionPriorCriterion.lua:28-73.  occ (B,C,h,w) with C = 2 or 3.                         */
B2F_API int b2f_occprior_criterion(const float* occ, int B, int C, int h, int w, float penalty,
                           int size_average, float* grad,
                           double* loss_dev, double* loss_host, b2f_stream_t stream);

/* ---- warpingUnit fused ------------------------------------------------
 * models/pwc.lua:68-73 `warpingUnit(I, F)` = Transpose(I), Transpose(F) -> nn.BilinearSamplerBHWD ->
 * Transpose back, fed by `nn.MulConstant(s)` on the flow (:402-408 feature warps, :441-446 image
 * warps).  These entries take the network's planar tensors directly -- img (B,C,H,W), flow (B,2,H,W)
 * with channel 0 = x, 1 = y in network units, flow_scale = the MulConstant factor -- and produce what
 * that chain produces: out (B,C,H,W); backward: gradImg (B,C,H,W) ACCUMULATED into a caller-zeroed
 * buffer (NULL = flow gradient only), gradFlow (B,2,H,W) = s * gradGrid.  Four layout passes and one
 * scaling pass of the reference disappear; results equal the chain's within fp32 rounding.       */
B2F_API int b2f_warp_bdhw_forward(const float* img, const float* flow, float flow_scale, float* out,
                                  int B, int C, int H, int W, b2f_stream_t stream);
B2F_API int b2f_warp_bdhw_backward(const float* img, const float* flow, float flow_scale,
                                   const float* gradOut, float* gradImg, float* gradFlow,
                                   int B, int C, int H, int W, b2f_stream_t stream);

/* ---- conv trunk of the PWC network ---------------------------------------
 * The reference builds its network from cudnn.SpatialConvolution(nIn, nOut, 3, 3, s, s, 1, 1) + nn.LeakyReLU(0.2)
 * (models/pwc.lua:58-65 convUnit, :76-85 decoder) wired by nngraph (:139-492).  These entries are the per-module
 * replacements; the graph itself is host code (back2future_b200/pwc.py mirrors createModelMulti, lua/models/pwc_b2f.lua
 * is the shim).  All tensors fp32 BDHW.
 *
 * Weights live in a PACKED layout [Cin * 9][CoutP], CoutP = Cout rounded up to 64, tap index ky * 3 + kx
 * (b2f_conv3x3_packed_floats floats); b2f_conv3x3_pack_weights converts Torch's (Cout, Cin, 3, 3) weight into it
 * (unpack != 0: the other way, e.g. to torch.save a checkpoint the reference can read).                       */
B2F_API int64_t b2f_conv3x3_packed_floats(int Cin, int Cout);
B2F_API int b2f_conv3x3_pack_weights(const float* w_torch, float* w_packed, int Cout, int Cin, int unpack,
                                     b2f_stream_t stream);
/* out = LeakyReLU_slope(conv3x3(x, w) + bias), zero padding 1, stride 1 or 2 (leaky_slope = 1: no activation).
 * x (B, Cin, H, W) and out (B, Cout, Ho, Wo) may be channel slices of wider buffers: pass the slice's base pointer
 * and the buffer's batch stride in elements (0 = contiguous).  out2 (may be NULL) receives a second copy with its
 * own batch stride -- the reference features of a level go to the next convUnit AND into the decoder's joined
 * input (nn.JoinTable of pwc.lua:298-305, 334 disappears).  bias may be NULL.
 * Replaces SpatialConvolution:updateOutput + LeakyReLU:updateOutput (pwc.lua:58-65, 76-85).                    */
B2F_API int b2f_conv3x3_forward(const float* x, int64_t x_batch_stride, const float* w_packed, const float* bias,
                                float* out, int64_t out_batch_stride, float* out2, int64_t out2_batch_stride,
                                int B, int Cin, int H, int W, int Cout, int stride, float leaky_slope,
                                b2f_stream_t stream);
/* nn.SpatialAveragePooling(2, 2, 2, 2) (pwc.lua:153): x (B, C, H, W) -> out (B, C, H/2, W/2).                  */
B2F_API int b2f_avgpool2x2_forward(const float* x, float* out, int B, int C, int H, int W, b2f_stream_t stream);
/* nn.SpatialUpSamplingBilinear(2) (pwc.lua:359-380; THNN's align-corners mapping src = dst * (in-1)/(out-1)):
 * x (B, C, H, W) with batch stride -> up to three destinations (B, C, 2H, 2W), each with its own batch stride
 * (0 = contiguous): the up-sampled flow feeds the next level's two decoders, its warps and the next up-sampling.
 * mul scales the result (1 = nn.SpatialUpSamplingBilinear alone; rescale_flow's MulConstant(2) otherwise).     */
B2F_API int b2f_upsample_bilinear2x_forward(const float* x, int64_t x_batch_stride, int B, int C, int H, int W,
                                            float* const* outs, const int64_t* out_batch_strides, int n_outs,
                                            float mul, b2f_stream_t stream);
/* nn.SpatialUpSamplingNearest(scale) (pwc.lua:308-316; two stacked x2 modules = scale 4).                     */
B2F_API int b2f_upsample_nearest_forward(const float* x, float* out, int B, int C, int H, int W, int scale,
                                         b2f_stream_t stream);
/* nn.SpatialSoftMax (pwc.lua:305): softmax over the channel dimension of (B, C, H, W).                         */
B2F_API int b2f_softmax_channels_forward(const float* x, float* out, int B, int C, int H, int W, b2f_stream_t stream);

/* ---- the decoders' convolutions on the tensor cores (tcgen05 / TMEM; inference) ---------------------------------
 * Same module as b2f_conv3x3_forward with stride 1 (SpatialConvolution + LeakyReLU, pwc.lua:76-85), as a TF32 implicit
 * GEMM with every operand split x = x_hi + x_lo (x_hi = x & 0xFFFFE000) and three MMA passes (hi*hi + lo*hi + hi*lo):
 * fp32-level accuracy (7e-6 relative measured, tools/ubench/tc_gemm.cu) where one TF32 pass gives 3e-3.
 * Activations between such layers are CHANNEL-MINOR (B, H, W, Cp), Cp = C rounded up to 32, as a (hi, lo) pair of
 * tensors; b2f_nhwc_split_from_bdhw converts a planar (B, C, H, W) tensor (batch-strided: the decoder's joined input).
 * Weights: [9 taps][Cout][Cin_p] hi / lo, b2f_conv3x3_tc_packed_floats(Cin, Cout) floats each.
 * Outputs: out_hi / out_lo (both or neither; (B, H, W, Cout rounded up to 32)) for the next tensor-core layer and / or
 * out_planar (B, Cout, H, W) fp32 with its batch stride (0 = dense) for every other consumer.  Any Cout: the
 * output columns run as slices of 32 .. 128 of ONE launch (wide layers; and coarse levels, where narrow slices shorten the
 * serial MMA chain of a launch that cannot fill the machine).                                                          */
B2F_API int64_t b2f_conv3x3_tc_packed_floats(int Cin, int Cout);
/* Measurement hook (thread-local): while a device buffer of 16 x (number of CTAs) uint64 is registered, every CTA of
 * b2f_conv3x3_tc_forward stores clock64 stamps of its phases there (tools/tc_trace.py); NULL switches it off.        */
B2F_API int b2f_debug_tc_trace(unsigned long long* device_buffer);
B2F_API int b2f_conv3x3_tc_pack_weights(const float* w_torch, float* w_hi, float* w_lo, int Cout, int Cin,
                                        b2f_stream_t stream);
/* The same (hi, lo) tensor-core weights from the PACKED layout of b2f_conv3x3_pack_weights, on the device (the training
 * path re-packs after every Adam step).  K = number of input channels the operand is padded to (>= Cin; the coarsest
 * flow decoder reads a wider joined input with zero weights).  transpose = 1 builds the operand of the INPUT-GRADIENT
 * convolution instead: N = Cin output rows, K (>= Cout) input channels, taps mirrored.                              */
B2F_API int b2f_conv3x3_tc_pack_from_packed(const float* w_packed, float* w_hi, float* w_lo, int Cout, int Cin, int K,
                                            int transpose, b2f_stream_t stream);
/* Many b2f_conv3x3_tc_pack_from_packed calls as ONE launch (a training step re-packs ~70 small weight tensors in the
 * forward plan and ~70 transposed ones in the backward plan: 0.4 ms of 6 us launches each).  `jobs_device` is an array
 * of njobs descriptors in DEVICE memory, valid until the launch has run.                                             */
typedef struct b2f_pack_job {
  const float* w_packed;
  float* w_hi;
  float* w_lo;
  int32_t Cout, Cin, K, transpose;
} b2f_pack_job;
B2F_API int b2f_conv3x3_tc_pack_from_packed_batch(const b2f_pack_job* jobs_device, int njobs, b2f_stream_t stream);
B2F_API int b2f_nhwc_split_from_bdhw(const float* x, int64_t x_batch_stride, float* hi, float* lo, int B, int C, int H,
                                     int W, b2f_stream_t stream);
B2F_API int b2f_conv3x3_tc_forward(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                                   const float* bias, float* out_hi, float* out_lo, float* out_planar,
                                   int64_t out_planar_batch_stride, int B, int Cin, int H, int W, int Cout,
                                   float leaky_slope, b2f_stream_t stream);
/* SpatialConvolution:updateGradInput of the same layer on the tensor cores (train.lua:480 model:backward): the forward
 * kernel on the transposed, mirrored weights (b2f_conv3x3_tc_pack_from_packed with transpose = 1: N = Cin rows, K = Cout)
 * over the channel-minor (hi, lo) gradient of the layer's OUTPUT; `act` = planar forward output of the layer BELOW
 * (B, Cin, H, W): the result is multiplied by 1 where act > 0 and by leaky_slope elsewhere (NULL: no factor).  `act_hi`
 * is the alternative to `act` (at most one of the two): the HI half of the same activation's channel-minor split,
 * (B, H, W, Cin rounded up to 32) -- hi = x & 0xFFFFE000 has the sign of every normal float, and a pixel's channels
 * are one 128-byte line per 32 instead of 32 strided words.  Outputs
 * as b2f_conv3x3_tc_forward: (hi, lo) channel-minor for the next tensor-core input gradient and / or planar fp32 for
 * the weight-gradient kernel.  Any Cin (the first decoder layer: 162 .. 356): slices of one launch, as in the
 * forward.  accumulate != 0: the planar result is ADDED to
 * gin_planar (nngraph's gradient accumulation at the joined decoder input).                                            */
B2F_API int b2f_conv3x3_tc_backward_data(const float* g_hi, const float* g_lo, const float* wt_hi, const float* wt_lo,
                                         const float* act, int64_t act_batch_stride, const float* act_hi, float* gin_hi,
                                         float* gin_lo, float* gin_planar, int64_t gin_planar_batch_stride, int B, int Cout,
                                         int H, int W,
                                         int Cin, float leaky_slope, int accumulate, b2f_stream_t stream);
/* SpatialConvolution:updateGradInput of a STRIDE-2 3x3 convolution (the down-sampling half of pwc.lua:58-65's convUnit)
 * on the tensor cores: g_hi / g_lo = channel-minor (B, Ho, Wo, Cout rounded up to 32) split of the output gradient, wt_hi
 * / wt_lo as for b2f_conv3x3_tc_backward_data; gin_planar (B, Cin, H, W) with Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1.
 * gin[2y + py, 2x + px] only receives the taps with ky = py + 1, kx = px + 1 (mod 2): four accumulators (one per parity
 * class of the output pixel) over ONE low-resolution patch, a quarter of the multiply-adds of the zero-inserted form.
 * Cin <= 128.  accumulate != 0: added to gin_planar.                                                                */
B2F_API int b2f_conv3x3_tc_backward_data_s2(const float* g_hi, const float* g_lo, const float* wt_hi, const float* wt_lo,
                                            float* gin_planar, int64_t gin_planar_batch_stride, int B, int Cout, int Ho,
                                            int Wo, int Cin, int H, int W, int accumulate, b2f_stream_t stream);
/* SpatialConvolution:accGradParameters of the same layer on the tensor cores: gw_packed += d loss / d weight from the
 * channel-minor (hi, lo) INPUT activation (x_hi / x_lo, a (B, H, W, Cx rounded up to 32) tensor whose first Cin channels
 * are this convolution's input -- the coarsest flow decoder reads the first 162 channels of a wider joined input) and
 * the channel-minor (hi, lo) OUTPUT gradient: MN-major operands, the contraction runs over the pixels and a tap is a
 * descriptor offset (wgrad_tc.cu).  gbias (may be NULL) += d loss / d bias, summed from the PLANAR output gradient
 * g_planar (B, Cout, H, W; batch stride 0 = dense) or, when g_planar is NULL, from g_hi + g_lo.  Any Cin, any Cout
 * (output-channel slices of <= 128).                                                                                   */
B2F_API int b2f_conv3x3_tc_backward_weights(const float* x_hi, const float* x_lo, int Cx, const float* g_hi, const float* g_lo,
                                            const float* g_planar, int64_t g_planar_batch_stride, float* gw_packed,
                                            float* gbias, int B, int Cin, int H, int W, int Cout, b2f_stream_t stream);

/* ---- training: backward of the conv trunk + optimizer -------------------------------
 * Weight gradients live in the same PACKED layout as the weights ([Cin * 9][CoutP]), so that parameters, gradients
 * and the Adam moments of the whole network are one flat buffer each (what getParameters() gives the reference,
 * train.lua:24) and the gradient all-reduce (b2f_comm.h) is one call; padding entries stay zero.                 */
/* wt[(co * 9 + 8 - tap)][CinP] <- w_packed: the weights backward-data reads (CinP = Cin rounded up to 64;
 * b2f_conv3x3_packed_floats(Cout, Cin) floats).                                                                  */
B2F_API int b2f_conv3x3_transpose_packed(const float* w_packed, float* wt_packed, int Cout, int Cin, b2f_stream_t stream);
/* SpatialConvolution:updateGradInput fused with the LeakyReLU:updateGradInput of the layer BELOW: gin (B, Cin, H, W)
 * [+]= conv^T(gout (B, Cout, Ho, Wo)) * (act > 0 ? 1 : leaky_slope), act = that layer's forward output (NULL: no
 * factor).  accumulate != 0 adds to gin (fan-out nodes of the graph).  Sizes are the FORWARD convolution's.      */
B2F_API int b2f_conv3x3_backward_data(const float* gout, int64_t gout_batch_stride, const float* wt_packed,
                                      const float* act, int64_t act_batch_stride, float* gin, int64_t gin_batch_stride,
                                      int accumulate, int B, int Cin, int H, int W, int Cout, int stride,
                                      float leaky_slope, b2f_stream_t stream);
/* SpatialConvolution:accGradParameters: gw_packed += d loss / d weight, gbias (may be NULL) += d loss / d bias.    */
B2F_API int b2f_conv3x3_backward_weights(const float* x, int64_t x_batch_stride, const float* gout,
                                         int64_t gout_batch_stride, float* gw_packed, float* gbias, int B, int Cin,
                                         int H, int W, int Cout, int stride, b2f_stream_t stream);
/* out (planes, 2H, 2W) <- x (planes, H, W) at the even coordinates, zeros elsewhere: the dilated output gradient with
 * which a stride-2 convolution's input gradient is b2f_conv3x3_backward_data(..., stride = 1) on the 2H x 2W grid
 * (the fast path; the direct stride-2 form inside b2f_conv3x3_backward_data is the fallback for odd widths).       */
B2F_API int b2f_zero_insert2x(const float* x, float* out, int64_t planes, int H, int W, b2f_stream_t stream);
/* LeakyReLU:updateGradInput in place on `rows` rows of `row_elems` floats: grad *= (act > 0 ? 1 : slope).         */
B2F_API int b2f_leaky_relu_backward(float* grad, int64_t grad_row_stride, const float* act, int64_t act_row_stride,
                                    int64_t row_elems, int64_t rows, float slope, b2f_stream_t stream);
/* dst += alpha * src on rows with independent strides: nngraph's gradient accumulation at fan-out nodes and the
 * level_weights * opt.* scaling of the criterion gradients (train.lua:421-468).                                  */
B2F_API int b2f_axpy2d(float* dst, int64_t dst_row_stride, const float* src, int64_t src_row_stride, int64_t row_elems,
                       int64_t rows, float alpha, b2f_stream_t stream);
/* :updateGradInput of SpatialUpSamplingBilinear(2) (sizes are the forward INPUT's; the result is multiplied by mul and
 * optionally added to grad_in), SpatialUpSamplingNearest(scale), SpatialSoftMax (softmax_out = the forward result). */
B2F_API int b2f_upsample_bilinear2x_backward(const float* grad_out, float* grad_in, int B, int C, int H, int W, float mul,
                                             int accumulate, b2f_stream_t stream);
B2F_API int b2f_upsample_nearest_backward(const float* grad_out, float* grad_in, int B, int C, int H, int W, int scale,
                                          b2f_stream_t stream);
B2F_API int b2f_softmax_channels_backward(const float* softmax_out, const float* grad_out, float* grad_in, int B, int C,
                                          int H, int W, b2f_stream_t stream);
/* optim.adam (train.lua:485-486) on flat buffers; t = 1, 2, ... is the step counter (bias correction).            */
B2F_API int b2f_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int64_t t, b2f_stream_t stream);

/* ---- Middlebury .flo files -- HOST buffers, no device work ----------------
 * File layout (flowExtensions.lua:254-287): float32 tag 202021.25 ("PIEH"), int32 width, int32
 * height, then height*width interleaved (u, v) float32 pairs, all little-endian.  The reference
 * keeps flow as a (2, h, w) tensor and permutes on the way in and out (loadFLO :268-269, writeFLO
 * :275); these entries take / return that planar (2, h, w) layout.                              */
/* flowExtensions.lua:274-286 writeFLO(filename, F).                                              */
B2F_API int b2f_flo_write(const char* path, const float* flow_chw, int h, int w);
/* Reads the header only (tag check, loadFLO :256-264): *w, *h.                                   */
B2F_API int b2f_flo_read_header(const char* path, int* w, int* h);
/* flowExtensions.lua:254-271 loadFLO(filename) into a caller-owned (2, h, w) buffer whose h, w
 * must equal the header's.                                                                       */
B2F_API int b2f_flo_read(const char* path, float* flow_chw, int h, int w);

#ifdef __cplusplus
}
#endif
#endif /* B2F_H_ */
