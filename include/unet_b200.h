/*
 * unet_b200.h - C ABI of libunet_b200.so: the B200 (sm_100a) replacement for the U-Net hot path of
 * masktrump19-sudo/unet-lane-detection.
 *
 * Reference interfaces replaced (paths relative to the reference tree):
 *   README.md:1421-1481   class UNet(nn.Module): __init__/_conv_block/forward  -> plan_* + forward
 *   src/unet.py:24-42     RKNNLaneInference.preprocess_image (+ README.md:3110-3111 mean/std)
 *                                                                               -> preprocess_u8
 *   src/unet.py:44-72     RKNNLaneInference.postprocess_output (sigmoid, > thr, *255) -> forward(mask)
 *   src/py_utils/rknn_executor.py:26-38  RKNN_model_container.run               -> infer_u8_host
 *
 * Conventions: every function returns 0 on success and a negative code on failure; the message is
 * available from unet_b200_last_error() (thread-local). Nothing here throws or calls exit().
 * All `*_dev` pointers are caller-owned CUDA device memory; `stream` is a cudaStream_t passed as
 * void* (NULL = default stream); work is enqueued, not synchronised, unless stated otherwise.
 * A plan is not thread-safe; distinct plans are independent (each carries the switches it was created with, and everything
 * the library remembers about a GPU is kept per device ordinal, so one process may drive plans on several devices).
 * Every call works on the CUDA device that is current for the calling thread. The library targets sm_100a only:
 * on any other device the calls fail with UB_ERR_DEVICE - there is no CPU or library fallback.
 */
#ifndef UNET_B200_H
#define UNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB_OK 0
#define UB_ERR_ARG (-1)     /* bad argument / unsupported shape */
#define UB_ERR_CUDA (-2)    /* a CUDA runtime or driver call failed */
#define UB_ERR_DEVICE (-3)  /* not an sm_100 device */
#define UB_ERR_STATE (-4)   /* plan not bound / weights not set */

#define UB_MAX_LEVELS 6

typedef struct unet_b200_plan unet_b200_plan;

const char* unet_b200_last_error(void);
int unet_b200_version(void);
/* 0 if the current device can run the kernels (compute capability 10.x), UB_ERR_DEVICE otherwise. */
int unet_b200_device_ok(void);

/* ---- plan: UNet(in_channels, out_channels, features) at a fixed HxW and batch capacity ---------- *
 * Mirrors UNet.__init__ (README.md:1424-1447). H and W must be divisible by 2^levels; features are any positive widths
 * (widths that are not multiples of 64 - the deployed topology [32,64,128], or 48 / 100 / 264 - are stored zero-extended to
 * the next multiple of 64; results are unchanged; features[0] <= 256; the fp32-class plan and the trainer keep their own
 * limits, see there), in_channels <= 4, out_channels in [1,64] (the reference trains and
 * deploys out_channels = 1, README.md:2165; with more the 1x1 head runs as its own kernel instead of inside the last conv). */
int unet_b200_plan_create(unet_b200_plan** out, int max_batch, int H, int W, int in_channels, int out_channels,
                          const int* features, int levels);
/* The same with an explicit arithmetic class (plan_create = UB_PRECISION_BF16):
 *   UB_PRECISION_BF16  bf16 operands, fp32 accumulation: logits within 2e-2 of the fp32 reference (BASELINE.json north_star);
 *   UB_PRECISION_FP32  "fp32-class": every activation and BN-folded weight is carried as two bf16 numbers hi + lo (16 mantissa
 *                      bits) and a product is hi*hi + lo*hi + hi*lo on the same tcgen05 kernel - logits within 1e-4 (north_star's
 *                      second gate), about 3x the tensor work. Needs features % 64 == 0 and out_channels == 1. The network
 *                      input of such a plan is FP32 NHWC4 (unet_b200_nchw_to_nhwc4_f32 / unet_b200_preprocess_u8_f32), and
 *                      the host-buffer entry points (infer_u8_host*) pick the fp32 preprocess themselves. */
#define UB_PRECISION_BF16 0
#define UB_PRECISION_FP32 1
int unet_b200_plan_create_ex(unet_b200_plan** out, int max_batch, int H, int W, int in_channels, int out_channels,
                             const int* features, int levels, int precision);
int unet_b200_plan_precision(const unet_b200_plan* p);
void unet_b200_plan_destroy(unet_b200_plan* p);
/* Bytes of device memory the plan needs for activations (workspace) and packed weights. Layer outputs share the workspace
 * by liveness (an output's space is reused after its last reader): the default network needs 19.3 MB per frame of batch
 * capacity; plan_workspace_unshared_bytes reports what one private buffer per layer output would take (64.1 MB). */
size_t unet_b200_plan_workspace_bytes(const unet_b200_plan* p);
size_t unet_b200_plan_workspace_unshared_bytes(const unet_b200_plan* p);
size_t unet_b200_plan_weight_bytes(const unet_b200_plan* p);
/* Hand the plan its two caller-owned device buffers (256-byte aligned); builds the TMA tensor maps. */
int unet_b200_plan_bind(unet_b200_plan* p, void* workspace_dev, void* weights_dev);

/* Number of 3x3 convolutions (2*(2*levels+1)) and their order:
 * enc0.0, enc0.3, enc1.0, ..., bottleneck.0, bottleneck.3, dec0.0, dec0.3, ...  (state_dict order
 * of README.md:1432-1444 with decoder blocks taken at odd indices). */
int unet_b200_plan_num_convs(const unet_b200_plan* p);
/* Fold eval-mode BatchNorm into conv `idx` and pack it into the plan's weight buffer.
 * w: fp32 [Cout][Cin][3][3]; gamma/beta/mean/var: fp32 [Cout] (all four NULL = no BN). */
int unet_b200_plan_set_conv(unet_b200_plan* p, int idx, const float* w_dev, const float* gamma_dev,
                            const float* beta_dev, const float* mean_dev, const float* var_dev, float eps,
                            void* stream);
/* ConvTranspose2d(2f, f, 2, 2) of decoder level idx (0 = deepest): w fp32 [2f][f][2][2], bias fp32 [f]. */
int unet_b200_plan_set_convT(unet_b200_plan* p, int idx, const float* w_dev, const float* bias_dev, void* stream);
/* Output Conv2d(f0, out_channels, 1): w fp32 [out_channels][f0], bias fp32 [out_channels] (read synchronously on `stream`). */
int unet_b200_plan_set_head(unet_b200_plan* p, const float* w_dev, const float* bias_dev, void* stream);

/* ---- forward (UNet.forward, README.md:1460-1481, eval mode) -------------------------------------- *
 * x_nhwc4_dev: bf16 [batch][H][W][4] (channels >= in_channels are ignored/zero); fp32 [batch][H][W][4] for a
 * UB_PRECISION_FP32 plan.
 * Any of the three outputs may be NULL (OC = out_channels):
 *   logits_dev fp32 [batch][OC][H][W]   (NCHW, as the reference module returns them)
 *   probs_dev  fp32 [batch][OC][H][W]   sigmoid(logits)
 *   mask_dev   u8   [batch][OC][H][W]   (sigmoid(logit) > threshold) ? 255 : 0   (src/unet.py:63-67) */
int unet_b200_forward(unet_b200_plan* p, const void* x_nhwc4_dev, int batch, float* logits_dev, float* probs_dev,
                      uint8_t* mask_dev, float threshold, void* stream);
/* Number of kernels one unet_b200_forward call launches (for launch accounting). */
int unet_b200_forward_launches(const unet_b200_plan* p);
/* Same as unet_b200_forward, but brackets every kernel with CUDA events on `stream`, synchronises,
 * and returns the per-kernel durations in ms_out[0 .. plan_num_layers) (last entry = head kernel). */
int unet_b200_forward_profile(unet_b200_plan* p, const void* x_nhwc4_dev, int batch, float* logits_dev,
                              float* probs_dev, uint8_t* mask_dev, float threshold, void* stream, float* ms_out,
                              int n_out);
int unet_b200_plan_num_layers(const unet_b200_plan* p);
/* info8 = {kind (0 stem, 1 conv3x3, 2 convT2x2, 3 head), H, W, Cin, Cout, taps, BLOCK_N,
 * flags (1 fused pool | 2 halo kernel | 4 fused head)} of kernel idx;
 * H, W are the GEMM-row grid (input resolution). */
int unet_b200_plan_layer_info(const unet_b200_plan* p, int idx, int* info8);

/* Kernel-selection switches. unet_b200_set_option changes the PROCESS DEFAULTS; a plan / trainer copies the defaults when it is
 * created and keeps its copy for life (changing a default later does not touch existing plans), the single-layer entry points
 * read the defaults when they are called:
 *   "halo" (default 1)      3x3 convs with Cout 64/128 and W % 8 == 0 run on the halo-patch kernel
 *   "fuse_head" (default 1) the 1x1 head + sigmoid + mask run in the last conv's epilogue when it is a halo layer
 *   "stem_umma" (default 1) the Cin<=4 -> 64 stem runs on tensor cores (in-kernel im2col) instead of the FP32 pipes
 *   "halo2" / "umma2" (default 1) the conv kernels run as CTA pairs (tcgen05 cta_group::2, M = 256 per MMA, half of the weight
 *                           tile per SM) whenever a layer has at least two pixel tiles; 0 = one CTA per tile. Bit-identical results.
 *   "wgrad_rows64" (default 1) 64-pixel reduction tiles in the BLOCK_N = 256 weight-gradient kernel
 *   "wgrad2" (default 1)    CTA-pair weight-gradient kernel for Cout >= 128 (read when a trainer / wgrad call is set up)
 *   "wgrad_stream" (default 1) backward: weight-gradient GEMMs on a side stream, overlapping the elementwise backward passes
 *   "pdl" (default -1)      programmatic dependent launch for every kernel: the conv / stem kernels run their prologue before
 *                           griddepcontrol.wait, so it overlaps the predecessor's tail. -1 = per plan: on for plans of up to
 *                           600 k pixels (one 224 x 224 frame: 0.319 -> 0.262 ms per pass), off above (2-3 % slower from 32
 *                           frames on; trainers and the single-layer entry points: off); 0 / 1 force it
 *   "bwd_fuse" (default 2)  training: the BatchNorm-backward reduction of a layer runs inside the pass that produces its
 *                           incoming gradient where that pass is an elementwise kernel: 2 = the head backward only
 *                           (measured 16.81-16.99 against 17.00-17.24 ms/step), 1 = also the four max-pool backward passes
 *                           (measured slower: those kernels then need 140 registers), 0 = off
 *   "dgrad_fuse" (default 1) training: where a layer's incoming gradient is written by a tcgen05 dgrad kernel (13 of the 18
 *                           BatchNorm layers of the default network), that kernel's epilogue also takes the layer's
 *                           BatchNorm-backward sums (the y sub-box is TMA-loaded beside the staged output tile), so the
 *                           separate reduction pass over g and y does not run
 *   "wgrad_halo" (default 1) training: weight gradients of the 3x3 convs with Cout == 64 (the full-resolution level) run on
 *                           wgrad_halo_kernel - one halo'd activation patch serves all nine taps - instead of one CTA per tap pair
 *   "host_pieces" (default 8) unet_b200_infer_u8_host_stream: pieces per pass (see there); 0 or 1 = pass-granular pipeline
 *   "small_n" (default 1)   small plans (the per-frame executor path: a 14 x 14 level at batch 1 is 16 tiles of 128 x 256
 *                           on 148 SMs) run their per-tap layers with column blocks of 128 or 64 where a wave model
 *                           (rounds of the persistent grid x measured tile cost 1 : 0.56 : 0.375) says that is faster;
 *                           2 = narrower blocks only while a layer cannot give every SM a tile (the first rule: 5-8 %
 *                           slower at batches 2-8), 0 = always 256; bit-identical logits. Pass of one frame 0.38 -> 0.27 ms
 *   "host_geometric" (default 1) unet_b200_infer_u8_host_stream: the pieces of a pass are 16, 32, 64, ... frames (the copy of a
 *                           piece hides behind the front layers of the piece half its size before it) and the last layer
 *                           runs them largest first: 16 frames of input / output copy exposed and 4 instead of 8 launches per
 *                           piece-wise layer; 0 = "host_pieces" equal pieces. Measured e2e 16.64-16.70 k -> 16.79-16.83 k frames/s
 *   "host_hybrid" (default 1) unet_b200_infer_u8_host_stream with source frames more than 1.5x the network input (their
 *                           copies take longer than the layers a piece hides them behind): short first pass AND pieces
 *                           inside every pass; 0 = pass-granular pipeline without pieces for such frames
 *   "pre_bulk" (default 1)  resizing preprocess: the source rows of a tile are staged by the copy engine (cp.async.bulk
 *                           into two shared-memory stages, preprocess_bulk_u8_kernel); 0 = the thread-staged tile kernel.
 *                           Bit-identical results. "pre_rows" / "pre_stages" (default 0 = 8 rows per tile, halved until the
 *                           stages fit 72 KB / 2 stages) override the tile shape for A/B runs.
 *   "stem_fuse" (default 0) inference plans: the stem is computed inside the patch producer of the first block's second
 *                           conv (stem_halo2_kernel: a small tensor-core GEMM per tile fills the halo'd shared-memory patch), so
 *                           the stem's 64-channel output is never written to or read from HBM; bit-identical results.
 *                           Measured on B200: 1.46-1.58 ms against 0.61 + 0.82 ms for the two kernels (256 frames) - the fused
 *                           kernel is bound by shared-memory bandwidth - so it is off
 *   "stem_wide" (default 0) tensor-core stem on 4 x 32 pixel tiles (one contiguous 4 KB output row per TMA store) instead
 *                           of 16 x 8; bit-identical, measured +0.3 % (noise): the stem is bound by the write rate, not by
 *                           the store pattern
 * All paths are hand-written sm_100a kernels; the switches exist for A/B measurement and tests. */
int unet_b200_set_option(const char* name, int value);

/* NCHW fp32 [batch][C<=4][H][W] -> NHWC4 bf16 (the nn.Module boundary); _f32: -> NHWC4 fp32 (fp32-class plans). */
int unet_b200_nchw_to_nhwc4(const float* x_dev, int batch, int C, int H, int W, void* y_nhwc4_dev, void* stream);
int unet_b200_nchw_to_nhwc4_f32(const float* x_dev, int batch, int C, int H, int W, float* y_nhwc4_dev, void* stream);

/* ---- preprocess (src/unet.py:24-42 + in-graph normalisation README.md:3110-3111) ---------------- *
 * src_dev: uint8 [batch][Hs][Ws][3] with the given row pitch / frame stride in bytes.
 * cv2.resize(INTER_LINEAR)-exact bilinear resize to HxW, optional R<->B swap (BGR input), then
 * (x - mean[c]) / std[c] with mean/std given in output (RGB) channel order, written as NHWC4 bf16.
 * resized_u8_dev (optional, may be NULL) receives the resized uint8 [batch][H][W][3] frame. */
int unet_b200_preprocess_u8(const uint8_t* src_dev, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride,
                            int H, int W, int swap_rb, const float* mean3, const float* std3, void* y_nhwc4_dev,
                            uint8_t* resized_u8_dev, void* stream);
/* Same, written as NHWC4 fp32 (no bf16 rounding of the normalised image): the input of a UB_PRECISION_FP32 plan. */
int unet_b200_preprocess_u8_f32(const uint8_t* src_dev, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride,
                                int H, int W, int swap_rb, const float* mean3, const float* std3, float* y_nhwc4_dev,
                                uint8_t* resized_u8_dev, void* stream);

/* ---- IPM front end of the ROS node fused into the preprocess (src/unet_ros_node.py:297-313) --------------------- *
 * cv2.warpPerspective(bgr, M, (Ww, Hw)) -> BGR2RGB (swap_rb) -> cv2.resize to HxW -> normalise -> NHWC4 bf16, bit-exact
 * with cv2's uint8 arithmetic; the Hw x Ww bird's-eye image is not materialised unless warped_u8_dev is given
 * ([batch][Hw][Ww][3], source channel order). m_inv9: HOST, row-major 3x3 double INVERSE map (dst -> src), i.e.
 * cv::invert of the matrix the reference passes to warpPerspective. y_nhwc4_dev may be NULL when only the warp is wanted. */
int unet_b200_preprocess_warp_u8(const uint8_t* src_dev, int batch, int Hs, int Ws, size_t pitch, size_t frame_stride,
                                 const double* m_inv9, int Hw, int Ww, int H, int W, int swap_rb, const float* mean3,
                                 const float* std3, void* y_nhwc4_dev, uint8_t* resized_u8_dev, uint8_t* warped_u8_dev,
                                 void* stream);
/* Mask back to the caller's resolution (src/unet.py:70): dst [batch][Hd][Wd] = cv2.resize(src [batch][Hs][Ws]) for
 * single-channel uint8, INTER_LINEAR, bit-exact (including cv2's border-row rule and its 2x2-decimation special case). */
int unet_b200_resize_gray_u8(const uint8_t* src_dev, int batch, int Hs, int Ws, uint8_t* dst_dev, int Hd, int Wd,
                             void* stream);

/* ---- executor entry point (RKNN_model_container.run, src/py_utils/rknn_executor.py:26-38) -------- *
 * HOST uint8 [batch][Hs][Ws][3] frames in, HOST outputs out (any may be NULL). Copies H2D, runs
 * preprocess + forward on `stream`, copies D2H and synchronises the stream before returning.
 * staging_dev must hold unet_b200_infer_staging_bytes(...) bytes. */
size_t unet_b200_infer_staging_bytes(const unet_b200_plan* p, int Hs, int Ws);
int unet_b200_infer_u8_host(unet_b200_plan* p, void* staging_dev, const uint8_t* frames_host, int batch, int Hs,
                            int Ws, int swap_rb, const float* mean3, const float* std3, float threshold,
                            float* logits_host, float* probs_host, uint8_t* mask_host, void* stream);

/* Same contract for ANY number of frames: the batch is cut into passes of the plan's capacity; the H2D copies of pass
 * i+1 and the D2H copies of pass i-1 overlap the kernels of pass i (two staging slots, two internal copy streams created
 * on first use). Inside a pass the input copy, the preprocess and the two full-resolution layers at the start of the
 * network, and the fused-head conv and the output copies at its end, run per PIECE of the pass (pieces of 16, 32, 64, ...
 * frames, options "host_geometric" / "host_pieces"; bf16 plans whose stem / first conv / last conv run on the tensor-core
 * stem and halo kernels), so only
 * the first piece's input copy and the last piece's output copy are not hidden behind kernels. Source frames more than
 * 1.5x the size of the network input (whose copies take longer than those two layers) additionally get a SHORT first
 * pass (a quarter of the capacity) whose kernels cover the copies of the rest (option "host_hybrid"); plans without
 * pieces use a pass-granular pipeline with a short first pass. Host buffers should be pinned.
 * staging_dev: unet_b200_infer_stream_staging_bytes(...) bytes.
 * plan_host_pieces: pieces a full pass is cut into (0: pass-granular fallback); infer_stream_launches: kernels one call
 * for `total` frames launches. */
size_t unet_b200_infer_stream_staging_bytes(const unet_b200_plan* p, int Hs, int Ws);
int unet_b200_plan_host_pieces(const unet_b200_plan* p);
int unet_b200_infer_stream_launches(const unet_b200_plan* p, int total, int Hs, int Ws);
int unet_b200_infer_u8_host_stream(unet_b200_plan* p, void* staging_dev, const uint8_t* frames_host, int total, int Hs,
                                   int Ws, int swap_rb, const float* mean3, const float* std3, float threshold,
                                   float* logits_host, float* probs_host, uint8_t* mask_host, void* stream);

/* ---- single layers (parity tests and reuse outside a plan) --------------------------------------- *
 * conv3x3: y = relu?(conv3x3(cat(x0, x1)) + bias); x0/x1 bf16 NHWC [B,H,W,C0|C1] (C1 may be 0, then
 * x1 is ignored); wp bf16 [Cout][9][C0+C1]; bias fp32 [Cout]; y bf16 [B,H,W,Cout]; pool (optional)
 * bf16 [B,H/2,W/2,Cout] = maxpool2x2(y). C0, C1, Cout multiples of 64. */
int unet_b200_conv3x3(const void* x0_dev, int C0, const void* x1_dev, int C1, const void* wp_dev,
                      const float* bias_dev, int B, int H, int W, int Cout, int relu, void* y_dev, void* pool_dev,
                      void* stream);
/* convT2x2: y[b,2h+dy,2w+dx,co] = bias[co] + sum_ci x[b,h,w,ci]*w[ci,co,dy,dx]; wp bf16 [4f][Cin]. */
int unet_b200_convT2x2(const void* x_dev, int Cin, const void* wp_dev, const float* bias_dev, int B, int H, int W,
                       int f, void* y_dev, void* stream);
/* Packing helpers producing the layouts above from PyTorch-layout fp32 tensors. */
int unet_b200_pack_conv3x3(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* mean_dev,
                           const float* var_dev, float eps, int Cout, int Cin, void* wp_dev, float* bias_dev,
                           void* stream);
int unet_b200_pack_convT2x2(const float* w_dev, int Cin, int f, void* wp_dev, void* stream);
/* Stem conv (Cin <= 4) on NHWC4 input; ws fp32 [9][4][Cout] from pack_stem. */
int unet_b200_pack_stem(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* mean_dev,
                        const float* var_dev, float eps, int Cout, int Cin, float* ws_dev, float* bias_dev,
                        void* stream);
int unet_b200_stem_conv(const void* x_nhwc4_dev, const float* ws_dev, const float* bias_dev, int B, int H, int W,
                        int Cin, int Cout, int relu, void* y_dev, void* stream);
/* Tensor-core stem (Cout == 64): wp bf16 [64][64] with k = tap*4 + ci from pack_stem_tc. */
int unet_b200_pack_stem_tc(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* mean_dev,
                           const float* var_dev, float eps, int Cout, int Cin, void* wp_dev, float* bias_dev,
                           void* stream);
int unet_b200_stem_conv_tc(const void* x_nhwc4_dev, const void* wp_dev, const float* bias_dev, int B, int H, int W,
                           int relu, void* y_dev, void* stream);
/* 1x1 head on bf16 NHWC [npix][C]: w fp32 [C] (device), bias by value. Outputs optional. */
int unet_b200_head(const void* x_dev, const float* w_dev, float bias, size_t npix, int C, float* logits_dev,
                   float* probs_dev, uint8_t* mask_dev, float threshold, void* stream);
int unet_b200_maxpool2x2(const void* x_dev, int B, int H, int W, int C, void* y_dev, void* stream);

/* ---- training step (README.md:2060-2084 train_one_epoch, model.train(): BatchNorm uses batch statistics) --------- *
 * A trainer is a fixed-batch plan that keeps every layer's activations for the backward pass.
 * Parameters and gradients are FLAT fp32 device arrays in model.parameters() order (registration order of
 * README.md:1427-1447: encoder_blocks.i.{0.weight,1.weight,1.bias,3.weight,4.weight,4.bias}, decoder_blocks.2j.{weight,
 * bias}, decoder_blocks.2j+1.{...}, bottleneck.{...}, output.{weight,bias}); tensor i starts at
 * trainer_tensor_offset(i) and the array holds trainer_num_params() floats. Features must be multiples of 32 whose 64-aligned
 * width is a power of two in [64,1024] (32 -> stored zero-extended to 64, as in the inference plan: the deployed topology
 * [32,64,128] trains; parameters and gradients keep the reference's shapes), features[0] in {32,64,128}; out_channels in [1,8]
 * (the reference trains 1; with more, logits / dlogits are NCHW [batch][out_channels][H][W]);
 * batch must give every level's 128-pixel box a multiple of 16 rows. */
typedef struct unet_b200_trainer unet_b200_trainer;
int unet_b200_trainer_create(unet_b200_trainer** out, int batch, int H, int W, int in_channels, int out_channels,
                             const int* features, int levels);
void unet_b200_trainer_destroy(unet_b200_trainer* t);
size_t unet_b200_trainer_workspace_bytes(const unet_b200_trainer* t);
long long unet_b200_trainer_num_params(const unet_b200_trainer* t);
int unet_b200_trainer_num_tensors(const unet_b200_trainer* t);
long long unet_b200_trainer_tensor_offset(const unet_b200_trainer* t, int idx); /* idx == num_tensors: total */
/* workspace_dev: caller-owned, 1024-byte aligned, trainer_workspace_bytes() bytes. */
int unet_b200_trainer_bind(unet_b200_trainer* t, void* workspace_dev);
/* UNet.forward in train mode (README.md:1460-1481 under model.train()): x bf16 NHWC4 [batch][H][W][4] -> logits fp32
 * [batch][H][W] (out_channels > 1: NCHW [batch][out_channels][H][W]). running_mean / running_var: HOST arrays of plan_num_convs device pointers (fp32 [C] each, plan conv
 * order) updated with `momentum` and the unbiased batch variance as nn.BatchNorm2d does; either may be NULL. */
int unet_b200_train_forward(unet_b200_trainer* t, const void* x_nhwc4_dev, const float* params_dev,
                            float* const* running_mean, float* const* running_var, float momentum, float eps,
                            float* logits_dev, void* stream);
/* loss.backward() (README.md:2078) from d loss / d logits (fp32 [batch][H][W]): overwrites grads_dev (flat, same layout
 * as params_dev) with the gradient of every parameter. Needs the activations of the preceding train_forward. */
int unet_b200_train_backward(unet_b200_trainer* t, const float* dlogits_dev, const float* params_dev, float* grads_dev,
                             void* stream);
/* The same backward in stages, cut where a contiguous range of the flat gradient becomes final, so that a data-parallel
 * caller can exchange that range while the rest of the backward runs ("bucketed, launched on a side stream
 * as each level's wgrad finishes"). trainer_num_stages = 2*levels + 2: stage 0 = head (also clears grads_dev), then the
 * decoder levels shallowest first, the bottleneck, the encoder levels deepest first; stages must be called in order.
 * trainer_stage_range gives the flat range [lo, hi) that is final once the stage's work has completed. A stage only enqueues
 * work on `stream` and on the trainer's internal weight-gradient stream; trainer_join makes `waiter_stream` (which may be
 * `stream` itself) wait for everything enqueued so far. */
int unet_b200_trainer_num_stages(const unet_b200_trainer* t);
int unet_b200_trainer_stage_range(const unet_b200_trainer* t, int stage, long long* lo, long long* hi);
int unet_b200_train_backward_stage(unet_b200_trainer* t, int stage, const float* dlogits_dev, const float* params_dev,
                                   float* grads_dev, void* stream);
int unet_b200_trainer_join(unet_b200_trainer* t, void* stream, void* waiter_stream);
/* Data-parallel training over NVLink without a separate collective (replaces loss.backward() + DDP all-reduce +
 * optimizer.step() of README.md:2076-2079 when world > 1). The flat gradient is cut into `world` contiguous shards of
 * S = roundup4(ceil(n/world)) elements, rank r owning [r*S, min(n,(r+1)*S)). Buffers are peer-mapped (torch symmetric
 * memory / CUDA IPC): *_bases_dev are device arrays of the `world` base pointers as mapped in THIS process.
 *
 * push: train_backward_p2p sends every gradient atomic of the backward straight to the OWNER rank's buffer through
 *   grad_bases_dev[owner], i.e. the reduce-scatter happens inside the wgrad / BN / bias epilogues. grads_local
 *   (== grad_bases_dev[rank]) is NOT cleared here; adamw_step_p2p(grad_bases_dev = NULL) clears the owned shard after use.
 * pull: plain train_backward into the local (peer-mapped) buffer, then adamw_step_p2p(grad_bases_dev != NULL) sums the
 *   owned shard over all ranks' buffers with 16-byte NVLink loads.
 * Either way AdamW runs on the owned shard only (optimizer state exists only for the shard, exp_avg*_shard_dev hold S
 * elements) and stores the new values into EVERY rank's flat parameter buffer: the all-gather is those NVLink stores.
 * The caller puts a device-side barrier across ranks between the backward and adamw_step_p2p, and after the latter. */
int unet_b200_train_backward_p2p(unet_b200_trainer* t, const float* dlogits_dev, const float* params_dev, float* grads_local_dev,
                                 float* const* grad_bases_dev, int world, void* stream);
int unet_b200_adamw_step_p2p(float* const* param_bases_dev, float* const* grad_bases_dev, int world, int rank,
                             float* grads_local_dev, float* exp_avg_shard_dev, float* exp_avg_sq_shard_dev, long long n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, const int* step_dev, float grad_scale,
                             void* stream);
/* The same kernel on an explicit flat range [lo, hi) (lo a multiple of 4; m_dev / v_dev hold hi - lo elements: the caller's
 * optimizer state for exactly this range) - what a bucketed exchange calls once per bucket with this rank's part of the
 * bucket. lr_dev as in adamw_step_dev. */
int unet_b200_adamw_range_p2p(float* const* param_bases_dev, float* const* grad_bases_dev, int world, int rank,
                              float* grads_local_dev, long long lo, long long hi, float* m_dev, float* v_dev, float lr,
                              const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, const int* step_dev,
                              float grad_scale, void* stream);
int unet_b200_adamw_range_multimem(float* params_mc_dev, const float* grads_mc_dev, const float* params_local_dev, long long lo,
                                   long long hi, float* m_dev, float* v_dev, float lr, const float* lr_dev, float beta1,
                                   float beta2, float eps, float weight_decay, const int* step_dev, float grad_scale,
                                   void* stream);
/* NVSwitch (NVLS) form of adamw_step_p2p's pull mode: params_mc_dev / grads_mc_dev are MULTICAST addresses of the symmetric
 * flat parameter / gradient buffers (cuMulticast* / torch symmetric memory multicast_ptr); the shard's gradient sum is formed
 * by multimem.ld_reduce inside the switch and the new parameters reach every replica through multimem.st. params_local_dev is
 * this rank's own (unicast) parameter buffer. Same barriers as adamw_step_p2p. */
int unet_b200_adamw_step_multimem(float* params_mc_dev, const float* grads_mc_dev, const float* params_local_dev, int world,
                                  int rank, float* exp_avg_shard_dev, float* exp_avg_sq_shard_dev, long long n, float lr,
                                  float beta1, float beta2, float eps, float weight_decay, const int* step_dev, float grad_scale,
                                  void* stream);
/* out_dev[i] = sum over replicas of x[lo + i] through the multicast address x_mc_dev (multimem.ld_reduce); used to check the
 * reduced gradient of the NVLS exchange. */
int unet_b200_multimem_reduce(const float* x_mc_dev, long long lo, long long n, float* out_dev, void* stream);
/* BCEDiceLoss (README.md:1855-1893): losses3_dev = {total, bce, dice}; dlogits_dev (optional) = d total / d logits.
 * target fp32, same shape as logits; scratch4_dev: 4 doubles. */
int unet_b200_bce_dice_loss(const float* logits_dev, const float* target_dev, size_t n, float pos_weight, float bce_weight,
                            float dice_weight, float smooth, double* scratch4_dev, float* losses3_dev, float* dlogits_dev,
                            void* stream);
/* validate() metrics (README.md:2086-2120) in one pass: out4_dev = {total loss, bce, dice loss, compute_dice(sigmoid(z) >
 * threshold, target)}; scratch6_dev: 6 doubles. */
int unet_b200_validation_metrics(const float* logits_dev, const float* target_dev, size_t n, float pos_weight,
                                 float bce_weight, float dice_weight, float smooth, float threshold, double* scratch6_dev,
                                 float* out4_dev, void* stream);
/* torch.optim.AdamW step (README.md:2173-2174) on flat fp32 arrays; grads are multiplied by grad_scale first
 * (1/world_size after a gradient all-reduce(sum)); step counts from 1. */
int unet_b200_adamw_step(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, size_t n,
                         float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                         void* stream);

/* Same, with the step count read from device memory (int32, >= 1) so the call can be captured in a CUDA graph; lr_dev
 * (optional, fp32 scalar in device memory) overrides lr, so a learning-rate schedule does not need a new graph either. */
int unet_b200_adamw_step_dev(float* params_dev, const float* grads_dev, float* exp_avg_dev, float* exp_avg_sq_dev, size_t n,
                             float lr, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay,
                             const int* step_dev, float grad_scale, void* stream);

/* ---- single training ops (same kernels the trainer runs; exposed for parity tests and reuse) ------------------------ *
 * All activations bf16 NHWC, gradients of activations bf16 NHWC, weight gradients fp32 in the PyTorch layout and
 * ACCUMULATED into dw (zero it first). */
/* dgrad operand of a 3x3 conv: wd bf16 [Cin][9][Cout], wd[ci][t][co] = w[co][ci][8-t]; the input gradient is then
 * unet_b200_conv3x3(dy, Cout, NULL, 0, wd, zero_bias, ..., Cout=Cin, relu=0). */
int unet_b200_pack_conv3x3_dgrad(const float* w_dev, int Cout, int Cin, void* wd_dev, void* stream);
/* dgrad operand of ConvTranspose2d(Cin, f, 2, 2): wd bf16 [Cin][4f], wd[ci][q*f+co] = w[ci][co][q]. */
int unet_b200_pack_convT2x2_dgrad(const float* w_dev, int Cin, int f, void* wd_dev, void* stream);
/* dw[Cout][C0+C1][3][3] += conv-weight gradient for x = cat(x0, x1) [B,H,W,C0|C1] and dy [B,H,W,Cout].
 * C0+C1 must be 64 or a multiple of 128 (with C0 a multiple of 64). */
int unet_b200_conv3x3_wgrad(const void* x0_dev, int C0, const void* x1_dev, int C1, const void* dy_dev, int B, int H, int W,
                            int Cout, float* dw_dev, void* stream);
/* Stem: x NHWC4 bf16, dy [B,H,W,Cout] -> dw[Cout][Cin][3][3] += ... (Cout == 64 on tensor cores, else Cout <= 128). */
int unet_b200_stem_wgrad(const void* x_nhwc4_dev, const void* dy_dev, int B, int H, int W, int Cin, int Cout, float* dw_dev,
                         void* stream);
/* ConvT backward. x [B,H,W,Cin]; dup = gradient w.r.t. the [B,2H,2W,f] output with a pixel pitch of dup_pitch elements
 * (>= f: it may be a channel slice of a wider tensor). dw fp32 [Cin][f][2][2] +=, dbias fp32 [f] += (optional). */
int unet_b200_convT2x2_wgrad(const void* x_dev, int Cin, const void* dup_dev, int dup_pitch, int B, int H, int W, int f,
                             float* dw_dev, float* dbias_dev, void* stream);
int unet_b200_convT2x2_dgrad(const void* dup_dev, int dup_pitch, const void* wd_dev, int B, int H, int W, int Cin, int f,
                             void* dx_dev, void* stream);
/* BatchNorm2d (training mode) + ReLU (+ optional 2x2 max-pool) on a raw conv output y [B,H,W,C]:
 * stats4 fp32 [4][C] receives {mean, invstd, scale, shift}; running stats (optional) are updated with `momentum`;
 * scratch2: 2*C doubles. a = relu(y*scale+shift); pool (optional) = maxpool2x2(a). */
int unet_b200_bn_relu_train_fwd(const void* y_dev, const float* gamma_dev, const float* beta_dev, int B, int H, int W, int C,
                                float eps, float momentum, float* running_mean_dev, float* running_var_dev, void* a_dev,
                                void* pool_dev, float* stats4_dev, double* scratch2_dev, void* stream);
/* Backward of the same: g holds d loss / d a on entry and d loss / d y on return (in place); dgamma/dbeta fp32 [C]. */
int unet_b200_bn_relu_bwd(void* g_dev, const void* y_dev, const float* stats4_dev, int B, int H, int W, int C,
                          float* dgamma_dev, float* dbeta_dev, void* stream);
/* dA = dskip (optional, pixel pitch skip_pitch elements) + max-pool backward of dP through a [B,H,W,C]
 * (first maximum in scan order wins, as ATen). */
int unet_b200_maxpool2x2_bwd(const void* a_dev, const void* dP_dev, const void* dskip_dev, int skip_pitch, int B, int H,
                             int W, int C, void* dA_dev, void* stream);

/* ---- split-precision ("fp32-class") layers: the 1e-4 logit gate of the fp32 path (BASELINE.json north_star) -------------
 * Replaces the same reference ops as the bf16 entry points above (README.md:1449-1458 conv block, :1441-1443 ConvT,
 * :1429 pool) at higher precision. Every activation and (BN-folded) weight v is carried as two bf16 numbers
 * hi = bf16(v), lo = bf16(v - hi); a product is hi*hi + lo*hi + hi*lo accumulated in fp32 on the tensor cores (the same
 * tcgen05 implicit-GEMM kernel, three K passes). Activation tensors are bf16 [B,H,W,2C] = [hi C | lo C].
 * C0 / C1 / Cin / Cout / f are LOGICAL channel counts (multiples of 64). */
int unet_b200_pack_conv3x3_split(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* mean_dev,
                                 const float* var_dev, float eps, int Cout, int C0, int C1, void* wp_dev /* bf16 [Cout][9][3*(C0+C1)] */,
                                 float* bias_dev, void* stream);
int unet_b200_conv3x3_split(const void* x0_dev, int C0, const void* x1_dev, int C1, const void* wp_dev, const float* bias_dev,
                            int B, int H, int W, int Cout, int relu, void* y_dev /* [B,H,W,2*Cout] */, void* stream);
int unet_b200_pack_convT2x2_split(const float* w_dev, int Cin, int f, void* wp_dev /* bf16 [4f][3*Cin] */, void* stream);
int unet_b200_convT2x2_split(const void* x_dev, int Cin, const void* wp_dev, const float* bias_dev, int B, int H, int W, int f,
                             void* y_dev /* [B,2H,2W,2f] */, void* stream);
/* Stem on the fp32 pipes straight from the module's fp32 NCHW input; weights from unet_b200_pack_stem_fp32 (same layout as
 * unet_b200_pack_stem, [9][4][Cout] fp32, without the bf16 rounding). */
int unet_b200_pack_stem_fp32(const float* w_dev, const float* gamma_dev, const float* beta_dev, const float* mean_dev,
                             const float* var_dev, float eps, int Cout, int Cin, float* ws_dev, float* bias_dev, void* stream);
int unet_b200_stem_conv_split(const float* x_nchw_dev, const float* ws_dev, const float* bias_dev, int B, int H, int W, int Cin,
                              int Cout, int relu, void* y_dev /* [B,H,W,2*Cout] */, void* stream);
/* 2x2/2 max-pool on hi + lo; x [B,H,W,2C] -> y [B,H/2,W/2,2C]. The head is unet_b200_head with C = 2*f0 and the weight
 * vector repeated twice. */
int unet_b200_maxpool2x2_split(const void* x_dev, int B, int H, int W, int C, void* y_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H */
