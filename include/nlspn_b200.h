/*
 * nlspn_b200.h -- C ABI of the B200-native NLSPN propagation library (libnlspn_b200.so).
 *
 * This is the drop-in boundary for ONE hot path of XJTUXYC/NLSPN_ECCV20: the iterative
 * non-local spatial propagation.  Reference paths below are relative to the reference
 * repository root.  Every entry point
 *   - takes plain device pointers (fp32, NCHW, contiguous) and sizes; no torch types,
 *   - allocates nothing: outputs and workspace are caller-owned (size: *_workspace_bytes),
 *   - only enqueues work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream) and never synchronises, unless documented (the *_host entry),
 *   - returns 0 on success, a NEGATIVE nlspn_status on a validation error (nothing was
 *     enqueued), or a POSITIVE cudaError_t if the CUDA runtime refused a launch;
 *     nlspn_last_error() returns a thread-local message for the last non-zero return.
 * The library keeps no global mutable state apart from per-device attributes that are
 * queried once (SM count, L2 size) under a mutex; calls are re-entrant.
 *
 * Tensor vocabulary (K = prop_kernel, KK = K*K, N = KK-1 neighbours, T = prop_time):
 *   feat_init   [B,1,H,W]    initial depth                 (pred_init, nlspnmodel.py:297)
 *   guidance    [B,3N,H,W]   2N offset channels (pair n = channels 2n:dh, 2n+1:dw) followed
 *                            by N raw affinities            (off_aff, nlspnmodel.py:301-305)
 *   confidence  [B,1,H,W]    raw confidence or NULL         (nlspnmodel.py:313)
 *   feat_fix    [B,1,H,W]    sparse depth or NULL           (dep, nlspnmodel.py:273)
 *   offset      [B,2KK,H,W]  offsets after the zero centre pair was inserted (nlspnmodel.py:252-259)
 *   aff         [B,KK,H,W]   normalised affinities incl. the centre weight    (nlspnmodel.py:179-201,261-269)
 *   conf_fixed  [B,1,H,W]    (1-m)*confidence + m, m = [feat_fix>0]          (nlspnmodel.py:328-334)
 *   gamma       [1]          aff_scale_const, a DEVICE scalar (nlspnmodel.py:93-104); passing the
 *                            parameter's device address avoids a host sync per call
 *   list_feat   [T,B,1,H,W]  state after every iteration    (list_pred, nlspnmodel.py:363)
 *   src         [S,B,1,H,W]  the planes the gather read: src[t] = x_t * conf_fixed (x_0 = blended
 *                            feat_init).  S = T keeps all of them for backward; S = 2 ping-pongs
 *                            (inference); without confidence src[0] only (S = 1) -- later
 *                            iterations gather list_feat directly.
 */
#ifndef NLSPN_B200_H
#define NLSPN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NLSPN_ABI_VERSION 1

#if defined(__GNUC__)
#define NLSPN_API __attribute__((visibility("default")))
#else
#define NLSPN_API
#endif

/* negative return codes */
enum nlspn_status {
    NLSPN_OK = 0,
    NLSPN_ERR_NULL = -1,        /* a required pointer is NULL                                  */
    NLSPN_ERR_SHAPE = -2,       /* B/H/W/T <= 0 or sizes overflow int32 indexing               */
    NLSPN_ERR_KERNEL = -3,      /* prop_kernel not one of 3, 5, 7                              */
    NLSPN_ERR_DOMAIN = -4,      /* DCN call outside the supported domain (see nlspn_dcn_*)     */
    NLSPN_ERR_WORKSPACE = -5,   /* workspace too small or misaligned                           */
    NLSPN_ERR_AFFINITY = -6,    /* unknown affinity mode                                       */
    NLSPN_ERR_ALIGN = -7        /* a tensor pointer is not 16-byte aligned                     */
};

/* affinity normalisation modes, nlspnmodel.py:93-104,179-197 */
enum nlspn_affinity { NLSPN_AFF_AS = 0, NLSPN_AFF_ASS = 1, NLSPN_AFF_TC = 2, NLSPN_AFF_TGASS = 3 };

/* flags */
#define NLSPN_FLAG_PRESERVE_INPUT 1u /* args.preserve_input, nlspnmodel.py:328,341-344,355-357 (needs feat_fix) */
#define NLSPN_FLAG_ALWAYS_CLIP    2u /* args.always_clip,    nlspnmodel.py:346-348,359-361                       */
/* args.offset == False (the fork's default, src/config.py:272-275): fixed-local propagation,
 * nlspnmodel.py:209-224 -- 3x3 replicate-padded weighted sum instead of the deformable gather.
 * guidance is then [B,N,H,W] (raw affinities only), `offset`/`g_offset_ext` are NULL, K must be 3. */
#define NLSPN_FLAG_NO_OFFSET 4u
/* UPSTREAM (zzangjinsun/NLSPN_ECCV20) semantics, the ones BASELINE.json's north-star prose describes
 * (right-hand column).  PARITY UNPINNED: that code is not in the reference tree; these
 * modes are checked against oracle/torchvision_port.py's restatement only.
 *   BLEND_PRE     the input-preserving blend runs BEFORE every gather only; list_feat[t] and
 *                 feat_result are the raw gathers (the fork blends after every iteration).
 *   CONF_SAMPLED  the confidence is not multiplied into the state; instead each neighbour's affinity
 *                 is multiplied, before the abs-sum normalisation, by the confidence sampled with a
 *                 1x1 deformable gather at that neighbour's offset.  `confidence` is then required by
 *                 nlspn_backward too, `conf_fixed` must be NULL, and g_confidence receives a scatter.
 *   LEGACY        with CONF_SAMPLED: add the tap displacement to the sampling offset (--legacy).
 * Not combinable with ALWAYS_CLIP, NO_OFFSET, BWD_PER_ITERATION. */
#define NLSPN_FLAG_BLEND_PRE    8u
#define NLSPN_FLAG_CONF_SAMPLED 16u
#define NLSPN_FLAG_LEGACY       32u
/* debugging aid: run the backward as T per-iteration kernels that re-read/re-write the
 * gradient accumulators and scatter with scalar atomics (the reference's structure); results
 * agree with the default two-pass backward up to fp32 summation order. */
#define NLSPN_FLAG_BWD_PER_ITERATION 0x100u
/* nlspn_backward only: bit-identical gradients from run to run.  The default forms of the state-gradient pass
 * (and the reference's col2im, modulated_deform_im2col_cuda.cuh:229-252; its own test admits the difference,
 * deformconv/test.py:627-631) add floating-point contributions in scheduling order.  With this flag the
 * transposed operator is tabulated exactly (CSR over destination pixels, rows sorted by source) and every sum
 * runs in a fixed order; gamma's gradient uses a fixed-order tree.  Slower (about 2x on the state-gradient pass)
 * and needs nlspn_backward_workspace_bytes_ex(..., flags) bytes of workspace.  Not combinable with NO_OFFSET,
 * BWD_PER_ITERATION, CONF_SAMPLED.  Limits: H*W <= 2^24, 4*(K*K-1)*B*H*W < 2^32. */
#define NLSPN_FLAG_DETERMINISTIC 0x200u

NLSPN_API int nlspn_abi_version(void);
NLSPN_API const char *nlspn_last_error(void);

/* Number of kernels this library has launched in this process so far (monotonic; used by
 * bench.py to report gpu_launches). */
NLSPN_API unsigned long long nlspn_launch_count(void);

/* Tuning options.  The NLSPN_<NAME> environment variables seed the defaults ONCE, when the
 * library is loaded; nothing on the call path reads the environment.  Names: tiled, persist, pdl, fwd_th,
 * param_th, state_tma, state_gather, gather_compact, state_zero3, state_minb, group_images, stream_hint,
 * param_factored, dcn_blocked, state_local, local_prefetch, local_minb, sched_minb.
 * value -1 = auto where the default depends on the shape.
 * Options that select the form of the backward (state_gather) change nlspn_backward_workspace_bytes: query
 * after setting them (a too-small workspace is always refused, never overrun). */
NLSPN_API int nlspn_set_option(const char *name, int value);
NLSPN_API int nlspn_get_option(const char *name, int *value);
NLSPN_API int nlspn_reset_options(void);

/* Optional per-kernel timing for bench.py's roofline.  While enabled, every kernel launch of
 * this library is bracketed by CUDA events on the caller's stream (this perturbs throughput a
 * little: never enable it inside a timed region whose `value` is reported).
 * nlspn_profile_enable(1) clears the table and starts, (0) stops.  nlspn_profile_read
 * synchronises the recorded events and returns per-class summed milliseconds and launch counts
 * (arrays of at least nlspn_profile_classes() entries). */
NLSPN_API int nlspn_profile_enable(int on);
NLSPN_API int nlspn_profile_classes(void);
NLSPN_API const char *nlspn_profile_class_name(int cls);
NLSPN_API int nlspn_profile_read(double *ms, long long *launches, int n);

/* Device facts the host side sizes its shards with: SM count and L2 bytes of `device`. */
NLSPN_API int nlspn_device_info(int device, int *sm_count, int *l2_bytes);

/* ---- fused prologue ------------------------------------------------------------------
 * Replaces _off_insert + _affinity_normalization + _aff_insert + the mask/confidence fix-up
 * + the first input-preserving blend + the first confidence pre-multiply
 * (nlspnmodel.py:323-325,328-334,341-351: ~20 elementwise kernels in the reference).
 * Writes offset, aff, conf_fixed (if confidence != NULL) and src0 = x_0 * conf_fixed where
 * x_0 = blend(feat_init) (clamped at 0 under ALWAYS_CLIP). */
NLSPN_API int nlspn_prologue_fwd(const float *guidance, const float *confidence, const float *feat_init,
                       const float *feat_fix, const float *gamma, int affinity, unsigned flags,
                       int B, int H, int W, int K,
                       float *offset, float *aff, float *conf_fixed, float *src0, void *stream);

/* ---- T propagation iterations ---------------------------------------------------------
 * Replaces the loop nlspnmodel.py:340-363, i.e. T x { x*c ; ModulatedDeformConvFunction
 * (modulated_deform_conv_func.py:17-36 -> modulated_deform_conv_cuda.cu:19-121 ->
 * modulated_deform_im2col_cuda.cuh:127-194) ; blend ; [clamp] ; append }.
 * `src` must already hold plane 0 (from nlspn_prologue_fwd).  S as described above. */
NLSPN_API int nlspn_propagate_fwd(const float *offset, const float *aff, const float *conf_fixed,
                        const float *feat_fix, unsigned flags, int B, int H, int W, int K, int T,
                        float *src, int S, float *list_feat, void *stream);

/* ---- fused forward: prologue + T iterations in one call ---------------------------------
 * Same results as nlspn_prologue_fwd followed by nlspn_propagate_fwd, but enqueued group-major
 * (all T iterations of a group of images before the next group) so a group's offsets and
 * affinities are still in L2 when the iterations read them.  This is what the NLSPN module
 * calls. */
NLSPN_API int nlspn_forward(const float *guidance, const float *confidence, const float *feat_init,
                            const float *feat_fix, const float *gamma, int affinity, unsigned flags,
                            int B, int H, int W, int K, int T,
                            float *offset, float *aff, float *conf_fixed, float *src, int S,
                            float *list_feat, void *stream);

/* ---- backward of prologue + loop ------------------------------------------------------
 * Replaces T x ModulatedDeformConvFunction.backward (modulated_deform_conv_func.py:38-56 ->
 * modulated_deform_conv_cuda.cu:124-280 -> cuh:196-328), autograd's elementwise backward of
 * nlspnmodel.py:344-357 and the backward of the normalisation (nlspnmodel.py:179-201,252-269).
 *   g_list        HOST array of T device pointers, entry t = DIRECT upstream gradient
 *                 [B,1,H,W] of list_feat[t]; a NULL entry means a zero gradient (not read)
 *   g_offset_ext  [B,2KK,H,W]  upstream gradient of the `offset` output or NULL
 *   g_aff_ext     [B,KK,H,W]   upstream gradient of the `aff` output or NULL
 * `confidence` (raw) is read only under NLSPN_FLAG_CONF_SAMPLED and may be NULL otherwise.
 * Outputs (overwritten): g_feat_init [B,1,H,W], g_guidance [B,3N,H,W], g_confidence [B,1,H,W]
 * (NULL iff conf_fixed is NULL), g_gamma: one double (device memory).
 * `src` must be the S = T array written by the forward (S = 1 without confidence).
 * Gradient wrt feat_fix is not produced (mask_fix is detached, nlspnmodel.py:330).
 * The scatter uses fp32 atomics: summation order, hence the last bits, vary run to run,
 * as in the reference (deformconv/test.py:627-631).
 * Workspace: query with the same (B, H, W, K, T) -- and the same environment -- as the call.  The size
 * depends on the form the state-gradient pass takes: tile-local transpose (K = 3, W % 4 == 0): two padded planes
 * and a packed record of 128 B per pixel, ~0.87 GB at KITTI B = 8, K = 3, T = 18; RED scatter (otherwise, or
 * T < 8 at K >= 5): three sets of blocked planes, ~0.55 GB; tabulated gather (K >= 5 and T >= 8): a table of
 * 16 B x CAP(K) x blocks, ~3.2 GB at K = 5 and ~4.9 GB at K = 7 for the same batch. */
NLSPN_API size_t nlspn_backward_workspace_bytes(int B, int H, int W, int K, int T);
/* the same query for a call with `flags` (NLSPN_FLAG_DETERMINISTIC changes the workspace layout) */
NLSPN_API size_t nlspn_backward_workspace_bytes_ex(int B, int H, int W, int K, int T, unsigned flags);
NLSPN_API int nlspn_backward(const float *guidance, const float *feat_init, const float *feat_fix,
                   const float *confidence, const float *offset, const float *aff, const float *conf_fixed,
                   const float *src, int S, const float *list_feat, const float *const *g_list,
                   const float *g_offset_ext, const float *g_aff_ext, const float *gamma, int affinity,
                   unsigned flags, int B, int H, int W, int K, int T,
                   float *g_feat_init, float *g_guidance, float *g_confidence, double *g_gamma,
                   void *workspace, size_t workspace_bytes, void *stream);

/* ---- single-step operator: the reference's existing native boundary (B1) ---------------
 * Same argument order and meaning as DCN.modulated_deform_conv_forward / _backward
 * (src/model/deformconv/src/modulated_deform_conv.h:10-25,46-62; pybind names in
 * src/model/deformconv/src/vision.cpp:9-10).  Supported domain = what nlspnmodel.py:107-121,
 * 205-208 passes: C_in = C_out = 1, group = deformable_group = 1, stride 1, dilation 1,
 * pad = (K-1)/2, kernel_h = kernel_w = K in {3,5,7}; any weight[1,1,K,K] and bias[1].
 * Anything else returns NLSPN_ERR_DOMAIN (never a silent wrong answer).  im2col_step is
 * accepted and ignored (no batch-divisibility restriction).
 *   input [B,1,H,W], offset [B,2KK,H,W] (all KK taps deformable), mask [B,KK,H,W].
 * backward overwrites grad_input, grad_offset, grad_mask, grad_weight[KK], grad_bias[1]. */
NLSPN_API int nlspn_dcn_forward(const float *input, const float *weight, const float *bias,
                      const float *offset, const float *mask,
                      int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                      int dilation_h, int dilation_w, int group, int deformable_group,
                      int im2col_step, int B, int C, int H, int W, float *output, void *stream);
NLSPN_API int nlspn_dcn_backward(const float *input, const float *weight, const float *bias,
                       const float *offset, const float *mask, const float *grad_output,
                       int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                       int dilation_h, int dilation_w, int group, int deformable_group,
                       int im2col_step, int B, int C, int H, int W,
                       float *grad_input, float *grad_offset, float *grad_mask,
                       float *grad_weight, float *grad_bias, void *stream);

/* Same operator with a caller-owned workspace (nlspn_dcn_backward_workspace_bytes): grad_input is scattered with
 * ONE vector reduction per tap into four phase-shifted 2x2-blocked copies of the plane and collected afterwards,
 * and the gather source travels as TMA boxes -- the form nlspn_eccv20_b200.dcn uses.  Falls back to the scalar
 * form of nlspn_dcn_backward when W % 4 != 0 or the workspace is NULL / too small. */
NLSPN_API size_t nlspn_dcn_backward_workspace_bytes(int B, int H, int W, int K);
NLSPN_API int nlspn_dcn_backward_ws(const float *input, const float *weight, const float *bias,
                       const float *offset, const float *mask, const float *grad_output,
                       int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                       int dilation_h, int dilation_w, int group, int deformable_group,
                       int im2col_step, int B, int C, int H, int W,
                       float *grad_input, float *grad_offset, float *grad_mask,
                       float *grad_weight, float *grad_bias, void *workspace, size_t workspace_bytes,
                       void *stream);

/* ---- ONE fused iteration of the loop body (SURVEY 8b nlspn_step_fwd/_bwd; row f2) -----------------------
 * For callers whose affinities change between iterations -- the fork's GRU mode, nlspnmodel.py:365-373 --
 * and therefore cannot use the T-iteration entries.  Replaces, per call, nlspnmodel.py:350-361:
 *   out      = clamp(blend(G(src_prev; offset, aff)))       G = ModulatedDeformConvFunction with w = 1, b = 0
 *              (:205-208) or, under NO_OFFSET / offset == NULL, the fixed-local 3x3 sum (:209-224)
 *   src_next = out * conf_fixed                              (the next call's src_prev; NULL without confidence,
 *              the next call then reads `out`)
 * src_prev [B,1,H,W] is the PRE-MULTIPLIED state (x * conf_fixed); offset [B,2KK,H,W] with the zero centre pair
 * (nlspnmodel.py:252-259); aff [B,KK,H,W] normalised, centre included (:261-269).
 * flags: PRESERVE_INPUT (needs feat_fix), ALWAYS_CLIP, NO_OFFSET.
 * Backward: g_out / g_src_next are the upstream gradients of the two outputs (either may be NULL = zero);
 * `out` is the forward's result (read when g_src_next or ALWAYS_CLIP is given).  Overwrites g_src_prev [B,1,H,W],
 * g_offset [B,2KK,H,W] (centre pair zero; NULL under NO_OFFSET), g_aff [B,KK,H,W], g_conf [B,1,H,W] (NULL iff
 * conf_fixed is NULL; gradient wrt conf_fixed of THIS step only). */
NLSPN_API int nlspn_step_fwd(const float *src_prev, const float *offset, const float *aff, const float *conf_fixed,
                             const float *feat_fix, unsigned flags, int B, int H, int W, int K,
                             float *out, float *src_next, void *stream);
NLSPN_API size_t nlspn_step_bwd_workspace_bytes(int B, int H, int W, int K, unsigned flags);
NLSPN_API int nlspn_step_bwd(const float *src_prev, const float *offset, const float *aff, const float *conf_fixed,
                             const float *feat_fix, const float *out, const float *g_out, const float *g_src_next,
                             unsigned flags, int B, int H, int W, int K,
                             float *g_src_prev, float *g_offset, float *g_aff, float *g_conf,
                             void *workspace, size_t workspace_bytes, void *stream);

/* ---- the three final head convolutions in front of the propagation (first half) ----------
 * Replaces, in NLSPNModel.forward (nlspnmodel.py:297,301,313 with the layers of :69-86):
 *     pred_init  = id_dec0     (cat(id_fd1,      fe1))      3x3 conv 128 -> 1,  ReLU
 *     off_aff    = off_aff_dec0(cat(off_aff_fd1, fe1))      3x3 conv 128 -> 3N, no activation   ("guidance")
 *     confidence = cf_dec0     (cat(cf_fd1,      fe1))      3x3 conv 128 -> 1,  Sigmoid
 * by ONE tcgen05 (kind::tf32, fp32 accumulation in tensor memory) implicit GEMM that reads the four 64-channel
 * NCHW tensors where they lie (no concatenation; for K = 3, 5 and W % 4 == 0 as MN-major operands straight from TMA
 * boxes, kernels_head2.cuh).  TF32 is what cuDNN computes these layers in under PyTorch's
 * default torch.backends.cudnn.allow_tf32 = True.
 *   nlspn_heads_packed_floats(K)  floats of the packed weight matrix
 *   nlspn_heads_pack              packs w_id [1,128,3,3], w_oa [3N,128,3,3], w_cf [1,128,3,3] (device pointers;
 *                                 input channels 0..63 = the head's own branch, 64..127 = fe1, as torch.cat orders
 *                                 them) into `packed` (device, 16-byte aligned); call again when the weights change
 *   nlspn_heads_fwd               bias [3N + 2] = (b_id, b_oa[0..3N), b_cf); outputs pred_init [B,1,H,W],
 *                                 guidance [B,3N,H,W], confidence [B,1,H,W] */
NLSPN_API size_t nlspn_heads_packed_floats(int K);
NLSPN_API int nlspn_heads_pack(const float *w_id, const float *w_oa, const float *w_cf, int K, float *packed, void *stream);
NLSPN_API int nlspn_heads_fwd(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                              const float *packed, const float *bias, int B, int H, int W, int K,
                              float *pred_init, float *guidance, float *confidence, void *stream);

/* The same GEMM with the propagation's prologue as its epilogue: `guidance` never reaches HBM (pass NULL), the
 * kernel writes what nlspn_prologue_fwd would have written from it -- offset [B,2K^2,H,W], aff [B,K^2,H,W],
 * conf_fixed [B,1,H,W] (NULL = no confidence propagation), src0 [B,1,H,W] -- next to pred_init and confidence
 * (nlspnmodel.py:297-313 followed by :252-269,179-201,328-351); continue with nlspn_propagate_fwd.  A non-NULL
 * `guidance` is written as well (a training step needs it for the backward).  flags: PRESERVE_INPUT (needs
 * feat_fix), ALWAYS_CLIP.  Implemented for K = 3, 5 and W % 4 == 0 (nlspn_heads_prologue_supported); anything else
 * returns NLSPN_ERR_SHAPE / NLSPN_ERR_DOMAIN and the caller uses nlspn_heads_fwd + nlspn_prologue_fwd. */
NLSPN_API int nlspn_heads_prologue_supported(int W, int K);
NLSPN_API int nlspn_heads_prologue_fwd(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                                       const float *packed, const float *bias, const float *feat_fix, const float *gamma,
                                       int affinity, unsigned flags, int B, int H, int W, int K,
                                       float *pred_init, float *confidence, float *guidance,
                                       float *offset, float *aff, float *conf_fixed, float *src0, void *stream);

/* Weight (and bias) gradients of the same three layers on tcgen05 (training; kernels_head_wgrad.cuh).  Replaces what
 * autograd runs for nlspnmodel.py:69-86,297,301,313 in the reference -- the ReLU / Sigmoid backward, three
 * cudnn_convolution_backward_weight calls on the 128-channel concatenations and three bias sums -- by
 *   nlspn_heads_grad_prep   g_init / g_guidance / g_confidence (upstream gradients of pred_init [B,1,H,W], guidance
 *                           [B,3N,H,W], confidence [B,1,H,W]; NULL = zero) -> g_shift [3][B][3N+2][H][W]: the activation
 *                           derivatives applied (pred_init, confidence = the forward outputs), channels concatenated
 *                           (0 = init, 1 = confidence, 2..3N+1 = guidance), and the whole written three times, shifted
 *                           by +1 / 0 / -1 pixels along x (copy 1, g_shift + B (3N+2) H W, is the plain concatenation: the
 *                           input of the data gradients); g_bias [3N+2] (NULL = skip) = (db_id, db_cf, db_oa[0..3N))
 *   nlspn_heads_wgrad       dw_all [3N+2,128,3,3] (overwritten): rows 0 = dw_id, 1 = dw_cf, 2..3N+1 = dw_oa, input
 *                           channels 0..63 = the head's own branch, 64..127 = fe1.  A NULL branch tensor (id_fd1 /
 *                           oa_fd1 / cf_fd1) leaves its 64-channel block zero.  TF32 products, fp32 accumulation over the
 *                           pixels in tensor memory, split-K partials added with fp32 atomics.
 *   nlspn_heads_dgrad_one   data gradients of the two one-channel heads with respect to their own branches: d_id_fd1,
 *                           d_cf_fd1 [B,64,H,W] (NULL = skip) from g_all = copy 1 of g_shift and the heads' [1,128,3,3]
 *                           weights (input channels 0..63): an fp32 nine-tap stencil at the rate of its store
 *                           (cudnn_convolution_backward_input in the reference)
 *   nlspn_heads_dgrad_pack  (prop_kernel 3) packs the three weight tensors for the wide data gradients into `packed`
 *                           (nlspn_heads_dgrad_packed_floats(K) floats, 16-byte aligned); call again when the weights change
 *   nlspn_heads_dgrad_wide  (prop_kernel 3, nlspn_heads_dgrad_supported) the two wide data gradients as one tcgen05 implicit
 *                           GEMM over the three shifted copies: d_oa_fd1 [B,64,H,W] (NULL = skip; the guidance head's own
 *                           branch) and d_fe1 [B,64,H,W] (the shared fe1: all three heads); TF32 products, fp32 accumulation
 * W % 4 == 0 and 16-byte aligned tensors (nlspn_heads_wgrad_supported); otherwise NLSPN_ERR_SHAPE / NLSPN_ERR_ALIGN. */
NLSPN_API int nlspn_heads_wgrad_supported(int W, int K);
NLSPN_API int nlspn_heads_grad_prep(const float *g_init, const float *pred_init, const float *g_guidance,
                                    const float *g_confidence, const float *confidence, int B, int H, int W, int K,
                                    float *g_shift, float *g_bias, void *stream);
NLSPN_API int nlspn_heads_wgrad(const float *id_fd1, const float *oa_fd1, const float *cf_fd1, const float *fe1,
                                const float *g_shift, int B, int H, int W, int K, float *dw_all, void *stream);
NLSPN_API int nlspn_heads_dgrad_one(const float *g_all, const float *w_id, const float *w_cf, int B, int H, int W, int K,
                                    float *d_id_fd1, float *d_cf_fd1, void *stream);
NLSPN_API size_t nlspn_heads_dgrad_packed_floats(int K);
NLSPN_API int nlspn_heads_dgrad_supported(int W, int K);
NLSPN_API int nlspn_heads_dgrad_pack(const float *w_id, const float *w_oa, const float *w_cf, int K, float *packed, void *stream);
NLSPN_API int nlspn_heads_dgrad_wide(const float *g_shift, const float *packed, int B, int H, int W, int K,
                                     float *d_oa_fd1, float *d_fe1, void *stream);

/* Double-precision variants of the single-step operator: the reference dispatches this op over
 * float and double (AT_DISPATCH_FLOATING_TYPES, modulated_deform_conv_cuda.cu:93,224) and its
 * tests run gradcheck in double (src/model/deformconv/test.py).  Same domain, argument order and
 * error behaviour as the fp32 entries; simple (untuned) kernels. */
NLSPN_API int nlspn_dcn_forward_f64(const double *input, const double *weight, const double *bias,
                      const double *offset, const double *mask,
                      int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                      int dilation_h, int dilation_w, int group, int deformable_group,
                      int im2col_step, int B, int C, int H, int W, double *output, void *stream);
NLSPN_API int nlspn_dcn_backward_f64(const double *input, const double *weight, const double *bias,
                       const double *offset, const double *mask, const double *grad_output,
                       int kernel_h, int kernel_w, int stride_h, int stride_w, int pad_h, int pad_w,
                       int dilation_h, int dilation_w, int group, int deformable_group,
                       int im2col_step, int B, int C, int H, int W,
                       double *grad_input, double *grad_offset, double *grad_mask,
                       double *grad_weight, double *grad_bias, void *stream);

/* ---- debug: integer corners chosen for every tap (exact-index parity test) -------------
 * idx [B,KK,3,H,W] int32: floor(h_im), floor(w_im), valid (validity test of cuh:180),
 * with h_im = (float)(h - pad + i) + offset_h formed exactly as cuh:178-179. */
NLSPN_API int nlspn_debug_indices(const float *offset, int B, int H, int W, int K, int32_t *idx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NLSPN_B200_H */
