/*
 * gssd.h — C ABI of the B200-native GSSD multibox head (libgssd_b200.so).
 *
 * This is the drop-in boundary for the hot path of L0SG/grouped-ssd-pytorch:
 *   PriorBox -> match/encode -> MultiBoxLoss (OHNM) -> Detect (decode / threshold / top-k / NMS)
 * The reference has no FFI of its own: the boundary that exists there is the Python package
 * `ssd_liverdet/layers` (layers/__init__.py:1-2).  Each entry point below names the reference
 * function it replaces (paths relative to /root/reference/ssd_liverdet/).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller (PyTorch's caching allocator, in the shipped host layer) owns every buffer,
 *     including the workspace; nothing is allocated, freed or synchronised in here, so every
 *     call is legal inside CUDA-graph capture;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 on success, a negative GSSD_ERR_* for argument errors (checked before any
 *     launch), or a positive cudaError_t if a launch failed;
 *   - boxes are float32; "center form" = (cx,cy,w,h), "point form" = (xmin,ymin,xmax,ymax);
 *   - ground truth is packed: gt[sum_G,5] rows (xmin,ymin,xmax,ymax,label) and gt_off[B+1]
 *     int32 row offsets (image b owns rows gt_off[b]..gt_off[b+1]), the device-side image of the
 *     reference's `targets` list (multibox_loss.py:67-69);
 *   - float arithmetic is IEEE, unfused (no FMA contraction), in the reference's association order.
 */
#ifndef GSSD_H_
#define GSSD_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSSD_ABI_VERSION 2

#if defined(__GNUC__)
#define GSSD_API __attribute__((visibility("default")))
#else
#define GSSD_API
#endif

#define GSSD_OK          0
#define GSSD_ERR_ARG    -1   /* null pointer / non-positive size / bad enum */
#define GSSD_ERR_LIMIT  -2   /* size beyond what the kernels support (see GSSD_MAX_*) */
#define GSSD_ERR_WS     -3   /* workspace smaller than gssd_workspace_bytes() */
#define GSSD_ERR_VALUE  -4   /* reference-visible ValueError (variance <= 0, nms_thresh <= 0) */
#define GSSD_ERR_EMPTY  -5   /* an image without ground truth (reference: IndexError) */
#define GSSD_ERR_UNSUPPORTED -6  /* this shape has no one-launch form (gssd_mbox_loss_fused): use the two stages */

#define GSSD_MAX_GT_PER_IMAGE   128    /* G per image held in shared memory */
#define GSSD_MAX_PRIORS       49152    /* P: keys of one image must fit one SM's shared memory */
#define GSSD_MAX_TOP_K         1024
#define GSSD_MAX_CLASSES         64
#define GSSD_MAX_FEATURE_MAPS     8
#define GSSD_MAX_ASPECT_RATIOS    8

GSSD_API int         gssd_abi_version(void);
GSSD_API const char *gssd_error_string(int code);
/* number of kernels launched by this library since load (bench.py's gpu_launches claim) */
GSSD_API uint64_t    gssd_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * PriorBox — replaces PriorBox.__init__/forward, layers/functions/prior_box.py:14-172
 * ---------------------------------------------------------------------------------------- */
enum {
    GSSD_PRIOR_V2 = 0,             /* 'v2' and 'v2_512' : prior_box.py:35-56, 116-138 */
    GSSD_PRIOR_V2_CUSTOM = 1,      /* 'v2_custom', 'v2_custom_squareonly', 'v2_custom_512': 58-114 */
    GSSD_PRIOR_LEGACY = 2          /* any other name (e.g. 'v1'), corner form: 141-167 */
};

typedef struct gssd_prior_cfg {
    int32_t version;                                   /* GSSD_PRIOR_* */
    int32_t n_maps;                                    /* len(cfg['feature_maps']) */
    int32_t clip;                                      /* cfg['clip'] */
    int32_t feature_maps[GSSD_MAX_FEATURE_MAPS];
    int32_t n_ar[GSSD_MAX_FEATURE_MAPS];               /* len(cfg['aspect_ratios'][k]) */
    double  min_dim;                                   /* cfg['min_dim'] */
    double  steps[GSSD_MAX_FEATURE_MAPS];
    double  min_sizes[GSSD_MAX_FEATURE_MAPS];
    double  max_sizes[GSSD_MAX_FEATURE_MAPS];
    double  aspect_ratios[GSSD_MAX_FEATURE_MAPS][GSSD_MAX_ASPECT_RATIOS];
    double  variance[2];                               /* validated > 0 (prior_box.py:28-30) */
} gssd_prior_cfg;

/* host-only: number of boxes P the config generates, or a negative GSSD_ERR_* */
GSSD_API int gssd_priorbox_count(const gssd_prior_cfg *cfg_host);
/* out[P,4] float32.  fp64 arithmetic in the reference's order, one rounding to fp32, then clamp. */
GSSD_API int gssd_priorbox(const gssd_prior_cfg *cfg_host, float *out, void *stream);

/* ------------------------------------------------------------------------------------------
 * box_utils — replaces layers/box_utils.py
 * ---------------------------------------------------------------------------------------- */
/* point_form, box_utils.py:4-13 : (cx,cy,w,h) -> (xmin,ymin,xmax,ymax), [n,4] -> [n,4] */
GSSD_API int gssd_point_form(const float *boxes, int n, float *out, void *stream);
/* center_size, box_utils.py:16-25 (documented intent; the reference body is malformed) */
GSSD_API int gssd_center_size(const float *boxes, int n, float *out, void *stream);
/* intersect, box_utils.py:28-46 : a[A,4], b[Bn,4] point form -> out[A,Bn] */
GSSD_API int gssd_intersect(const float *a, int A, const float *b, int Bn, float *out, void *stream);
/* jaccard, box_utils.py:49-67 : IoU matrix out[A,Bn]; union = (area_a + area_b) - inter */
GSSD_API int gssd_jaccard(const float *a, int A, const float *b, int Bn, float *out, void *stream);
/* encode, box_utils.py:114-135 : matched[n,4] point form, priors[n,4] center form -> out[n,4] */
GSSD_API int gssd_encode(const float *matched, const float *priors, int n, float var0, float var1,
                float *out, void *stream);
/* decode, box_utils.py:139-157 : loc[n,4], priors[n,4] -> out[n,4] point form (not clipped) */
GSSD_API int gssd_decode(const float *loc, const float *priors, int n, float var0, float var1,
                float *out, void *stream);
/* log_sum_exp, box_utils.py:160-168 : x[rows,C] -> out[rows]; subtracts the max of the WHOLE
 * tensor.  ws: gssd_workspace_bytes(GSSD_WS_LSE, ...) */
GSSD_API int gssd_log_sum_exp(const float *x, int rows, int C, float *out, void *ws, size_t ws_bytes,
                     void *stream);

/* match, box_utils.py:70-111, batched over B images (the loop at multibox_loss.py:67-72).
 *   priors[P,4] center form; gt/gt_off packed ground truth (sum_G rows; g_max = max rows/image);
 *   loc_t[B,P,4] float32 and conf_t[B,P] int64 are written for every prior (box_utils.py:109-111);
 *   best_truth_idx[B,P] int32 (optional, may be NULL) = final matched GT row within the image.
 * Ties: argmax -> lowest index; shared best prior -> highest GT row wins (box_utils.py:104-105). */
GSSD_API int gssd_match(const float *priors, int P, const float *gt, const int32_t *gt_off, int B,
               int sum_G, int g_max, float threshold, float var0, float var1,
               float *loc_t, int64_t *conf_t, int32_t *best_truth_idx,
               void *ws, size_t ws_bytes, void *stream);

/* nms, box_utils.py:174-238 : boxes[n,4] point form, scores[n].
 *   keep[n] int64 zero-padded, count[1] int32 (device).  Candidates are the top_k scores BEFORE
 *   suppression; union = (area_j - inter) + area_i; kept iff IoU <= overlap.
 *   Equal scores: higher index first (stable ascending sort read from the end). */
GSSD_API int gssd_nms(const float *boxes, const float *scores, int n, float overlap, int top_k,
             int64_t *keep, int32_t *count, void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * MultiBoxLoss — replaces MultiBoxLoss.forward, layers/modules/multibox_loss.py:46-120, and the
 * autograd backward of its two outputs.  Two stages so that a multi-GPU host can all-reduce the
 * two batch-global scalars between them (stats[0] = max of conf as float bits, MAX;
 * stats[1] = number of positives N as int32, SUM).
 * ---------------------------------------------------------------------------------------- */
typedef struct gssd_loss_stats {     /* device-resident header, 16 bytes */
    uint32_t conf_max_ord;           /* x_max of log_sum_exp (box_utils.py:167) over the local batch, as an
                                        order-preserving uint32 (0 = "no value"); combine with MAX */
    int32_t  num_pos_total;          /* local part of N (multibox_loss.py:117); combine with SUM */
    uint32_t done_counter;           /* internal (last-CTA reduction of stage 2) */
    uint32_t reserved;
} gssd_loss_stats;
/* stats_buf = [gssd_loss_stats header][int32 num_pos[B]] : gssd_stats_bytes(B) bytes */
GSSD_API size_t gssd_stats_bytes(int B);

/* Stage 1: matching (box_utils.py:70-108) + batch max of conf (conf may be NULL: header max stays 0).
 *   tags[B,P] uint16 : bit15 = positive, bits0-14 = matched GT row within the image
 *   stats_buf is zeroed and filled by this call. */
GSSD_API int gssd_mbox_match(const float *priors, int P, const float *conf, int C,
                    const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                    float threshold, uint16_t *tags, void *stats_buf, void *stream);

/* Stage 2: encode + smooth-L1 (multibox_loss.py:80-88), mining key with the global-max LSE
 * (91-101), hard-negative selection of min(negpos_ratio*num_pos, P-1) keys per image (102-106;
 * descending key, ties -> lower prior index), cross-entropy over pos|neg (108-113), division by N
 * (117-119) and the gradients of both losses.
 *   stats_buf: as filled by stage 1 on this device.  global_stats[n_global_stats]: the headers of
 *   every rank of a data-parallel job (all-gathered by the host; MAX / SUM are taken in-kernel), or
 *   NULL,0 to use the local header alone.
 *   losses[2] float32 = (loss_l/N, loss_c/N) with this device's images in the numerators;
 *   grad_loc[B,P,4], grad_conf[B,P,C] = d(loss_l)/d(loc), d(loss_c)/d(conf) (NULL, NULL for forward
 *   only); pos_mask/neg_mask[B,P] uint8 optional. */
GSSD_API int gssd_mbox_loss(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                   const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                   const uint16_t *tags, void *stats_buf,
                   const gssd_loss_stats *global_stats, int n_global_stats,
                   int negpos_ratio, float var0, float var1,
                   float *losses, float *grad_loc, float *grad_conf,
                   uint8_t *pos_mask, uint8_t *neg_mask,
                   void *ws, size_t ws_bytes, void *stream);

/* Peer exchange of the loss statistics (data-parallel jobs on one NVLink/NVSwitch box): instead of an all-gather
 * between the two stages, the LAST CTA of stage 1 stores this rank's 16-byte statistics straight into every peer's
 * exchange buffer (peer-mapped memory, one 16-byte slot per rank and step parity, epoch-tagged, system-scope
 * release), and stage 2 spins on its LOCAL buffer until the slots of all ranks carry the current epoch.  No host
 * involvement, legal inside CUDA graphs (the epoch lives in device memory), a few microseconds instead of a
 * collective launch.  Every rank must issue the same sequence of stage-1 / stage-2 calls. */
#define GSSD_XCHG_MAX_RANKS 16
#define GSSD_XCHG_HANDLE_BYTES 64
typedef struct gssd_xchg {
    void   *peers[GSSD_XCHG_MAX_RANKS];   /* device pointers to the exchange buffer of every rank (own included) */
    int32_t rank, world;
    uint32_t timeout_ms;                  /* a kernel that waits longer than this for a peer's statistics traps instead of
                                             wedging the GPU; 0 = wait for ever.  Size it for the longest time two ranks can
                                             be apart when they reach the criterion (a stalled data loader counts) */
    uint32_t reserved;
} gssd_xchg;
/* allocate (cudaMalloc) and zero this rank's exchange buffer; export it for the peers (cudaIpcMemHandle_t) */
GSSD_API int gssd_xchg_create(void **xbuf_out, void *ipc_handle_out_host /* 64 bytes */);
GSSD_API int gssd_xchg_open(const void *ipc_handle_host /* 64 bytes */, void **peer_ptr_out);
GSSD_API int gssd_xchg_close(void *peer_ptr);
GSSD_API int gssd_xchg_destroy(void *xbuf);
/* the two stages of MultiBoxLoss with the peer exchange in between (same arguments as gssd_mbox_match /
 * gssd_mbox_loss otherwise; global_stats is replaced by the exchange) */
GSSD_API int gssd_mbox_match_x(const float *priors, int P, const float *conf, int C,
                      const float *gt, const int32_t *gt_off, int B, int sum_G, int g_max,
                      float threshold, uint16_t *tags, void *stats_buf, const gssd_xchg *x_host, void *stream);
GSSD_API int gssd_mbox_loss_x(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                     const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                     const uint16_t *tags, void *stats_buf, const gssd_xchg *x_host,
                     int negpos_ratio, float var0, float var1,
                     float *losses, float *grad_loc, float *grad_conf,
                     uint8_t *pos_mask, uint8_t *neg_mask,
                     void *ws, size_t ws_bytes, void *stream);

/* MultiBoxLoss.forward (multibox_loss.py:46-120, matching included) and the gradients of its two outputs in ONE launch,
 * for batches whose CTAs are all resident at once (the reference's training batch of 32 is): the two batch-wide scalars
 * meet at two counters in global memory instead of at a kernel boundary, and conf is read from HBM once.
 *   state: gssd_fused_state_bytes() bytes of device memory, zeroed ONCE by the caller and then reused by launches that are
 *          ordered on one stream (the kernel resets it on its way out);  x_host: peer exchange of a data-parallel job or NULL;
 *   num_pos[B] int32 optional;  ws: gssd_workspace_bytes(GSSD_WS_LOSS, ...);  everything else as gssd_mbox_loss.
 * Returns GSSD_ERR_UNSUPPORTED when the shape cannot be co-resident (ask gssd_mbox_fused_supported first, or fall back to
 * gssd_mbox_match + gssd_mbox_loss: same results bit for bit). */
GSSD_API size_t gssd_fused_state_bytes(void);
GSSD_API int gssd_mbox_fused_supported(int B, int P, int C, int g_max);
GSSD_API int gssd_mbox_loss_fused(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                         const float *gt, const int32_t *gt_off, int sum_G, int g_max,
                         float threshold, int negpos_ratio, float var0, float var1,
                         void *state, const gssd_xchg *x_host,
                         float *losses, float *grad_loc, float *grad_conf,
                         uint8_t *pos_mask, uint8_t *neg_mask, int32_t *num_pos,
                         void *ws, size_t ws_bytes, void *stream);

/* Backward helper: grad_loc *= g[0], grad_conf *= g[1] in place (g = upstream gradients of the two
 * scalar losses, device).  Touches no memory when g == (1,1), the `(loss_l+loss_c).backward()` case
 * of train_lesion_multiphase_v2.py:247-248. */
GSSD_API int gssd_mbox_scale_grads(float *grad_loc, size_t n_loc, float *grad_conf, size_t n_conf,
                          const float *g_loc, const float *g_conf, void *stream);

/* ------------------------------------------------------------------------------------------
 * Detect — replaces Detect.forward, layers/functions/detection_pytorch_ver_1point5.py:33-89
 * (and the legacy detection.py:13-62).
 *   loc[B,P,4], conf[B,P,C] (already softmaxed), priors[P,4]
 *   out[B,C,top_k,5] rows (score,xmin,ymin,xmax,ymax) in descending score, zero-padded; class 0
 *   slab is all zero; count[B,C] int32 and keep_idx[B,C,top_k] int32 (prior index, -1 padded) are
 *   optional (may be NULL). */
GSSD_API int gssd_detect(const float *loc, const float *conf, const float *priors, int B, int P, int C,
                int top_k, float conf_thresh, float nms_thresh, float var0, float var1,
                float *out, int32_t *count, int32_t *keep_idx, void *stream);

/* Detect with the model's softmax fused in (ssd_multiphase_custom_group.py:384-390: `detect(loc, softmax(conf), priors)`):
 * conf_logits[B,P,C] are the raw head outputs, the class score is softmax(conf_logits + class_bias)[class], evaluated
 * the way torch's softmax does (row max, exp, sum in class order, IEEE divide).  class_bias_host[C] (host, may be NULL)
 * is an additive per-class logit offset (prior / calibration shift); everything else as gssd_detect. */
GSSD_API int gssd_detect_logits(const float *loc, const float *conf_logits, const float *class_bias_host, const float *priors,
                       int B, int P, int C, int top_k, float conf_thresh, float nms_thresh, float var0, float var1,
                       float *out, int32_t *count, int32_t *keep_idx, void *stream);

/* ------------------------------------------------------------------------------------------
 * L2Norm — replaces L2Norm.forward, layers/modules/l2norm.py:19-23, and its backward.
 *   x[B,Cn,HW] (NCHW), weight[Cn]; y = weight[c] * x / (sqrt(sum_c x^2) + eps)
 * ---------------------------------------------------------------------------------------- */
GSSD_API int gssd_l2norm_fwd(const float *x, const float *weight, int B, int Cn, int HW, float eps,
                    float *y, float *norm /* [B,HW], saved for backward, may be NULL */, void *stream);
/* gx[B,Cn,HW], gw[Cn]; ws: gssd_l2norm_bwd_ws_bytes() bytes of scratch (per-CTA partial sums of gw) */
GSSD_API int gssd_l2norm_bwd(const float *x, const float *weight, const float *norm, const float *gy,
                    int B, int Cn, int HW, float eps, float *gx, float *gw,
                    void *ws, size_t ws_bytes, void *stream);
GSSD_API size_t gssd_l2norm_bwd_ws_bytes(int B, int Cn, int HW);

/* ------------------------------------------------------------------------------------------
 * Source block — replaces the per-source chain of SSD.forward,
 * models/ssd_multiphase_custom_group.py:258-297 (source 1), 300-325 (source 2), 329-372 (sources
 * 3-6) and the heads at 375-380:
 *     grouped conv (groups = 4) -> BN -> ReLU -> [L2Norm, source 1] -> dense 1x1 fuse_X1 ->
 *     bn_fuse_X1 -> ReLU -> loc.k / conf.k 3x3 -> permute(0,2,3,1) -> flatten -> concat
 * as a chain of calls of ONE implicit-GEMM convolution kernel (tcgen05.mma with the accumulators in
 * TMEM, operands staged by TMA; bf16 inputs, fp32 accumulation).
 *
 * Activations travel "pixel-major padded" (PM): bf16 [n_img][H+2][W+2][C], a zero 1-pixel border
 * around every image, channels innermost.  Seen as a matrix X[rows, C] with
 * rows = n_img*(H+2)*(W+2), a 3x3 / stride 1 / pad 1 tap (dy,dx) is the same matrix shifted by
 * dy*(W+2)+dx rows, so every A tile is one 2-D TMA box; outputs are produced for every padded
 * position and the border rows are written as zeros.
 * ---------------------------------------------------------------------------------------- */
typedef struct gssd_conv_desc {
    int32_t n_img, height, width;   /* interior extent H, W */
    int32_t c_in, c_out, groups;    /* c_in/groups must be a multiple of 64 */
    int32_t taps;                   /* 1 (1x1 conv) or 9 (3x3, stride 1, pad 1) */
    int32_t relu;                   /* clamp at 0 after scale/shift */
    const void  *x;                 /* bf16 PM [rows, c_in] */
    const void  *w;                 /* bf16 [c_out, taps*c_in/groups], k = tap*(c_in/groups) + c, tap = ky*3+kx
                                       (layout of gssd_conv_pack_weights) */
    const float *scale;             /* [c_out] or NULL (= 1) */
    const float *shift;             /* [c_out] or NULL (= 0):  y = acc*scale + shift  (conv bias, folded BN) */
    const float *row_ss_in;         /* [rows] or NULL: acc *= 1/(sqrt(row_ss_in[m]) + l2_eps) first — the L2Norm of
                                       the INPUT (l2norm.py:19-23) deferred past the GEMM; its per-channel weight
                                       is folded into `w` by the caller */
    float        l2_eps;
    void        *y;                 /* bf16 PM [rows, c_out] or NULL */
    float       *row_ss_out;        /* [rows] or NULL: sum_c y^2 of each pixel (the consumer's row_ss_in) */
    float       *chan_sum;          /* [2*c_out] or NULL: += per-channel (sum, sum of squares) of acc*scale+shift over
                                       the interior pixels — train-mode BatchNorm statistics; the caller zeroes it */
    /* head mode (loc != NULL, y == NULL): output columns [0, 4A) go to loc and [4A, 4A + A*n_cls) to
       conf at prior index prior_off + (h*W + w)*A + a — the NHWC flatten + concat of GSSD:375-380 */
    float       *loc;               /* fp32 [n_img, n_priors, 4] */
    float       *conf;              /* fp32 [n_img, n_priors, n_cls] */
    int32_t      n_anchor, n_cls, prior_off, n_priors;
} gssd_conv_desc;

/* one convolution of the chain.  Returns GSSD_ERR_ARG / GSSD_ERR_LIMIT for shapes the kernel does not take. */
GSSD_API int gssd_conv_igemm(const gssd_conv_desc *d_host, void *stream);

/* weight packing: w[c_out, c_in/groups, kh, kw] fp32 (nn.Conv2d.weight) -> bf16 [c_out, taps*c_in/groups];
 * in_scale[c_in] (optional) multiplies input channel c of every filter (L2Norm.weight folding). */
GSSD_API int gssd_conv_pack_weights(const float *w, int c_out, int c_in_per_group, int groups, int taps,
                           const float *in_scale, void *out_bf16, void *stream);
/* layout converters between torch's NCHW fp32 and PM bf16 (borders written as zeros) */
GSSD_API int gssd_nchw_to_pm(const float *x, int n_img, int c, int h, int w, void *y_bf16, void *stream);
GSSD_API int gssd_pm_to_nchw(const void *x_bf16, int n_img, int c, int h, int w, float *y, void *stream);
/* nn.MaxPool2d on a PM tensor (the pools between the grouped backbone convs, ssd_multiphase_custom_group.py:437-446):
 * y PM [n_img, out_h, out_w, c]; out_h / out_w (host, optional) receive the output extent; with x == y == NULL the call
 * only computes them. */
GSSD_API int gssd_maxpool_pm(const void *x_bf16, int n_img, int c, int h, int w, int kernel, int stride, int pad, int ceil_mode,
                    void *y_bf16, int *out_h_host, int *out_w_host, void *stream);
/* train-mode BatchNorm (+ReLU) applied in place on a PM tensor from the statistics a conv call left in
 * chan_sum (biased variance, as F.batch_norm normalises): y = relu?((y - mean)*rstd*gamma + beta) on interior
 * pixels; row_ss_out (optional) receives sum_c y^2 per pixel; mean_var_out[2*c] (optional) the batch mean and
 * UNBIASED variance for the caller's running-stat update. */
GSSD_API int gssd_bn_act_pm(void *y_bf16, int n_img, int c, int h, int w, const float *chan_sum,
                   const float *gamma, const float *beta, float bn_eps, int relu,
                   float *row_ss_out, float *mean_var_out, void *stream);

/* the same out of place: y (the raw conv output) stays intact for the backward, the activations go to y_out */
GSSD_API int gssd_bn_act_pm_to(const void *y_bf16, void *y_out_bf16, int n_img, int c, int h, int w, const float *chan_sum,
                      const float *gamma, const float *beta, float bn_eps, int relu,
                      float *row_ss_out, float *mean_var_out, void *stream);

/* ---- backward of the source block (autograd through ssd_multiphase_custom_group.py:258-380) --------------------------
 * Data gradients need no entry point of their own: the gradient of a stride-1 "same" convolution w.r.t. its input is such
 * a convolution of dY with the filter rotated by 180 degrees and the channel roles swapped inside each group, i.e.
 * gssd_conv_igemm on re-packed weights.
 *
 * Weight gradient: dw[tap][co][ci] = sum over pixels of dY[pix][co] * X[pix + offset(tap)][ci]  (tcgen05, split over the
 * pixels, fp32).  dy: PM bf16 [rows, dy_channels] (dy_channels >= c_out, multiple of 64; the columns beyond c_out are
 * ignored), x: PM bf16 [rows, c_in]; dw: fp32 [taps][c_out_pad][c_in/groups] with c_out_pad = c_out rounded up to 128
 * (gssd_conv_wgrad_bytes), zeroed and filled by the call.  c_in/groups must be a multiple of 128. */
GSSD_API int gssd_conv_wgrad(const void *dy_bf16, const void *x_bf16, int n_img, int height, int width, int c_in, int c_out,
                    int dy_channels, int groups, int taps, float *dw, void *stream);
GSSD_API size_t gssd_conv_wgrad_bytes(int c_in, int c_out, int groups, int taps);
/* upstream gradients of one source's slice of loc[B,P,4] / conf[B,P,C] -> PM bf16 [rows, c_pad] with channels
 * [loc 4A | conf A*C | zeros] (the head's output channels; inverse of the head-mode scatter of gssd_conv_igemm) and their
 * per-channel sums bias_grad[A*(4+C)] (optional) = the gradients of loc.k.bias / conf.k.bias */
GSSD_API int gssd_head_grad_pm(const float *d_loc, const float *d_conf, int n_img, int height, int width, int n_priors, int prior_off,
                      int n_anchor, int n_cls, int c_pad, void *out_bf16, float *bias_grad, void *stream);
/* ReLU + BatchNorm (batch statistics) backward of one stage on PM tensors, with the L2Norm backward of its consumer folded in:
 *   g   = dy [- y * (sum_c dy*y) / ((n+eps)*n), n = sqrt(row_ss_l2)]  [+ add]      then  g = 0 where y <= 0 (relu)
 *   dx  = gamma*rstd*(g - mean(g) - x_hat*mean(g*x_hat))   with chan_sum (the forward's sum / sum of squares; x_hat from yraw)
 *       = g*gamma                                            without (eval-mode BN folded into gamma, or no BN: gamma NULL)
 *   out = dx [* 1/(sqrt(row_ss_out)+eps_out)]
 *   sums[3c] = (sum g = dbeta, sum g*x_hat = dgamma, sum dx = gradient of the conv bias); for an eval-mode BatchNorm that the
 *   forward folded into the conv epilogue, eval_bn_weight / eval_bn_bias (optional) give dgamma / dbeta with
 *   x_hat = (y - beta)/gamma taken from the activations (wherever y > 0; elsewhere g = 0) */
GSSD_API int gssd_bn_relu_bwd_pm(const void *dy_bf16, const void *y_bf16, const void *yraw_bf16, const void *add_bf16, int n_img, int c,
                        int height, int width, const float *chan_sum, const float *gamma, float bn_eps, int relu,
                        const float *eval_bn_weight, const float *eval_bn_bias,
                        const float *row_ss_l2, float l2_eps, const float *row_ss_out, float l2_eps_out,
                        void *out_bf16, float *sums, void *stream);

/* ------------------------------------------------------------------------------------------
 * Training-mode BatchNorm2d (+ ReLU) on NCHW fp32 — the "-> BN -> ReLU" behind every grouped backbone convolution that stays a
 * torch convolution (models/ssd_multiphase_custom_group.py:434-460, applied at :254-259 and :300-301): nn.BatchNorm2d in training
 * mode (batch statistics, biased variance for the normalisation, running statistics updated with `momentum` and the UNBIASED
 * variance) followed by F.relu, as two streaming kernels forward and two backward.
 *   x, y, dy, dx: [N, C, HW] contiguous; gamma / beta [C]; save_mean_rstd [2C] (mean, 1/sqrt(var+eps) per channel, written by the
 *   forward, read by the backward); running_mean / running_var [C] or both NULL; ws: 2C doubles of scratch (zeroed by the call).
 *   mean_shift [C] or NULL: added to the batch mean in the running-mean update only — the bias of the convolution in front when the
 *   caller ran that convolution without it (a per-channel constant cancels in the normalisation: BN(x + b) == BN(x)).
 *   backward: the ReLU mask is recomputed from x (y > 0  <=>  x*a + b > 0 with the forward's own coefficients).
 * ---------------------------------------------------------------------------------------- */
GSSD_API int gssd_bn_relu_nchw_fwd(const float *x, const float *gamma, const float *beta, int N, int C, int HW, float eps, int relu,
                          float *y, float *save_mean_rstd, float *running_mean, float *running_var, float momentum,
                          const float *mean_shift, double *ws, void *stream);
GSSD_API int gssd_bn_relu_nchw_bwd(const float *x, const float *dy, const float *gamma, const float *beta, const float *save_mean_rstd,
                          int N, int C, int HW, int relu, float *dx, float *d_gamma, float *d_beta, double *ws, void *stream);

/* nn.MaxPool2d backward on NCHW fp32 (the pools between the grouped backbone convolutions, ssd_multiphase_custom_group.py:437-446;
 * dilation 1): indices[planes, OH, OW] int64 = flat h*W + w of every window's maximum, as F.max_pool2d(..., return_indices=True)
 * returns them; dx[planes, H, W] is written completely (a gather, no atomics). */
GSSD_API int gssd_maxpool_nchw_bwd(const float *dy, const int64_t *indices, int planes, int H, int W, int OH, int OW, int kernel, int stride,
                          int pad, float *dx, void *stream);

/* The same three operators on channels-last tensors (torch.channels_last: [rows = N*H*W, C] with the channels innermost), for a
 * backbone whose convolutions run on cuDNN's NHWC kernels.  C must be 4 times a power of two, at most 1024. */
GSSD_API int gssd_bn_relu_nhwc_fwd(const float *x, const float *gamma, const float *beta, long rows, int C, float eps, int relu, float *y,
                          float *save_mean_rstd, float *running_mean, float *running_var, float momentum, const float *mean_shift,
                          double *ws, void *stream);
GSSD_API int gssd_bn_relu_nhwc_bwd(const float *x, const float *dy, const float *gamma, const float *beta, const float *save_mean_rstd,
                          long rows, int C, int relu, float *dx, float *d_gamma, float *d_beta, double *ws, void *stream);
GSSD_API int gssd_maxpool_nhwc_bwd(const float *dy, const int64_t *indices, int N, int C, int H, int W, int OH, int OW, int kernel, int stride,
                          int pad, float *dx, void *stream);

/* ------------------------------------------------------------------------------------------
 * Modulated deformable convolution (DCNv2) of GSSD++ — replaces `dcn_v2._DCNv2.apply` as called at
 * layers/dcn_v2_custom.py:49-55 and 84-88 (3x3, stride 1, padding 1, dilation 1).
 *   forward : gssd_dcn_columns, then gssd_conv_igemm with taps = 1 and c_in = 9*c_in on the columns
 *   backward: gssd_conv_igemm (d_columns = dY * W), gssd_conv_wgrad (dW = dY^T * columns), gssd_dcn_columns_bwd
 * x: PM bf16 [rows, c_in]; offset fp32 [n_img, 2*dg*9, H, W] with channel (g*9 + tap)*2 = dy, +1 = dx; mask fp32
 * [n_img, dg*9, H, W] (already sigmoid-ed by the caller, dcn_v2_custom.py:83); dg = deformable groups; c_in/dg must be a
 * multiple of 8.  columns: bf16 [rows, 9*c_in], k = tap*c_in + c — the K order of gssd_conv_pack_weights(taps = 9,
 * groups = 1), so the packed 3x3 filter IS the 1x1 filter over the columns; border rows are written as zeros.
 * A sample at (y, x) is zero unless -1 < y < H and -1 < x < W; neighbours outside the image count as zero.
 * ---------------------------------------------------------------------------------------- */
GSSD_API int gssd_dcn_columns(const void *x_bf16, const float *offset, const float *mask, int n_img, int c_in, int height, int width,
                     int deformable_groups, void *col_bf16, void *stream);
/* d_columns bf16 [rows, 9*c_in] -> dx_pm fp32 PM [rows, c_in] (zeroed by the call, then accumulated with vector reductions:
 * the summation order, hence the last bits, vary from run to run), d_offset / d_mask in the layouts of offset / mask */
GSSD_API int gssd_dcn_columns_bwd(const void *x_bf16, const float *offset, const float *mask, const void *dcol_bf16, int n_img, int c_in,
                         int height, int width, int deformable_groups, float *dx_pm, float *d_offset, float *d_mask, void *stream);
/* fp32 PM [rows, c] -> NCHW fp32 (the interior pixels) */
GSSD_API int gssd_pmf32_to_nchw(const float *x_pm, int n_img, int c, int h, int w, float *y, void *stream);

/* ------------------------------------------------------------------------------------------
 * Attention core of GSSD++'s Self_Attn — replaces layers/self_attn.py:69-81 (permute + bmm + softmax + permute + bmm):
 *   attn[b, n, m]   = softmax_m( sum_d theta[b, d, n] * phi[b, d, m] )        theta [B, D, N], phi [B, D, M]
 *   attn_g[b, c, n] = sum_m g[b, c, m] * attn[b, n, m]                         g [B, Cv, M]
 * fp32, contiguous, the layouts the module's .view() calls produce (self_attn.py:62, 68, 77).  D and Cv must be multiples of 32;
 * a strip of queries against all keys has to fit in shared memory (M, N up to ~7000).  One kernel forward, two backward.
 * Backward: gradients of theta / phi / g from d_attn_g [B, Cv, N] and the saved attn; ds_ws [B, N, M] floats of scratch.  The
 * attention map is an output for inspection only (self_attn.py:86): no gradient is taken through it.
 * ---------------------------------------------------------------------------------------- */
GSSD_API int gssd_attn_fwd(const float *theta, const float *phi, const float *g, int B, int D, int Cv, int N, int M, float *attn,
                  float *attn_g, void *stream);
GSSD_API int gssd_attn_bwd(const float *theta, const float *phi, const float *g, const float *attn, const float *d_attn_g, int B, int D,
                  int Cv, int N, int M, float *d_theta, float *d_phi, float *d_g, float *ds_ws, void *stream);

/* ------------------------------------------------------------------------------------------
 * Host-buffer pipeline — the training-step / inference call of the hot path with HOST inputs and outputs:
 * the H2D of train_lesion_multiphase_v2.py:198-200, the criterion call at :246 (MultiBoxLoss forward + the
 * gradients of its two outputs) and Detect (ssd_multiphase_custom_group.py:384-390), `depth` steps in flight:
 * step i's host->device copies run beside step i-1's kernels and step i-2's device->host copies.
 * The caller owns the device arena (gssd_pipe_arena_bytes); the pipeline creates only streams and events.
 * Host pointers should be page-locked (pageable memory works, the copies then serialise) and must stay valid
 * until gssd_pipe_wait() of the ticket returns.
 * ---------------------------------------------------------------------------------------- */
typedef struct gssd_pipe gssd_pipe;
typedef struct gssd_pipe_cfg {
    int32_t B, P, C, top_k, max_gt_rows, depth;       /* depth in [1, 8] */
    float   match_thresh, var0, var1, conf_thresh, nms_thresh;
    int32_t negpos_ratio;
} gssd_pipe_cfg;
typedef struct gssd_pipe_slot {                       /* device buffers of one in-flight step (inside the arena) */
    float *loc, *conf, *scores, *gt;
    int32_t *gt_off;
    uint16_t *tags;
    void *stats;                                      /* gssd_stats_bytes(B): header + num_pos[B] */
    float *losses, *grad_loc, *grad_conf, *detect_out;
    void *ws;
    size_t ws_bytes;
    void *fused_state;                                /* gssd_fused_state_bytes(): rendezvous state of the one-launch loss */
} gssd_pipe_slot;

GSSD_API size_t  gssd_pipe_arena_bytes(const gssd_pipe_cfg *cfg_host);
GSSD_API int     gssd_pipe_create(gssd_pipe **out, const gssd_pipe_cfg *cfg_host, const float *priors /* device [P,4] */,
                         void *arena /* device */, size_t arena_bytes);
GSSD_API void    gssd_pipe_destroy(gssd_pipe *p);
GSSD_API int     gssd_pipe_slot_info(const gssd_pipe *p, int slot, gssd_pipe_slot *out_host);
/* Enqueue one step.  Returns its ticket (>= 0; slot = ticket % depth) or a negative GSSD_ERR_* / -(1000 + cudaError_t).
 * Blocks only when all `depth` slots are in flight (it then waits for the oldest step).  gt_host[sum_g,5] /
 * gt_off_host[B+1] are the packed ground truth.  losses_host[2], detect_out_host[B,C,top_k,5].
 * The row offsets are checked on the host before anything is enqueued: gt_off_host[0] == 0, gt_off_host[B] == sum_g, every
 * image has between 1 and g_max rows (GSSD_ERR_EMPTY for an image without ground truth — the reference raises IndexError,
 * box_utils.py:94 — GSSD_ERR_ARG otherwise); GSSD_ERR_ARG too when the slot's previous gssd_pipe_begin was never finished.
 * gt_host == NULL makes it a Detect-only step (inference: H2D of loc / conf, Detect, D2H of the detections; no matching, no
 * loss, losses_host may be NULL); detect_out_host == NULL a loss-only step. */
GSSD_API int64_t gssd_pipe_submit(gssd_pipe *p, const float *loc_host, const float *conf_host, const float *scores_host,
                         const float *gt_host, const int32_t *gt_off_host, int sum_g, int g_max,
                         float *losses_host, float *detect_out_host);
/* The same step in two halves for data-parallel jobs: begin = H2D + matching (+ Detect and its D2H); the caller then
 * all-gathers the 16-byte headers (slot.stats) of every rank on *stream_out; finish = loss + D2H of the losses. */
GSSD_API int64_t gssd_pipe_begin(gssd_pipe *p, const float *loc_host, const float *conf_host, const float *scores_host,
                        const float *gt_host, const int32_t *gt_off_host, int sum_g, int g_max,
                        float *detect_out_host, void **stream_out);
GSSD_API int     gssd_pipe_finish(gssd_pipe *p, int64_t ticket, const gssd_loss_stats *global_stats /* device */, int n_global_stats,
                         float *losses_host);
/* Detect reads the conf logits of the step (softmax fused, gssd_detect_logits) instead of a separate scores tensor:
 * scores_host of submit/begin is then ignored and may be NULL (25 % fewer H2D bytes at C = 2) */
GSSD_API int     gssd_pipe_set_detect_logits(gssd_pipe *p, int enable, const float *class_bias_host /* [C] or NULL */);
/* data-parallel: route the statistics through the peer exchange, so gssd_pipe_submit() serves world_size > 1 too */
GSSD_API int     gssd_pipe_set_xchg(gssd_pipe *p, const gssd_xchg *x_host);
/* block until the step's outputs are in the host buffers; its gradients (slot.grad_loc / grad_conf) stay valid until
 * the slot is reused, `depth` submits later */
GSSD_API int     gssd_pipe_wait(gssd_pipe *p, int64_t ticket);

/* ------------------------------------------------------------------------------------------
 * The consumer of Detect's output in the reference's evaluator (test_ap_iobb.py), on the GPU
 * ---------------------------------------------------------------------------------------- */
/* test_ap_iobb.py:126-149 for a whole batch: rows of class `class_index` with score > 0 and score > thresh, boxes scaled by
 * (width, height, width, height), image id (first_image_id + b) in front -> rows[n,6] = (id, score, x1, y1, x2, y2) in (image,
 * descending score) order; offsets[B+1] = first row of every image, offsets[B] = n.  rows must hold B*top_k*6 floats. */
GSSD_API int gssd_collect_detections(const float *detect_out, int B, int C, int top_k, int class_index, float width, float height,
                            float thresh, int first_image_id, float *rows, int32_t *offsets, int32_t *counts_ws /* [B] */,
                            void *stream);
/* test_ap_iobb.py:251-326 + voc_ap (10-41): AP at the IoU thresholds and at the IoBB thresholds (IoBB = intersection over the
 * DETECTION's area) of rows[n_det,6] grouped by image (det_off[n_img+1]) against gt_boxes[sum_G,4] (gt_off[n_img+1], at most
 * 128 boxes per image) in the rows' coordinates.  thresholds[n_iou + n_iobb] (IoU ones first), npos = number of ground-truth
 * boxes, rec_points[11] = np.arange(0., 1.1, 0.1) for the VOC-07 metric.  ap_out[n_iou + n_iobb] float64.  Optional outputs:
 * tp_out[n_thr][n_det] (1 = true positive, 2 = false positive, 0 = image without boxes) and order_out[n_det] (row indices in
 * descending score, equal scores in row order).  Float64 arithmetic as in the reference. */
GSSD_API size_t gssd_ap_workspace_bytes(int n_det, int n_thr);
GSSD_API int gssd_ap_eval(const float *rows, const int32_t *det_off, const float *gt_boxes, const int32_t *gt_off, int n_img, int n_det,
                 const double *thresholds, int n_iou, int n_iobb, int npos, int use_07_metric, const double *rec_points,
                 double *ap_out, uint8_t *tp_out, uint32_t *order_out, void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * workspace sizing (host-only)
 * ---------------------------------------------------------------------------------------- */
enum { GSSD_WS_LSE = 0, GSSD_WS_MATCH = 1, GSSD_WS_LOSS = 2, GSSD_WS_NMS = 3 };
GSSD_API size_t gssd_workspace_bytes(int kind, int B, int P, int C, int sum_G, int top_k);

#ifdef __cplusplus
}
#endif
#endif /* GSSD_H_ */
