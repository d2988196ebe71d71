/* sequitr_b200 -- C ABI of the B200-native Sequitr hot path.
 *
 * The reference (quantumjot/sequitr) is pure Python with no FFI; its "operator
 * API" for this path is three Python seams.  This header is
 * the boundary a maintainer binds those seams to (see INTEGRATION.md for the
 * ctypes stubs).  Each entry point cites the reference interface it replaces,
 * relative to /root/reference/sequitr/.
 *
 * Conventions
 *  - every call returns an int status: SQ_OK (0) or a negative SQ_E* code;
 *    sq_last_error() returns a human-readable message for the calling thread.
 *    Nothing throws across the boundary.
 *  - plain pointers and sizes only.  Pointers named *_dev are caller-owned
 *    DEVICE pointers (e.g. torch tensors' data_ptr()); *_host are host pointers.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); device
 *    entry points are asynchronous on that stream.
 *  - scratch memory is caller-owned: ask sq_*_workspace_bytes(), allocate once,
 *    pass it in.  The library does not allocate device memory after plan
 *    creation (the *_host convenience calls own a pinned staging arena that is
 *    grown on first use only).
 *  - one handle per device; a handle is not thread-safe, distinct handles are
 *    independent (one process per GPU, frames sharded by the caller).
 */
#ifndef SEQUITR_B200_H
#define SEQUITR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_OK            0
#define SQ_EINVAL       -1   /* bad argument (ValueError on the Python side)           */
#define SQ_ECUDA        -2   /* CUDA runtime/driver error                               */
#define SQ_ENOMEM       -3   /* workspace too small / allocation failed                 */
#define SQ_EOVERFLOW    -4   /* more components than max_rows                           */
#define SQ_EUNSUPPORTED -5   /* valid request this build cannot serve (e.g. not sm_100) */
#define SQ_ESTATE       -6   /* plan not finalised / weights missing                    */

typedef struct sq_handle_s *sq_handle_t;
typedef struct sq_unet_s   *sq_unet_t;
typedef struct sq_trainer_s *sq_trainer_t;

/* bridge types, networks/unet.py:42 BRIDGE_TYPES */
#define SQ_BRIDGE_NONE   0
#define SQ_BRIDGE_ADD    1
#define SQ_BRIDGE_MUL    2
#define SQ_BRIDGE_SUB    3
#define SQ_BRIDGE_CONCAT 4

/* arithmetic contract of a UNet plan */
#define SQ_MODE_FP32_EXACT 0  /* CUDA cores, fixed fmaf order: bit-identical to oracle/unet_ref.c */
#define SQ_MODE_BF16_TC    1  /* tcgen05 tensor cores: bf16 storage, fp32 accumulation in TMEM    */

/* output dtype of the weight maps */
#define SQ_F32 0
#define SQ_F64 1
#define SQ_U8  2   /* raw camera frames (dataio/octopus.py:236: 'uint' + bit depth) */
#define SQ_U16 3

/* ------------------------------------------------------------------ runtime */
const char *sq_version(void);
const char *sq_last_error(void);
int sq_create(int device, sq_handle_t *out);
int sq_destroy(sq_handle_t h);
/* sm_count, compute capability major/minor, total device memory */
int sq_device_info(sq_handle_t h, int *sm_count, int *cc_major, int *cc_minor, size_t *total_mem);

/* Page-lock / release a host buffer the caller owns (cudaHostRegister), e.g. a frame stack read by a dataio reader
 * (dataio/octopus.py:231-245), so that the *_host entry points copy it by DMA.  A refusal by the OS (transient right
 * after another process released a large registration, or RLIMIT_MEMLOCK) returns SQ_ECUDA with the CUDA error state
 * CLEARED: the caller simply goes on with pageable memory. */
int sq_host_register(void *ptr, size_t bytes);
int sq_host_unregister(void *ptr);

/* ------------------------------------------------ label-and-localise (L1)
 * Replaces the per-frame / per-class loop of utils.CentroidWriter.write
 * (utils.py:531-566: scipy.ndimage.label :547 + center_of_mass :550).
 *
 * mask_dev  : uint8 (n, d, hgt, wid); d = 1 for planar stacks (N,H,W).  For
 *             volumes the caller passes the stack already swapped to
 *             (N,Y,X,Z) like utils.py:519, i.e. (d,hgt,wid) = (Y,X,Z).
 * Components are sets of equal non-zero value connected through faces
 * (4-connected planar / 6-connected volumetric = SciPy's default structure).
 * table_dev : float32 (n, max_rows, 5), rows [frame0+i, c0, c1, c2, class]
 *             (c2 = 0 for planar), ordered by ascending class then by the
 *             raster position of each component's first voxel (utils.py:559-574).
 * counts_dev: int32 (n) number of components found in each frame; if a count
 *             exceeds max_rows only the first max_rows rows of that frame are
 *             valid (the *_host variant returns SQ_EOVERFLOW).
 * labels_dev: optional int32 (n,d,hgt,wid): SciPy's label number of each voxel
 *             within its class (0 = background); NULL to skip (the reference
 *             never stores it).
 */
int sq_label_workspace_bytes(sq_handle_t h, int n, int d, int hgt, int wid, int max_rows,
                             size_t *bytes);
int sq_label_centroids(sq_handle_t h, const uint8_t *mask_dev, int n, int d, int hgt, int wid,
                       int frame0, int32_t *labels_dev, float *table_dev, int32_t *counts_dev,
                       int max_rows, void *workspace_dev, size_t workspace_bytes, void *stream);
/* host buffers in, host buffers out (H2D + kernels + D2H, synchronous) */
int sq_label_centroids_host(sq_handle_t h, const uint8_t *mask_host, int n, int d, int hgt,
                            int wid, int frame0, int32_t *labels_host, float *table_host,
                            int32_t *counts_host, int max_rows);

/* ------------------------------------------------------ weight maps (W1, W3)
 * sq_weightmap_edt replaces pipeline.ImageWeightMap.pipe (pipeline.py:475-479):
 *   d = EDT(1 - m);  w = w0*(1-m)*exp(-d*d/(2 sigma^2 + 1e-99)) + m + 1
 * for a {0,1} mask m (non-zero = foreground).  The distance transform is the
 * exact Euclidean one (integer squared distances, bit-identical to SciPy's);
 * a frame with no foreground follows SciPy's behaviour (distance to a virtual
 * seed at (-1,0)).  Arithmetic is fp64 like the reference; out_dtype selects
 * the stored type (SQ_F64 = the reference's dtype, SQ_F32 = what weightmap.py:205
 * saves).  d2_dev: optional int32 (n,hgt,wid) exact squared distances.
 *
 * sq_weightmap_unet is the north-star d1+d2 map on int32 instance labels:
 *   w = wc[m] + w0*(1-m)*exp(-(d1+d2)^2/(2 sigma^2 + 1e-99)),  m = labels > 0,
 * d1/d2 = exact distances to the nearest / second-nearest distinct instance;
 * wc = {1, 2} when NULL (the reference's "+ 1 + m" class term).
 */
int sq_weightmap_workspace_bytes(sq_handle_t h, int n, int hgt, int wid, int instance_mode,
                                 size_t *bytes);
int sq_weightmap_edt(sq_handle_t h, const uint8_t *mask_dev, int n, int hgt, int wid,
                     double w0, double sigma, int out_dtype, void *out_dev, int32_t *d2_dev,
                     void *workspace_dev, size_t workspace_bytes, void *stream);
int sq_weightmap_unet(sq_handle_t h, const int32_t *labels_dev, int n, int hgt, int wid,
                      double w0, double sigma, const double *wc_host, int out_dtype,
                      void *out_dev, void *workspace_dev, size_t workspace_bytes, void *stream);
int sq_weightmap_edt_host(sq_handle_t h, const uint8_t *mask_host, int n, int hgt, int wid,
                          double w0, double sigma, int out_dtype, void *out_host,
                          int32_t *d2_host);
int sq_weightmap_unet_host(sq_handle_t h, const int32_t *labels_host, int n, int hgt, int wid,
                           double w0, double sigma, const double *wc_host, int out_dtype,
                           void *out_host);

/* ---------------------------------------------------------------- UNet (U1-U8)
 * A plan is the device-side image of networks/unet.py's UNet: __init__ (:126-146)
 * -> sq_unet_create, variables -> sq_unet_load_weights, build (:224-262) ->
 * sq_unet_forward (+ the softmax/argmax head the consumer utils.py:492 implies).
 *
 * ndim 2: features (n,hgt,wid,cin) NHWC, d must be 1.  ndim 3: (n,d,hgt,wid,cin).
 * filters[nlev]: networks/unet.py:40 DEFAULT_FILTERS = (16,32,64,128,256).
 * Layer definitions (the reference leaves them abstract, unet.py:326-342):
 *   conv_layer          3x3(x3) SAME conv + bias (+ per-channel affine) + ReLU
 *   max_pool_layer      2x2(x2) max pool, stride 2
 *   conv_transpose_layer 2x2(x2) stride-2 transposed conv + bias
 *   conv_layer_1x1      1x1 conv + bias -> logits
 * Weights are HOST float32 arrays named by TF scope, TF layouts:
 *   "UNet/down{i}/conv{1,2}/kernel" (3,3[,3],Cin,Cout)  ".../bias" (Cout)
 *   optional ".../scale", ".../shift" (Cout): y = relu((conv+bias)*scale + shift)
 *   "UNet/up{i}/upscale/kernel" (2,2[,2],Cout,Cin)  ".../bias"
 *   "UNet/up{i}/conv{1,2}/..."; "UNet/to_image/kernel" (1,1[,1],Cin,K), ".../bias"
 */
int sq_unet_create(sq_handle_t h, int ndim, int num_inputs, int num_outputs,
                   const int *filters, int nlev, int bridge, int mode, sq_unet_t *out);
int sq_unet_destroy(sq_unet_t u);
int sq_unet_load_weights(sq_unet_t u, const char *name, const float *data_host,
                         const int64_t *shape, int rank);
int sq_unet_finalize(sq_unet_t u);            /* checks completeness, uploads, re-lays-out */
int sq_unet_workspace_bytes(sq_unet_t u, int n, int d, int hgt, int wid, size_t *bytes);
/* in_dev float32 channels-last; any of probs/mask/logits may be NULL.
 * probs/logits float32 (n,[d,]hgt,wid,K); mask uint8 (n,[d,]hgt,wid) = argmax (first max). */
int sq_unet_forward(sq_unet_t u, const float *in_dev, int n, int d, int hgt, int wid,
                    float *probs_dev, uint8_t *mask_dev, float *logits_dev,
                    void *workspace_dev, size_t workspace_bytes, void *stream);
/* number of kernel launches the last sq_unet_forward issued (for bench.py's gpu_launches) */
int sq_unet_last_launches(sq_unet_t u, int *launches);
/* per-layer device times (ms) of one instrumented forward: names_out receives
 * up to max_layers pointers to static strings, ms_out/flops_out the times and
 * the algorithmic FLOPs of each layer; returns the layer count in *n_layers. */
int sq_unet_profile(sq_unet_t u, const float *in_dev, int n, int d, int hgt, int wid,
                    void *workspace_dev, size_t workspace_bytes, void *stream,
                    const char **names_out, float *ms_out, double *flops_out,
                    int max_layers, int *n_layers);

/* Pre-inference clean-up pipes (the step right before the UNet).  Stacks
 * are float32 (n,h,w,c) channels-last on the device; workspace: sq_prep_workspace_bytes(n, c).
 *   sq_image_norm        replaces ImageNorm.pipe        (reference pipeline.py:338-356): per image and
 *                        channel (x - mean) / std, population std, float32 arithmetic on float32 moments
 *                        (the reference's 1e-99 epsilon vanishes in float32 too); in may equal out.
 *   sq_image_outliers    replaces ImageOutliers.pipe    (pipeline.py:266-296): size x size median
 *                        (scipy.ndimage.median_filter: 'reflect' boundary, window origin size/2, rank
 *                        size*size/2); pixels with |x - median| > threshold take the median.  size 1..5;
 *                        in and out must differ.  Bit-exact.
 *   sq_image_bgsubtract  replaces ImageBGSubtract.pipe  (pipeline.py:360-405): least-squares quadratic
 *                        surface over (column, row) subtracted from a single-channel image; fp64 fit;
 *                        out_dtype SQ_F64 (the reference's result dtype) or SQ_F32.
 *   sq_image_pipe_host   the same three on host buffers (which = 0 norm, 1 outliers, 2 background). */
int sq_prep_workspace_bytes(sq_handle_t h, int n, int c, size_t *bytes);
int sq_image_norm(sq_handle_t h, const float *in_dev, float *out_dev, int n, int hgt, int wid, int c,
                  void *workspace_dev, size_t workspace_bytes, void *stream);
/* sq_image_norm on a stack stored as SQ_F32, SQ_U16 or SQ_U8 (raw camera counts are widened on load). */
int sq_image_norm_raw(sq_handle_t h, const void *in_dev, int in_dtype, float *out_dev, int n, int hgt, int wid,
                      int c, void *workspace_dev, size_t workspace_bytes, void *stream);
int sq_image_outliers(sq_handle_t h, const float *in_dev, float *out_dev, int n, int hgt, int wid, int c,
                      int size, double threshold, void *stream);
int sq_image_bgsubtract(sq_handle_t h, const float *in_dev, void *out_dev, int out_dtype, int n, int hgt,
                        int wid, void *workspace_dev, size_t workspace_bytes, void *stream);
int sq_image_pipe_host(sq_handle_t h, int which, const float *in_host, void *out_host, int out_dtype,
                       int n, int hgt, int wid, int c, int size, double threshold);

/* Weighted softmax cross-entropy of the head (training step of BASELINE config 5; consumes the
 * per-pixel 'weights' map and the labels that networks/unet.py tr_augment :396-401 returns; the
 * loss itself is not shipped by the reference).  logits_dev float32 (npix,K); labels_dev uint8
 * (npix) class ids < K; weights_dev float32 (npix).
 *   loss  = (1/npix) * sum_i w_i * (logsumexp(logits_i) - logits_i[label_i])      -> *loss_dev (float64)
 *   grad  = w_i * (softmax(logits_i) - onehot(label_i)) / npix                    -> grad_dev (npix,K) or NULL
 * The reduction order is fixed (deterministic).  workspace: sq_weighted_ce_workspace_bytes(). */
int sq_weighted_ce_workspace_bytes(sq_handle_t h, size_t *bytes);
int sq_weighted_ce(sq_handle_t h, const float *logits_dev, const uint8_t *labels_dev,
                   const float *weights_dev, long long npix, int K, double *loss_dev,
                   float *grad_dev, void *workspace_dev, size_t workspace_bytes, void *stream);

/* tr_augment of the training input pipeline (reference networks/unet.py:348-401; config 5): random
 * rotation about the image centre, crop, one-hot label expansion -- fused into one gather kernel, so
 * only the (ch,cw) crop is computed (the reference rotates four full-size tensors with
 * tf.contrib.image.rotate :373-383, then crops :391-393).  Per frame i the caller supplies the
 * projective coefficients transforms_host[6i..6i+5] = (a0,a1,a2,b0,b1,b2) mapping an OUTPUT pixel
 * (x = column, y = row) to the input sample point (a0*x + a1*y + a2, b0*x + b1*y + b2) -- for a
 * rotation by theta TensorFlow's angles_to_projective_transforms gives (cos, -sin, x_off, sin, cos,
 * y_off) -- and the crop origin crop_host[2i..2i+1] = (rh, rw) (:387-388).  image_dev float32
 * (n,hgt,wid,c) is sampled BILINEAR, label_dev uint8 (n,hgt,wid) NEAREST (std::round), weights_dev
 * float32 (n,hgt,wid) BILINEAR plus 1 wherever the NEAREST sample point lies outside the frame (the
 * reference's wgt_mask :380-383); samples outside the frame read 0.  All coordinate arithmetic is
 * float32 without fused multiply-add (TensorFlow 1.x ImageProjectiveTransform).  Outputs:
 * image_out (n,ch,cw,c) float32, label_out (n,ch,cw,num_outputs) uint8 one-hot of classes
 * 0..num_outputs-1 (:396-398, num_outputs <= 5), weights_out (n,ch,cw) float32. */
int sq_tr_augment(sq_handle_t h, const float *image_dev, const uint8_t *label_dev, const float *weights_dev,
                  int n, int hgt, int wid, int c, const float *transforms_host, const int *crop_host,
                  int ch, int cw, int num_outputs, float *image_out_dev, uint8_t *label_out_dev,
                  float *weights_out_dev, void *stream);

/* Training step of the UNet (BASELINE config 5): what a tf.estimator model_fn does with the graph that
 * networks/unet.py `build` :224-262 makes in TRAIN mode (`training` :170-172, dropout after every conv block
 * :274-276) and the (image, {'label', 'weights'}) pairs tr_augment :348-401 yields.  The reference ships no loss
 * and no optimiser for the UNet (its only optimiser is the GAN's tf.train.AdamOptimizer, gan.py:740-751), so the
 * step is: forward with every activation kept -> sq_weighted_ce -> backward through every layer -> in-place update
 * of the plan's kernels and biases with TensorFlow's Adam rule (optimizer 1) or plain SGD (optimizer 0).
 *   - the plan must be SQ_MODE_FP32_EXACT; it serves inference with the updated weights right after a step (read them
 *     with sq_trainer_read to load a bf16 plan).  Kernels and biases are trained; a layer's optional per-channel
 *     affine (scale, shift = folded BN statistics) stays frozen: y = relu((conv + bias) * scale + shift)
 *   - dropout: rate in [0,1); mask = counter-based hash of (seed, step, block, element) -- see train.cu
 *   - image_dev float32 (n,[d,]hgt,wid,cin); labels_dev uint8 class ids (n,[d,]hgt,wid); weights_dev float32 same
 *     shape; *loss_dev float64 on the device; apply_update 0 computes loss and gradients only
 *   - data-parallel training (one process per GPU): every rank runs the step with apply_update 0 on
 *     its share of the batch, the ranks all-reduce (average) the gradient arena -- ONE contiguous float32 buffer,
 *     sq_trainer_grad_arena -- with one NCCL collective, then every rank calls sq_trainer_apply
 *   - sq_trainer_read: name "<scope>/kernel" or "<scope>/bias" (TF layouts, as sq_unet_load_weights takes them),
 *     what 0 = current value, 1 = gradient of the last step */
int sq_trainer_create(sq_unet_t u, int optimizer, float learning_rate, float beta1, float beta2, float epsilon,
                      float dropout, unsigned long long seed, sq_trainer_t *out);
int sq_trainer_destroy(sq_trainer_t t);
int sq_trainer_workspace_bytes(sq_trainer_t t, int n, int d, int hgt, int wid, size_t *bytes);
int sq_trainer_step(sq_trainer_t t, const float *image_dev, const uint8_t *labels_dev, const float *weights_dev,
                    int n, int d, int hgt, int wid, int apply_update, double *loss_dev, void *workspace_dev,
                    size_t workspace_bytes, void *stream);
int sq_trainer_apply(sq_trainer_t t, void *stream);
int sq_trainer_grad_arena(sq_trainer_t t, float **arena_dev, size_t *count);
int sq_trainer_read(sq_trainer_t t, const char *name, int what, float *out_host, size_t count);

/* The whole data-parallel hot path on HOST frames: H2D -> UNet -> argmax ->
 * label-and-localise -> D2H of the centroid tables (the call a Sequitr job
 * function makes per batch of frames).  frames_host float32 (n,hgt,wid,cin);
 * table_host (n,max_rows,5); counts_host (n); mask_host optional uint8 (n,hgt,wid). */
int sq_segment_localise_host(sq_unet_t u, const float *frames_host, int n, int hgt, int wid,
                             int frame0, float *table_host, int32_t *counts_host, int max_rows,
                             uint8_t *mask_host);

/* Same path on RAW camera frames as the reference's readers deliver them (dataio/octopus.py:236-247:
 * uint8 / uint16 memmap, one channel): the frames cross PCIe in their stored type (1 or 2 bytes per
 * pixel instead of 4), are widened to float32 on the device (exact) and, when normalise != 0, pass
 * through ImageNorm (sq_image_norm) before the UNet.  in_dtype: SQ_U8, SQ_U16 or SQ_F32. */
int sq_segment_localise_raw_host(sq_unet_t u, const void *frames_host, int in_dtype, int normalise, int n,
                                 int hgt, int wid, int frame0, float *table_host, int32_t *counts_host,
                                 int max_rows, uint8_t *mask_host);

/* Widen a raw stack to float32 on the device: in_dtype SQ_U8 / SQ_U16 (SQ_F32 copies). */
int sq_image_cast(sq_handle_t h, const void *in_dev, int in_dtype, float *out_dev, long long count,
                  void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SEQUITR_B200_H */
