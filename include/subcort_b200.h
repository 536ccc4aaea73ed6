/* subcort_b200 -- C-ABI of the B200-native voxelwise hot path.
 *
 * The reference (sergivalverde/sub-cortical_segmentation) has NO FFI of its own: its
 * native arithmetic is generated at run time by Theano behind nolearn's NeuralNet
 * (cnn_cort/nets.py:233-246) and its data layer is numpy (cnn_cort/base.py).  Each entry
 * point below therefore names the reference *Python* interface it replaces; the
 * Python-side binding a maintainer adds is the ctypes stub shown in INTEGRATION.md
 * (implemented in sub-cortical_segmentation_b200/cnn_cort/_native.py).
 *
 * Conventions: extern "C"; plain pointers and sizes; every call returns 0 on success or
 * a negative sc_status, with a human-readable message in sc_last_error() (thread-local);
 * no exceptions cross the boundary.  Pointers named *_dev are CUDA device pointers on the
 * context's device (e.g. torch tensors' data_ptr()); *_host are host pointers (pinned
 * memory makes the copies asynchronous).  `stream` is a cudaStream_t passed as void*
 * (NULL = legacy default stream).  One sc_ctx per device per thread; contexts are
 * independent.  There is no CPU fallback: without a CUDA device sc_create fails.
 *
 * Coordinates are (x, y, z) int32 triples, volumes are C-ordered [X][Y][Z] (z fastest),
 * the atlas is [X][Y][Z][15] float32 -- the layouts cnn_cort/base.py uses after nibabel.
 */
#ifndef SUBCORT_B200_H
#define SUBCORT_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define SC_API __attribute__((visibility("default")))
#else
#define SC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sc_ctx sc_ctx;

typedef enum {
  SC_OK = 0,
  SC_ERR_CUDA = -1,        /* a CUDA runtime/driver call failed */
  SC_ERR_ARG = -2,         /* bad argument (null pointer, negative size, unsupported patch size ...) */
  SC_ERR_STATE = -3,       /* call order violated (e.g. forward before sc_load_weights) */
  SC_ERR_NOMEM = -4,       /* device workspace could not be allocated */
  SC_ERR_UNSUPPORTED = -5  /* device is not sm_100 */
} sc_status;

#define SC_NUM_CLASSES 15
#define SC_PATCH 32
#define SC_PARAM_FLOATS 883455   /* floats in nets/<name>/<name>.pkl, pickle order */
#define SC_NUM_ARRAYS 107        /* non-empty arrays in that pickle */

/* ---- context ------------------------------------------------------------------------ */
SC_API int sc_version(void);
SC_API const char* sc_last_error(void);
/* replaces: THEANO_FLAGS device selection, cnn_cort/load_options.py:54-57 */
SC_API int sc_create(int device, sc_ctx** out);
SC_API int sc_destroy(sc_ctx* ctx);
/* knobs: "gemm" = 1 tcgen05 bf16x3 (the product path, always the default: sc_create fails where it cannot be brought up)
 *        | 0 exact-fp32 SIMT kernels (cross-check back-end for the parity tests only);
 *        "chunk_voxels" = voxels per head chunk in sc_segment_volume; "profile" = 0 | 1;
 *        "tc_compact" 0 | 1 (candidate-row compaction of the FC head, default 1), "gather_ctas_per_sm" 0..32,
 *        "tc_timing" = kernel class whose launches record per-role wait cycles (debug). */
SC_API int sc_set_option(sc_ctx* ctx, const char* key, int64_t value);
SC_API int64_t sc_get_counter(sc_ctx* ctx, const char* key); /* "launches": kernels launched so far */
/* per-kernel-class device times: after sc_set_option(ctx, "profile", 1) every launch is bracketed by
 * CUDA events on its stream; sc_profile_read synchronises, returns summed ms and launch counts per
 * class (arrays of at least sc_profile_classes() entries) and resets the accumulators. */
SC_API int sc_profile_classes(void);
SC_API const char* sc_profile_name(int cls);
SC_API int sc_profile_read(sc_ctx* ctx, double* ms_out, int64_t* launches_out, int n);

/* ---- parameters ----------------------------------------------------------------------
 * replaces: nolearn NeuralNet.load_params_from / save_params_to (call site nets.py:251).
 * `blob_host` is the concatenation of the 107 arrays of the OrderedDict in pickle order
 * (SC_PARAM_FLOATS floats).  The library keeps that master copy on the device and derives
 * the inference layouts (filter flip for flip_filters=True, BN folded to scale/shift,
 * transposed / padded / split-bf16 GEMM operands). */
SC_API int sc_load_weights(sc_ctx* ctx, const float* blob_host, int64_t n_floats);
SC_API int sc_get_params(sc_ctx* ctx, float* blob_host, int64_t n_floats);

/* ---- candidate voxels ----------------------------------------------------------------
 * replaces: get_mask_voxels(mask) cnn_cort/base.py:310-331 (np.nonzero order, C-order).
 * elem_bytes: 1 (uint8/bool mask) or 4 (float32/int32 image, tested != 0).
 * Writes up to `capacity` triples; *n_out_host receives the total count (synchronises). */
SC_API int sc_nonzero_coords(sc_ctx* ctx, const void* vol_dev, int elem_bytes, const int32_t dims[3],
                      int32_t* xyz_dev, int64_t capacity, int64_t* n_out_host, void* stream);

/* replaces: scipy.ndimage.binary_dilation(mask, iterations=10) of the crop path (base.py:369): default
 * 6-connected structuring element, zero border.  mask/out uint8 [X][Y][Z] (0 / non-zero in, 0 / 1 out), out != mask. */
SC_API int sc_dilate_mask(sc_ctx* ctx, const uint8_t* mask_dev, const int32_t dims[3], int iterations,
                   uint8_t* out_dev, void* stream);

/* ---- scan preparation on the device ---------------------------------------------------
 * replaces: the host-side head of load_patch_batch / test_scan (base.py:357-372).  A scan is uploaded once in the
 * array order the NIfTI file has; everything up to the label volume stays on the device. */
typedef enum { SC_DT_U8 = 1, SC_DT_I8, SC_DT_U16, SC_DT_I16, SC_DT_U32, SC_DT_I32, SC_DT_F32, SC_DT_F64 } sc_dtype;
/* NIfTI (Fortran) order -> the C order used everywhere else: src is [channels][Z][Y][X] (x fastest, what nibabel's
 * get_data() holds in memory), dst is [X][Y][Z][channels]; bit-preserving for elements of 1, 2, 4 or 8 bytes. */
SC_API int sc_import_volume(sc_ctx* ctx, const void* src_dev, int elem_bytes, const int32_t dims[3], int channels,
                     void* dst_dev, void* stream);
/* Crop mode (base.py:367-369): the atlas priors are only read at the candidate voxels, so only the candidates' bounding
 * box {x0,x1,y0,y1,z0,z1} of the HOST array (94 % of a scan's bytes are priors) is uploaded: strided DMA copies into
 * staging_dev (Fortran-ordered source: >= X * box_y * box_z * channels * elem_bytes bytes -- whole x rows are moved, the
 * x range is cut on the device; NULL for a C-ordered source) and a
 * reorder into the box region of dst_dev, the full-size C-ordered [X][Y][Z][channels] device volume.  Voxels outside
 * the box keep whatever dst_dev held.  src_host should be page-locked (pageable memory makes the copies synchronous). */
SC_API int sc_upload_volume_box(sc_ctx* ctx, const void* src_host, int elem_bytes, const int32_t dims[3], int channels,
                         int fortran_order, const int32_t box[6], void* staging_dev, void* dst_dev, void* stream);
/* replaces: image_norm = (image - image[np.nonzero(image)].mean()) / image[np.nonzero(image)].std() (base.py:358),
 * cast to float32 as the patches are (base.py:383).  numpy's result bit for bit: same dtype promotion, same pairwise
 * summation order (csrc/prep.cu).  vol_dev: C-ordered raw volume of `dtype`; out_dev float32 (NULL: statistics only);
 * mean_std_host (nullable): the two scalars.  Synchronises the stream. */
SC_API int sc_normalise_volume(sc_ctx* ctx, const void* vol_dev, int dtype, const int32_t dims[3], float* out_dev,
                        double* mean_std_host, void* stream);
/* replaces: image.astype('bool') (base.py:372) / the truth value scipy's binary_dilation takes of the registered mask
 * (base.py:369): mask_dev[v] = vol[v] != 0 as uint8. */
SC_API int sc_candidate_mask(sc_ctx* ctx, const void* vol_dev, int dtype, const int32_t dims[3], uint8_t* mask_dev, void* stream);
/* half-open bounding box {x0,x1,y0,y1,z0,z1} and number of the non-zero voxels of a uint8 mask (all zeros when empty):
 * the box sc_segment_volume restricts its convolutions to.  Synchronises the stream. */
SC_API int sc_mask_bbox(sc_ctx* ctx, const uint8_t* mask_dev, const int32_t dims[3], int32_t box_host[6],
                 int64_t* count_host, void* stream);

/* ---- orthogonal patch gather ---------------------------------------------------------
 * replaces: get_patches x3 views (base.py:272-308) + the atlas vector with background fix
 * (base.py:387-394) of one load_patch_batch batch.  Outputs are [n][1][32][32] float32 per
 * view and [n][15] float32; any output pointer may be NULL to skip it.  bg_fix=1 applies
 * "row sums to 0 -> row[14] = 1" (test path); 0 leaves rows untouched (train path, Q4). */
SC_API int sc_gather_patches(sc_ctx* ctx, const float* vol_dev, const int32_t dims[3],
                      const float* atlas_dev, int bg_fix, const int32_t* xyz_dev, int64_t n,
                      float* axial_dev, float* coronal_dev, float* saggital_dev,
                      float* atlas_out_dev, void* stream);
/* the same windows of an integer label volume (load_patch_vectors, base.py:156-160);
 * only the centre pixel is ever consumed (base.py:85), so this returns labels[n] uint8. */
SC_API int sc_gather_center_labels(sc_ctx* ctx, const uint8_t* labels_dev, const int32_t dims[3],
                            const int32_t* xyz_dev, int64_t n, uint8_t* y_dev, void* stream);

/* ---- network forward (patchwise) -----------------------------------------------------
 * replaces: net.predict_proba / net.predict on {'in1','in2','in3','in4'}
 * (call sites base.py:425-428, 435-438).  proba_dev [n][15] and/or label_dev [n] may be
 * NULL.  Deterministic mode (stored BN statistics, no dropout). */
SC_API int sc_forward(sc_ctx* ctx, const float* in1_dev, const float* in2_dev, const float* in3_dev,
               const float* in4_dev, int64_t n, float* proba_dev, int32_t* label_dev, void* stream);
/* host-buffer form of the same call: H2D of the four inputs and D2H of the results inside. */
SC_API int sc_forward_host(sc_ctx* ctx, const float* in1_host, const float* in2_host, const float* in3_host,
                    const float* in4_host, int64_t n, float* proba_host, int32_t* label_host, void* stream);
/* gather + forward without materialising patches on the host: one test_scan batch
 * (base.py:421-428) from a device-resident volume. */
SC_API int sc_forward_from_volume(sc_ctx* ctx, const float* vol_dev, const int32_t dims[3],
                           const float* atlas_dev, const int32_t* xyz_dev, int64_t n,
                           float* proba_dev, int32_t* label_dev, void* stream);

/* one dense layer of the head on its own (parity tests of the GEMM back-ends): which = 0..2 the
 * d1 layer of the axial / coronal / saggital branch (in [n][576] flattened conv5 maps, 540 used -> out [n][192],
 * columns 0..179 valid), 3 = FC1 (in [n][576] = the feature buffer, 192 columns per view of which 180 are used -> out [n][576], columns 0..539 valid), 4 = fc_2
 * (in [n][576] = FC1 output | atlas | zero pad -> out [n][272]).  backend: 0 SIMT fp32, 1 tcgen05 split-bf16 (three MMAs).
 * replaces: DenseLayer + PReLU, cnn_cort/nets.py:179-180, 217-218, 227-228. */
SC_API int sc_dense_layer(sc_ctx* ctx, int which, const float* in_dev, int64_t n, float* out_dev, int backend, void* stream);

/* ---- whole-volume inference (dense dilated formulation) -----------------------------
 * replaces: the body of test_scan (base.py:421-440) when the candidates are (nearly) all
 * voxels of a box: per view the branch runs as dilated convolutions over whole slices,
 * which is exactly the patchwise network evaluated at every pixel.
 * box = {x0,x1,y0,y1,z0,z1} half-open (NULL = whole volume).  cand_mask_dev (uint8
 * [X][Y][Z], NULL = every voxel of the box) selects which voxels are written; the others
 * keep whatever label_vol/proba_vol held.  label_vol_dev uint8 [X][Y][Z];
 * proba_vol_dev float32 [X][Y][Z][15] or NULL. */
SC_API int sc_segment_volume(sc_ctx* ctx, const float* vol_dev, const int32_t dims[3],
                      const float* atlas_dev, const int32_t* box, const uint8_t* cand_mask_dev,
                      uint8_t* label_vol_dev, float* proba_vol_dev, void* stream);
/* The atlas priors are first read after the convolution phase.  A caller that uploads them on a side stream passes the
 * cudaEvent_t recorded behind that upload here; the NEXT sc_segment_volume call makes its stream wait for the event right
 * before the FC head instead of the caller serialising upload and convolutions.  One-shot (cleared by that call). */
SC_API int sc_atlas_ready_event(sc_ctx* ctx, void* cuda_event);
/* host-buffer form: copies volume + atlas (+mask) in and the label (+proba) volume out. */
SC_API int sc_segment_volume_host(sc_ctx* ctx, const float* vol_host, const int32_t dims[3],
                           const float* atlas_host, const int32_t* box, const uint8_t* cand_mask_host,
                           uint8_t* label_vol_host, float* proba_vol_host, void* stream);

/* ---- post-processing ------------------------------------------------------------------
 * replaces: post_process_segmentation (base.py:460-480): per class 1..14 keep the 6-connected component that overlaps the
 * registered sub-cortical mask most (first maximum in scipy.ndimage.label's raster order; the reference's argmax == 0
 * quirk included, see csrc/postproc.cu).  seg / mask / out: uint8 [X][Y][Z]; mask non-zero = inside; out != seg. */
SC_API int sc_post_process(sc_ctx* ctx, const uint8_t* seg_dev, const uint8_t* mask_dev, const int32_t dims[3],
                    uint8_t* out_dev, void* stream);

/* ---- scatter -------------------------------------------------------------------------
 * replaces: image[x,y,z] = y_pred and image_proba[x,y,z,c] = proba[:,c] (base.py:430-440) */
SC_API int sc_scatter(sc_ctx* ctx, const int32_t* xyz_dev, int64_t n, const int32_t* label_dev,
               const float* proba_dev, const int32_t dims[3], uint8_t* label_vol_dev,
               float* proba_vol_dev, void* stream);

/* ---- training ------------------------------------------------------------------------
 * replaces: one minibatch of nolearn's train_fn inside net.fit (nets.py:233-246):
 * training-mode forward (BN batch statistics, dropout p=.5), categorical cross-entropy,
 * backward.  Gradients land in the context's flat gradient buffer (parameter layout); the
 * slots of the BN running statistics receive this batch's mean / inv_std (applied by
 * sc_adam_step).  drop_masks_dev: NULL to draw masks from `seed`, or [n][3*540 + 540 + 540]
 * uint8 keep-masks per sample (axial|coronal|saggital l1drop in (c,h,w) order, f1_drop, f2_drop).
 * loss_dev: 1 float (sum of -log p over the batch divided by `n_global`). */
SC_API int sc_train_forward_backward(sc_ctx* ctx, const float* in1_dev, const float* in2_dev,
                              const float* in3_dev, const float* in4_dev, const uint8_t* y_dev,
                              int64_t n, int64_t n_global, uint64_t seed,
                              const uint8_t* drop_masks_dev, float* loss_dev, void* stream);
/* device pointers to the flat gradient / parameter buffers (SC_PARAM_FLOATS floats) so the
 * caller can all-reduce gradients in place (NCCL via torch.distributed). */
SC_API int sc_grad_buffer(sc_ctx* ctx, float** grads_dev);
SC_API int sc_param_buffer(sc_ctx* ctx, float** params_dev);
/* replaces: lasagne.updates.adam (nets.py:236-237): a_t = lr*sqrt(1-b2^t)/(1-b1^t);
 * p -= a_t * m / (sqrt(v) + eps) over the trainable entries, gradient multiplied by grad_scale first.
 * The slots of the BN running statistics in the gradient buffer carry the batch mean / inv_std
 * (summed over ranks by the all-reduce): s <- 0.9 s + 0.1 * slot * stat_scale (stat_scale = 1/world).
 * Marks the inference layouts stale; they are re-derived lazily by the next inference call. */
SC_API int sc_adam_step(sc_ctx* ctx, float lr, float beta1, float beta2, float eps, float grad_scale, float stat_scale, void* stream);
SC_API int sc_reset_optimizer(sc_ctx* ctx);
/* Data-parallel step as ONE kernel over NVLink peer memory: sum-all-reduce of the gradient buffers fused with sc_adam_step
 * (each rank reduces and updates its 1/world slice out of the peers' buffers and stores the new parameters into every peer;
 * csrc/fused_adam.cu).  One process per GPU: sc_fused_export writes 3 CUDA IPC handles (192 bytes: gradients, parameters,
 * flag page); the caller gathers the handles of all ranks (torch.distributed all_gather) and passes the world x 192 bytes,
 * in rank order, to sc_fused_attach on every rank (world <= 8).  sc_allreduce_adam_step must then be called by every rank
 * once per step, after sc_train_forward_backward on the same stream; it replaces all_reduce + sc_adam_step (the statistics
 * slots are averaged over the ranks).  All ranks end every step with bit-identical parameters. */
#define SC_IPC_BYTES 192
SC_API int sc_fused_export(sc_ctx* ctx, unsigned char* handles_out);
SC_API int sc_fused_attach(sc_ctx* ctx, int rank, int world, const unsigned char* all_handles);
SC_API int sc_allreduce_adam_step(sc_ctx* ctx, float lr, float beta1, float beta2, float eps, void* stream);
/* Synchronised BatchNorm for data-parallel training: the reference normalises over the whole batch of its
 * single device; with the hook set, the per-channel BatchNorm sums of the forward pass ({sum x, sum x^2}) and of the backward
 * pass ({sum dy, sum dy*xhat}) plus the element count are handed to `fn` as one small float64 device buffer right after they
 * are reduced locally; `fn` must sum it over all ranks IN PLACE, ordered on `stream` (an all-reduce of torch.distributed /
 * NCCL).  N ranks then reproduce the single-device step on the same global batch.  fn == NULL (default): per-GPU statistics.
 * With a hook the step is launched kernel by kernel (no CUDA graph). */
typedef int (*sc_allreduce_fn)(void* user, void* buf_dev /* double[count] */, int64_t count, void* stream);
SC_API int sc_set_allreduce_hook(sc_ctx* ctx, sc_allreduce_fn fn, void* user);
/* evaluation pass of nolearn's eval_fn: mean CE loss and accuracy numerators over a batch
 * in deterministic mode; out2_dev = {sum of -log p[y], number of correct argmax}. */
SC_API int sc_eval_batch(sc_ctx* ctx, const float* in1_dev, const float* in2_dev, const float* in3_dev,
                  const float* in4_dev, const uint8_t* y_dev, int64_t n, float* out2_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SUBCORT_B200_H */
