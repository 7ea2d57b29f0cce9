/*
 * msda_b200.h -- C ABI of the B200-native multi-scale deformable attention operator.
 *
 * This is the drop-in boundary for the one hot path this repository rebuilds: the
 * MSDeformAttn operator of On-JungWoan/UVHand (reference tree: models/ops).  Plain pointers
 * and sizes only; no torch types.  Built as uvhand_b200/lib/libmsda_b200.so for sm_100a.
 *
 * Every entry point names the reference interface it replaces (paths relative to
 * /root/reference).  Tensor layouts are the reference's:
 *
 *   value            [N, S, M, D]        S = sum_l H_l*W_l value tokens, M heads, D channels/head
 *   spatial_shapes   [L, 2] int64        (H_l, W_l)                       -- DEVICE memory
 *   level_start      [L]    int64        exclusive cumsum of H_l*W_l      -- DEVICE memory
 *   sampling_loc     [N, Lq, M, L, P, 2] (x, y) normalised to [0,1]; may lie outside
 *   attn_weight      [N, Lq, M, L, P]
 *   output           [N, Lq, M, D]       (the reference returns this viewed as [N, Lq, M*D])
 *   grad_*           same shapes as the tensors they belong to
 *
 * Return value: 0 on success, a positive cudaError_t if the CUDA runtime reported one
 * (the reference only printf()s launch errors, ms_deform_im2col_cuda.cuh:948-952; this ABI
 * returns them), or a negative MSDA_ERR_* code for a rejected argument.  All launches are
 * asynchronous on `stream` (a cudaStream_t passed as void*; NULL = the legacy default stream).
 * There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef MSDA_B200_H_
#define MSDA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSDA_B200_ABI_VERSION 1

#define MSDA_OK 0
#define MSDA_ERR_NULL_POINTER (-1)  /* a required pointer is NULL                                  */
#define MSDA_ERR_BAD_DIMS (-2)      /* a dimension is negative, or a product overflows the kernel's index range */
#define MSDA_ERR_IM2COL_STEP (-3)   /* batch % min(batch, im2col_step) != 0  (ms_deform_attn_cuda.cu:50-52) */
#define MSDA_ERR_UNSUPPORTED (-4)   /* combination not built (e.g. bf16 value with D != 32)        */
#define MSDA_ERR_MISALIGNED (-5)    /* pointer not aligned to its element size                     */
#define MSDA_ERR_NO_DEVICE (-6)     /* no usable CUDA device                                       */
#define MSDA_ERR_BAD_OPTION (-7)    /* msda_b200_set_option: unknown key or value                  */

int msda_b200_abi_version(void);
/* Static string for any return code of this library (CUDA codes via cudaGetErrorString). */
const char *msda_b200_strerror(int rc);

/* The reference's chunking precondition, ms_deform_attn_cuda.cu:50-52 and :117-119:
 * im2col_step_ = min(batch, im2col_step); batch % im2col_step_ must be 0.  The kernels here
 * process the whole batch in one launch (results do not depend on the chunking), but callers
 * keep the argument and its error.  Returns MSDA_OK or MSDA_ERR_IM2COL_STEP. */
int msda_b200_check_im2col_step(int batch, int im2col_step);

/* ---- forward ---------------------------------------------------------------------------
 * Replaces ms_deformable_im2col_cuda<scalar_t>  (src/cuda/ms_deform_im2col_cuda.cuh:923-954)
 * and the kernel it launches, ms_deformable_im2col_gpu_kernel (cuh:237-299).  `output` is
 * fully written (no pre-zeroing needed; the reference zero-fills it first, .cu:54). */
int msda_fwd_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start,
                 const float *sampling_loc, const float *attn_weight,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point, float *output, void *stream);
int msda_fwd_f64(const double *value, const int64_t *spatial_shapes, const int64_t *level_start,
                 const double *sampling_loc, const double *attn_weight,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point, double *output, void *stream);
/* bf16 value (storage: raw bf16 bits), fp32 locations / weights / output.  channels must be 32.
 * No reference analogue (the reference dispatches float and double only, .cu:64). */
int msda_fwd_bf16(const uint16_t *value, const int64_t *spatial_shapes, const int64_t *level_start,
                  const float *sampling_loc, const float *attn_weight,
                  int batch, int spatial_size, int num_heads, int channels, int num_levels,
                  int num_query, int num_point, float *output, void *stream);

/* ---- backward --------------------------------------------------------------------------
 * Replaces ms_deformable_col2im_cuda<scalar_t> (cuh:956-1327) and its six kernels
 * (cuh:301-920).  grad_value is ACCUMULATED into and must be zero on entry (the reference
 * host code zero-fills it, .cu:121); grad_sampling_loc and grad_attn_weight are fully written (as zeros when
 * spatial_size is 0: no sample can land anywhere). */
int msda_bwd_f32(const float *grad_output, const float *value, const int64_t *spatial_shapes,
                 const int64_t *level_start, const float *sampling_loc, const float *attn_weight,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point,
                 float *grad_value, float *grad_sampling_loc, float *grad_attn_weight, void *stream);
int msda_bwd_f64(const double *grad_output, const double *value, const int64_t *spatial_shapes,
                 const int64_t *level_start, const double *sampling_loc, const double *attn_weight,
                 int batch, int spatial_size, int num_heads, int channels, int num_levels,
                 int num_query, int num_point,
                 double *grad_value, double *grad_sampling_loc, double *grad_attn_weight, void *stream);
/* bf16 value and bf16 grad_value, everything else fp32.  Two accumulation modes:
 *   grad_value_scratch != NULL  (accurate, what the torch layer uses by default): an fp32 buffer of
 *       batch*spatial_size*num_heads*channels floats, ZERO on entry; reductions go into it with
 *       red.global.add.v4.f32 and it is rounded once into grad_value (which is then fully written);
 *   grad_value_scratch == NULL  (direct): red.global.add.noftz.bf16x2 straight into grad_value, which
 *       must be zero on entry.  Half the reduction bytes, but every add rounds the running sum to
 *       bf16, so heavily hit rows drift (about 1.6e-2 relative L2 at the DETR-encoder shape). */
int msda_bwd_bf16(const float *grad_output, const uint16_t *value, const int64_t *spatial_shapes,
                  const int64_t *level_start, const float *sampling_loc, const float *attn_weight,
                  int batch, int spatial_size, int num_heads, int channels, int num_levels,
                  int num_query, int num_point,
                  uint16_t *grad_value, float *grad_sampling_loc, float *grad_attn_weight,
                  float *grad_value_scratch, void *stream);

/* ---- fused prologue ----------------------------------------------------------------------
 * The operator with the two elementwise steps that precede it in MSDeformAttn.forward folded in
 * (models/ops/modules/ms_deform_attn.py:100-108):
 *     attn_weight  = softmax(attn_logits over L*P)
 *     sampling_loc = reference_points[:, :, None, :, None, :] + sampling_offsets / (W_l, H_l)
 * so that locations and weights are never written to / re-read from HBM.
 *   reference_points [N, Lq, L, 2]        (the 42-keypoint variant passes the keypoint centre)
 *   sampling_offsets [N, Lq, M, L, P, 2]  raw output of the sampling_offsets Linear
 *   attn_logits      [N, Lq, M, L*P]      raw output of the attention_weights Linear
 * Location arithmetic is torch's operation for operation, so corner indices equal the unfused
 * path's bit for bit.  channels must be 32, num_levels <= 8, num_levels*num_point <= 32
 * (MSDA_ERR_UNSUPPORTED otherwise: callers then use the unfused entry points).
 * Backward: grad_value accumulated (zero on entry); grad_sampling_offsets / grad_attn_logits fully
 * written; grad_reference_points accumulated (zero on entry) or NULL when not wanted. */
int msda_fused_fwd_f32(const float *value, const int64_t *spatial_shapes, const int64_t *level_start,
                       const float *reference_points, const float *sampling_offsets,
                       const float *attn_logits,
                       int batch, int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point, float *output, void *stream);
int msda_fused_bwd_f32(const float *grad_output, const float *value, const int64_t *spatial_shapes,
                       const int64_t *level_start, const float *reference_points,
                       const float *sampling_offsets, const float *attn_logits,
                       int batch, int spatial_size, int num_heads, int channels, int num_levels,
                       int num_query, int num_point,
                       float *grad_value, float *grad_sampling_offsets, float *grad_attn_logits,
                       float *grad_reference_points, void *stream);

/* ---- value producer side ------------------------------------
 * MSDeformAttn.forward builds the operator's `value` as value_proj(x).masked_fill(padding_mask[..., None], 0)
 * (models/ops/modules/ms_deform_attn.py:96-98).  For the bf16-value kernels these do the mask and the cast in ONE
 * pass over the projection's fp32 output, and the matching backward:
 *   msda_value_pack_bf16        value_bf16[r, :] = padding_mask[r] ? 0 : bf16(value_f32[r, :])
 *   msda_value_unpack_grad_bf16 grad_f32[r, :]   = padding_mask[r] ? 0 : float(grad_bf16[r, :])
 * rows = batch * spatial_size tokens; cols = num_heads * channels (a multiple of 8); padding_mask: one byte per row,
 * non-zero = padded token, or NULL for no mask.  No reference analogue (the reference has no bf16 mode). */
int msda_value_pack_bf16(const float *value_f32, const uint8_t *padding_mask, int64_t rows, int cols,
                         uint16_t *value_bf16, void *stream);
int msda_value_unpack_grad_bf16(const uint16_t *grad_bf16, const uint8_t *padding_mask, int64_t rows, int cols,
                                float *grad_f32, void *stream);

/* ---- parity probe ----------------------------------------------------------------------
 * Dumps, for every (n, q, m, l, p), what the kernels' shared sample-geometry routine decides:
 * floor of the pixel coordinate and the corner mask (bit0..3 = corner v1..v4 inside the level,
 * bit4 = sample accepted by the coarse test of cuh:288).  Rejected samples report 0/0/0.
 * This is how "indices bit-exact" is checked on the device; it runs the same __device__
 * function the forward and backward kernels inline. */
int msda_sample_index_f32(const int64_t *spatial_shapes, const float *sampling_loc,
                          int batch, int num_query, int num_heads, int num_levels, int num_point,
                          int32_t *h_low, int32_t *w_low, uint8_t *mask, void *stream);

/* ---- host-buffer entry points (what a non-torch caller binds) ---------------------------
 * All pointers are HOST memory (pinned or pageable).  The call copies inputs to the device,
 * runs forward (and backward), copies results back and returns when they are in place.
 * A plan owns the device buffers, pinned staging and streams for one problem size; create it
 * once and reuse it.  These wrap the same device entry points as above. */
typedef struct msda_host_plan msda_host_plan;

int msda_host_plan_create(msda_host_plan **plan, int device,
                          int batch, int spatial_size, int num_heads, int channels,
                          int num_levels, int num_query, int num_point);
void msda_host_plan_destroy(msda_host_plan *plan);
/* Bytes moved host->device and device->host by one msda_host_fwd_bwd_f32 call. */
int msda_host_plan_traffic(const msda_host_plan *plan, int64_t *h2d_bytes, int64_t *d2h_bytes);
/* Pin / unpin a caller buffer so the copies can run asynchronously at full PCIe rate. */
int msda_host_register(void *ptr, int64_t bytes);
int msda_host_unregister(void *ptr);

int msda_host_fwd_f32(msda_host_plan *plan, const float *value, const int64_t *spatial_shapes,
                      const int64_t *level_start, const float *sampling_loc,
                      const float *attn_weight, float *output);
int msda_host_fwd_bwd_f32(msda_host_plan *plan, const float *value, const int64_t *spatial_shapes,
                          const int64_t *level_start, const float *sampling_loc,
                          const float *attn_weight, const float *grad_output, float *output,
                          float *grad_value, float *grad_sampling_loc, float *grad_attn_weight);

/* bf16 storage for the two value-shaped tensors (the only ones the operator can take in bf16): `value` is uploaded
 * and `grad_value` downloaded as bf16 (2 x 22.8 MB less PCIe traffic per call at the DETR-encoder shape); grad_value is
 * accumulated in fp32 on the device and rounded once.  channels must be 32 (MSDA_ERR_UNSUPPORTED otherwise).  A plan is
 * either fp32 or bf16: the other family of calls returns MSDA_ERR_UNSUPPORTED. */
int msda_host_plan_create_bf16(msda_host_plan **plan, int device,
                               int batch, int spatial_size, int num_heads, int channels,
                               int num_levels, int num_query, int num_point);
int msda_host_fwd_bwd_bf16(msda_host_plan *plan, const uint16_t *value, const int64_t *spatial_shapes,
                           const int64_t *level_start, const float *sampling_loc,
                           const float *attn_weight, const float *grad_output, float *output,
                           uint16_t *grad_value, float *grad_sampling_loc, float *grad_attn_weight);

/* ---- several host-buffer steps in flight ---------------------------------------------------
 * A pipe is `depth` (1..8) plans that share one set of upload / compute / download streams.
 * msda_host_pipe_submit_f32 enqueues one forward + backward step on slot `slot` (0 <= slot < depth) and returns without
 * waiting; msda_host_pipe_wait blocks until that slot's step has landed in the caller's output buffers (slot < 0: every
 * slot).  A submit on a slot whose previous step is still in flight waits for it first.  Because the streams are
 * shared, the uploads of step i + 1 overlap the kernels and downloads of step i: a stream of steps (a training loop
 * that prefetches its next batch) moves at the PCIe rate of large two-way copies instead of paying every step's
 * pipeline fill and drain.  Buffers must stay valid and untouched until the slot has been waited for; page-locked
 * buffers (msda_host_register) are required for the copies to be asynchronous at all.  One host thread drives a pipe
 * (like a plan, it is not thread-safe); use one pipe per thread or device.  Results are those of
 * msda_host_fwd_bwd_f32 up to summation order.  (MSDA_B200_PIPE_CHUNKS: query chunks per sample inside a step; default 0 =
 * the step is one packet: four uploads, one forward + backward over the batch, four downloads.) */
typedef struct msda_host_pipe msda_host_pipe;
int msda_host_pipe_create(msda_host_pipe **pipe, int device, int depth,
                          int batch, int spatial_size, int num_heads, int channels,
                          int num_levels, int num_query, int num_point);
void msda_host_pipe_destroy(msda_host_pipe *pipe);
int msda_host_pipe_submit_f32(msda_host_pipe *pipe, int slot, const float *value, const int64_t *spatial_shapes,
                              const int64_t *level_start, const float *sampling_loc,
                              const float *attn_weight, const float *grad_output, float *output,
                              float *grad_value, float *grad_sampling_loc, float *grad_attn_weight);
/* The same with value / grad_value travelling as bf16 (see msda_host_plan_create_bf16); a pipe is either fp32 or bf16. */
int msda_host_pipe_create_bf16(msda_host_pipe **pipe, int device, int depth,
                               int batch, int spatial_size, int num_heads, int channels,
                               int num_levels, int num_query, int num_point);
int msda_host_pipe_submit_bf16(msda_host_pipe *pipe, int slot, const uint16_t *value, const int64_t *spatial_shapes,
                               const int64_t *level_start, const float *sampling_loc,
                               const float *attn_weight, const float *grad_output, float *output,
                               uint16_t *grad_value, float *grad_sampling_loc, float *grad_attn_weight);
int msda_host_pipe_wait(msda_host_pipe *pipe, int slot);
int msda_host_pipe_traffic(const msda_host_pipe *pipe, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- tuning / introspection --------------------------------------------------------------
 * Keys: "fwd_path", "bwd_path": 0 = automatic, 1 = force the any-shape kernels (used by bench.py
 * and the tests to exercise every kernel on the same inputs); "bf16_direct_red": 0/1, read by the
 * torch host layer to pick msda_bwd_bf16's accumulation mode (0 = fp32 scratch, the default);
 * "split_mode": which D=32 kernels serve a call: 0 = by problem size (small calls such as 300/900
 * decoder queries use the one-item-per-warp kernels, large ones four items per warp), 1 = never
 * split, 2 = always split; "bwd_tiles": the backward that combines grad_value rows on chip before they leave
 * the SM (num_levels = num_point = 4): 0 = chosen by shape (encoder self-attention num_query == spatial_size, or a
 * pyramid of at most 1500 tokens), 1 = never, 2 = whenever the shape is eligible; "fwd_wide": the forward with 4 lanes
 * per (query, head) and 8 channels per lane (num_levels = num_point = 4): 0 = for bf16 values, 1 = never, 2 = also for
 * fp32 values; "fwd_slab": the fp32 forward that stages one head of one sample in shared memory and gathers from there
 * (num_levels = num_point = 4, spatial_size <= 1600): 0 = when the call has at least two 64-query passes per SM,
 * 1 = never, 2 = whenever the shape is eligible.  None of these changes results
 * beyond summation order.  Options are process-global (every thread and stream sees them); they exist for tests and
 * benchmarks, production callers leave them at 0.
 * Returns MSDA_OK or MSDA_ERR_BAD_OPTION. */
int msda_b200_set_option(const char *key, int value);
int msda_b200_get_option(const char *key, int *value);
/* Number of kernels this library has launched since load (all streams); bench.py's
 * gpu_launches claim is read from here. */
int64_t msda_b200_launch_count(void);
/* Name of the kernel family the last forward / backward call dispatched to. */
const char *msda_b200_last_path(int backward);

#ifdef __cplusplus
}
#endif
#endif /* MSDA_B200_H_ */
