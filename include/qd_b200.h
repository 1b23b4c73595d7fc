/*
 * qd_b200.h -- C ABI of libqd_b200.so: the B200 (sm_100a) implementation of the
 * fake-quantization hot path of antspy/quantized_distillation.
 *
 * The reference has no FFI layer: its boundary for this path is the Python
 * package `quantization` (quantization/__init__.py:3-8).  This header is the
 * boundary *behind* that package in the new implementation: every entry point
 * below replaces the body of one reference function (cited as
 * path:line in the reference checkout) and is what a ctypes / cffi / pybind
 * stub on the reference side would bind.
 *
 * Conventions
 *   - all tensor pointers are DEVICE pointers to contiguous float32 storage in
 *     C order unless the name ends in `_host`; no torch types cross the ABI.
 *   - `n` is the number of elements, `bucket` the bucket size, 0 meaning the
 *     reference's `bucket_size=None` (one bucket spanning the tensor).
 *   - bucket geometry follows create_bucket_tensor
 *     (quantization/help_functions.py:67-94): rows = ceil(n/bucket) unless
 *     n < bucket (one short row); the tail row behaves as if padded with copies
 *     of the last element, but the padding is never materialised except in the
 *     `xhat` output of qd_scale_down, whose length is padded_len.
 *   - per-row outputs (alpha, beta: float32[rows]; argmin, argmax: int64[rows],
 *     index inside the row, first occurrence) may be NULL when not wanted.
 *   - device: kernels launch on the calling thread's CURRENT CUDA device (one
 *     process per GPU is the deployment model); the device pointers and
 *     `stream` of a call must belong to it.  Only the *_host entry points take
 *     a device ordinal and switch (and restore) the device themselves.
 *   - `stream` is a cudaStream_t; every call only enqueues work on it (no host
 *     synchronisation) except the *_host entry points, which return after the
 *     result is in host memory.
 *   - `workspace` is device scratch of at least qd_workspace_bytes(n, bucket)
 *     bytes, owned by the caller, private to the stream for the call.
 *   - return value: 0 (QD_OK) or a qd_status; qd_last_error() gives the
 *     message for the calling thread.  Nothing throws, nothing aborts.
 *   - arithmetic: float32, one IEEE round-to-nearest-even per reference torch
 *     op, no FMA contraction, true division, rintf -- results are bit-identical
 *     to the reference's CPU path for q / idx / alpha / beta / argmin / argmax.
 */
#ifndef QD_B200_H
#define QD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* qd_stream_t; /* cudaStream_t */

typedef enum {
    QD_OK = 0,
    QD_ERR_INVALID_ARG = 1,  /* reference raises ValueError (quant_functions.py:22-33,138-139,230-236) */
    QD_ERR_UNSUPPORTED = 2,  /* reference raises NotImplementedError (quant_functions.py:329-334) */
    QD_ERR_CUDA = 3,         /* a CUDA runtime call failed; message holds cudaGetErrorString */
    QD_ERR_WORKSPACE = 4     /* workspace missing or too small */
} qd_status;

/* gradient fix-up styles of the training loop (cnn_models/conv_forward_model.py:249-266) */
typedef enum {
    QD_BWD_STE = 0,       /* 'none': straight-through, gout = g */
    QD_BWD_TRUNCATED = 1, /* 'truncated': gout = |x| > 1 ? 0 : g          (:263-264) */
    QD_BWD_MINMAX = 2     /* 'complicated': uniformQuantization_variable.backward (quant_functions.py:319-406) */
} qd_bwd_mode;

/* index rule of the non-uniform op */
typedef enum {
    QD_RULE_NEAREST = 0,  /* nonUniformQuantization direct path (quant_functions.py:267-273) */
    QD_RULE_MIDPOINT = 1  /* SearchSorted.query, pre-processed path  (quant_functions.py:531-573) */
} qd_rule;

/* ---- library ----------------------------------------------------------- */
int qd_version(void);
const char* qd_last_error(void);
/* SM count and compute capability of the current device. */
int qd_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- a1: create_bucket_tensor geometry (help_functions.py:67-94) -------- */
int qd_bucket_geometry(int64_t n, int64_t bucket, int64_t* rows, int64_t* row_len, int64_t* padded_len);
size_t qd_workspace_bytes(int64_t n, int64_t bucket);

/* ---- a2: ScalingFunction.scale_down, linear (quant_functions.py:56-107) --
 * xhat[padded_len] = (x' - beta)/alpha with x' = clamp(x - *mean, +-max_element);
 * mean NULL = no mean subtraction, max_element <= 0 = no clamp.  xhat may be
 * NULL to compute only the per-row state. */
int qd_scale_down(const float* x, float* xhat, float* alpha, float* beta, int64_t* argmin, int64_t* argmax,
                  int64_t n, int64_t bucket, const float* mean, float max_element,
                  void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* ---- a3: ScalingFunction.inv_scale_down (quant_functions.py:131-152) -----
 * out[n] = (y*alpha + beta) + *mean over the first n of y[padded_len]. */
int qd_inv_scale_down(const float* y, float* out, const float* alpha, const float* beta, const float* mean,
                      int64_t n, int64_t bucket, qd_stream_t stream);

/* ---- a10: absmax / absnorm scaling -- EXTENSION, parity unpinned ---------
 * quant_functions.py:109-127, 144-146 cannot execute in the reference (`tensor.max(p=2)`, a bound method stored
 * as the scale), so no output of the reference exists to compare with.  These entry points implement what the
 * lines intend once the two slips are repaired: sign = sign(x), v = |x|, norm_b = max v (ABSMAX) or
 * sqrt(sum v^2) over the padded bucket (ABSNORM), norm_b < 1e-10 -> 1, x_hat = v / norm_b; inverse
 * y * norm_b * sign (+ mean).  xhat / sign use the padded layout of qd_scale_down; norm: float32[rows]. */
typedef enum { QD_SCALE_ABSMAX = 1, QD_SCALE_ABSNORM = 2 } qd_abs_scaling;
int qd_scale_down_abs(const float* x, float* xhat, float* sign, float* norm, int64_t n, int64_t bucket, int kind,
                      const float* mean, float max_element, qd_stream_t stream);
int qd_inv_scale_down_abs(const float* y, const float* sign, const float* norm, const float* mean, float* out,
                          int64_t n, int64_t bucket, qd_stream_t stream);
/* uniformQuantization(type_of_scaling='absmax'|'absnorm'): q = ((rint(x_hat*S)/S) * norm_b) * sign (+ mean) */
int qd_uniform_fwd_abs(const float* x, float* q, uint8_t* idx_u8, float* norm, int64_t n, int64_t bucket, int levels,
                       int kind, const float* mean, float max_element, qd_stream_t stream);

/* ---- a4: uniformQuantization (quant_functions.py:155-194) ----------------
 * q[n] (may alias x: modify_in_place); idx_u8[n] optional integer levels
 * (levels <= 256); levels = s >= 2.  stochastic != 0 selects stochastic
 * rounding (:174-187) with a Philox stream (seed, offset) -- distributional
 * parity only, the reference draws from torch.rand on the host. */
int qd_uniform_fwd(const float* x, float* q, uint8_t* idx_u8, float* alpha, float* beta, int64_t* argmin,
                   int64_t* argmax, int64_t n, int64_t bucket, int levels, const float* mean, float max_element,
                   int stochastic, uint64_t seed, uint64_t offset,
                   void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* ---- a5: backward of the uniform op --------------------------------------
 * gout[n] (may alias g).  QD_BWD_MINMAX requires bucket != 0 like the
 * reference (quant_functions.py:332-334) and bucket <= QD_MAX_STAGED_BUCKET. */
int qd_uniform_bwd(const float* x, const float* g, float* gout, int64_t n, int64_t bucket, int levels, int mode,
                   void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* forward + backward in one pass over (x, g): 16 bytes per element. */
int qd_uniform_fwd_bwd(const float* x, const float* g, float* q, float* gout, int64_t n, int64_t bucket,
                       int levels, int mode, void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* ---- a6/a7: nonUniformQuantization (quant_functions.py:196-290, 509-573) -
 * points[K] device, sorted ascending; idx_u8 (K <= 256) and/or idx_i64 optional. */
int qd_nonuniform_fwd(const float* x, const float* points, int num_points, int rule, float* q, uint8_t* idx_u8,
                      int64_t* idx_i64, float* alpha, float* beta, int64_t n, int64_t bucket, const float* mean,
                      float max_element, void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* ---- a8: nonUniformQuantization_variable.backward (quant_functions.py:471-506)
 * grad_points[K] = sum_{i: idx_i = k} fl32(g_i * alpha_row(i)); exactly one of
 * idx_u8 / idx_i64 non-NULL.  Deterministic (fixed reduction tree, float64
 * accumulation), so data-parallel replicas stay bit-identical. */
int qd_nonuniform_bwd(const float* g, const uint8_t* idx_u8, const int64_t* idx_i64, const float* alpha,
                      int num_points, float* grad_points, int64_t n, int64_t bucket,
                      void* workspace, size_t workspace_bytes, qd_stream_t stream);

/* index search alone, on values that are already scaled to [0,1] (the
 * reference's pre-processed path hands SearchSorted an already scaled tensor,
 * quant_functions.py:432-447, 275): idx = rule(xhat_i), unit_out_i = points[idx].
 * Any of idx_u8 / idx_i64 / unit_out may be NULL. */
int qd_centroid_index(const float* xhat, const float* points, int num_points, int rule, uint8_t* idx_u8,
                      int64_t* idx_i64, float* unit_out, int64_t n, qd_stream_t stream);

/* ---- next row f2: index histogram for the Huffman statistics
 * (help_functions.py:223-225): counts[b] += #{ i : idx_i = b }, b < num_bins <= 256;
 * counts is a device int64[num_bins] the caller zeroes (it accumulates across tensors). */
int qd_index_histogram(const uint8_t* idx_u8, int64_t n, int num_bins, int64_t* counts, qd_stream_t stream);

/* ---- next row f2: packed integer codec (the compressed-model deliverable the reference only
 * accounts for: helpers/functions.py:216-262).  bits in {1, 2, 4, 8}; code i of element e sits in
 * byte e*bits/8 at bit offset (e*bits)%8 (little endian).  packed has ceil(n*bits/8) bytes. */
int qd_pack_indices(const uint8_t* idx_u8, uint8_t* packed, int64_t n, int bits, qd_stream_t stream);
/* q[n] rebuilt from packed uniform levels and the per-row (alpha, beta): bit-identical to the
 * output of qd_uniform_fwd that produced the levels. */
int qd_unpack_dequant_uniform(const uint8_t* packed, int bits, const float* alpha, const float* beta, float* q,
                              int64_t n, int64_t bucket, int levels, qd_stream_t stream);
/* same for centroid codes: q = points[code]*alpha + beta */
int qd_unpack_dequant_nonuniform(const uint8_t* packed, int bits, const float* points, int num_points,
                                 const float* alpha, const float* beta, float* q, int64_t n, int64_t bucket,
                                 qd_stream_t stream);

/* ---- next row f1: one launch over every parameter tensor of a model ------
 * (replaces the per-tensor loop of cnn_models/conv_forward_model.py:236-247).
 * A plan owns a device-side table of (src, dst, n, levels); pointers must stay
 * valid while the plan lives.  src[i] == dst[i] quantizes in place. */
typedef struct qd_plan qd_plan;
int qd_plan_create(qd_plan** plan, int count, const float* const* src, float* const* dst, const int64_t* n,
                   const int32_t* levels, int64_t bucket);
int qd_plan_destroy(qd_plan* plan);
/* Optional shadow buffers (one per tensor, n[i] floats): when set, qd_plan_uniform_fwd with
 * save != 0 also writes the untouched full-precision row there -- the reference's
 * `model_state_dict = model.state_dict()` (conv_forward_model.py:286) for free in the same pass. */
int qd_plan_set_shadow(qd_plan* plan, float* const* shadow);
int qd_plan_uniform_fwd(const qd_plan* plan, qd_stream_t stream);
int qd_plan_uniform_fwd_save(const qd_plan* plan, qd_stream_t stream);
/* gout_i = bwd(src_i, grad_i) for every tensor, in place in grad. */
int qd_plan_uniform_bwd(const qd_plan* plan, float* const* grad, int mode, qd_stream_t stream);

/* Fused end-of-step: for every tensor, in ONE pass (24 bytes per element):
 *   grad    <- fix-up of `mode` evaluated at the full-precision shadow copy   (conv_forward_model.py:249-266, :315)
 *   shadow, momentum <- torch.optim.SGD update (momentum, Nesterov, weight decay, dampening 0)   (:317)
 *   dst     <- uniformQuantization(shadow)   -- the next step's quantized weights            (:286-287)
 * QD_BWD_TRUNCATED also clamps the updated weights to [-1, 1] (the next step's clamp, :240-241).
 * momentum[i]: n[i] floats, zero before the first step.  Rows of at most 512 elements, else QD_ERR_UNSUPPORTED.
 * Arithmetic = torch's CUDA SGD kernels (a + alpha*b contracted to one FMA): see qd_plan.cuh. */
int qd_plan_set_momentum(qd_plan* plan, float* const* momentum);
int qd_plan_sgd_step(const qd_plan* plan, float* const* grad, int mode, double lr, double momentum,
                     double weight_decay, int nesterov, qd_stream_t stream);

/* ---- the same for the differentiable-quantization loop (cnn_models/conv_forward_model.py:501-551):
 * one launch re-quantizes every tensor with its own current list of points (midpoint rule of the
 * pre-processed path, quant_functions.py:531-573), two small launches produce every tensor's centroid
 * gradient (:471-506), deterministically.  Per tensor i: src[i] the fixed full-precision tensor,
 * dst[i] the live parameter (receives q), idx[i] uint8[n], alpha[i] / beta[i] float[rows] (written by
 * the forward, read by the backward), points[i] float[num_points[i]] ascending, device memory, re-read
 * at every launch, grad_points[i] float[num_points[i]].  1 <= num_points <= 32 and rows of at most
 * 1024 elements, otherwise QD_ERR_UNSUPPORTED (use the per-tensor entry points). */
typedef struct qd_nu_plan qd_nu_plan;
int qd_plan_nonuniform_create(qd_nu_plan** plan, int count, const float* const* src, float* const* dst,
                              uint8_t* const* idx, float* const* alpha, float* const* beta,
                              const float* const* points, float* const* grad_points, const int64_t* n,
                              const int32_t* num_points, int64_t bucket);
int qd_plan_nonuniform_destroy(qd_nu_plan* plan);
int qd_plan_nonuniform_fwd(const qd_nu_plan* plan, qd_stream_t stream);
/* grad[i]: dLoss/d(quantized tensor i), float[n[i]] */
int qd_plan_nonuniform_bwd(const qd_nu_plan* plan, const float* const* grad, qd_stream_t stream);

/* ---- next row f3: the reductions of the differentiable-quantization setup -----------------
 * Exact order statistics without a sort: out[r] = the ranks[r]-th smallest element of v (0-based; ranks
 * is a DEVICE array of num_ranks <= 512 entries, clamped to [0, n-1]).  This is what
 * np.percentile(x_hat, linspace(0, 100, K)) reads (help_functions.py:140-154).  Two reads of v
 * (value histogram, then compaction of the selected bins) plus a radix select on the compacted keys.
 * workspace: qd_order_statistics_workspace_bytes(n) bytes. */
size_t qd_order_statistics_workspace_bytes(int64_t n);
int qd_order_statistics(const float* v, int64_t n, const int64_t* ranks, int num_ranks, float* out,
                        void* workspace, size_t workspace_bytes, qd_stream_t stream);
/* out[i] = ||tensors[i]||_2 for count tensors (host array of device pointers) in two launches, float64
 * partial sums in a fixed order: the gradient norms of assign_bits_automatically
 * (cnn_models/conv_forward_model.py:424-448).  Setup-time call: synchronises the stream once. */
int qd_multi_l2norm(const float* const* tensors, const int64_t* n, int count, float* out, qd_stream_t stream);

/* ---- host-buffer entry points (what a CPU-tensor caller gets) ------------
 * Inputs and outputs in HOST memory (pinned for full PCIe rate); the call
 * pipelines H2D, the fused kernel and D2H in row-aligned chunks on internal
 * streams of `device` -- or, for pinned tensors of at most 8 Mi elements, runs
 * one launch straight on the (device-addressable) host pointers -- and returns
 * when the outputs are complete.  Pageable memory is accepted (slower copies). */
int qd_uniform_fwd_host(const float* x_host, float* q_host, int64_t n, int64_t bucket, int levels, int device);
int qd_uniform_fwd_bwd_host(const float* x_host, const float* g_host, float* q_host, float* gout_host,
                            int64_t n, int64_t bucket, int levels, int mode, int device);

/* ---- benchmark hook: override a path-selection threshold (tools/block_bench.py measures the
 * variants against each other with it); value -1 restores the built-in choice.
 *   key 0: longest row (floats) taken by the warp-per-row two-pass variant
 *   key 1: longest row (floats) that keeps two rows in flight per CTA in the staged path
 *   key 2: threads per CTA of the staged path (64 / 128 / 256 / 512 / 1024)
 *   key 3: warp-path min/max backward sums r_b per element in float64 (1) or in float32 groups of four (0);
 *          built-in choice: per element when q is written in the same pass, grouped for the backward alone
 *   key 4: longest row (floats) taken by the warp path (<= 1024)
 *   key 5 / 6 / 7: host entry points: pipeline slots (1..8) / chunk elements / staging path (0 = chunked copies,
 *          1 = one launch on pinned host pointers) */
int qd_debug_set_tuning(int key, int64_t value);

/* ---- self tests used by tests/ (device side arithmetic checks) ---------- */
int qd_selftest_division(int64_t pairs, uint64_t seed, int64_t* mismatches, qd_stream_t stream);

#define QD_MAX_STAGED_BUCKET 49152 /* floats; largest bucket staged in shared memory */

#ifdef __cplusplus
}
#endif
#endif /* QD_B200_H */
