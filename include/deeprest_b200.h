/*
 * deeprest_b200.h — C ABI of libdeeprest_b200.so, the B200-native replacement for the
 * DeepRest resource-estimator hot path (reference: resource-estimation/qrnn.py:6-67,
 * driven from resource-estimation/estimate.py:60-107).
 *
 * The reference has no FFI: the path is a torch nn.Module called in-process
 * (SURVEY §8b).  Each entry point below names the reference interface it replaces.
 * The ABI is cgo/ctypes/JNI-safe by construction: C types only, no callbacks, no
 * exceptions across the boundary, no caller pointer retained after a call returns.
 *
 * Conventions
 *   - every function returns 0 (DR_OK) or a negative DR_E* code; the message for the last
 *     failure on a handle is dr_last_error(handle) (dr_last_error(NULL) for dr_create).
 *   - host entry points (no _dev suffix) take HOST pointers, copy in/out, and return only
 *     when the result is in the caller's buffer.
 *   - *_dev entry points take DEVICE pointers and are asynchronous on the handle's stream.
 *   - a handle is not thread-safe; distinct handles are independent. One handle per GPU.
 *   - there is NO CPU fallback: without a CUDA device dr_create fails with DR_ECUDA.
 *
 * Shapes (reference notation): B windows, T seq_len, F input features, M experts
 * (= component_resource series, qrnn.py:21), H = 128, Q = 3 quantiles.
 * Weight blob = the reference state_dict() order without mask_init (see
 * deeprest_b200/layout.py).
 */
#ifndef DEEPREST_B200_H
#define DEEPREST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR_OK            0
#define DR_EINVAL       -1   /* bad argument / shape                                        */
#define DR_ECUDA        -2   /* CUDA runtime failure or no device                           */
#define DR_ENOMEM       -3   /* device or host allocation failed                            */
#define DR_ESTATE       -4   /* call order (e.g. forward before load_weights)               */
#define DR_EUNSUPPORTED -5   /* hyper-parameter the reference allows but this build does not */

#define DR_H 128
#define DR_Q 3

/* engine selection for the bi-GRU recurrence */
#define DR_ENGINE_AUTO  0    /* tcgen05 where it applies (F <= 64 for the forward), else FFMA */
#define DR_ENGINE_FFMA  1    /* fp32 CUDA-core kernel (exact fp32 accumulate)               */
#define DR_ENGINE_TC    2    /* tcgen05 tensor-core kernels, split-fp16 (3-pass) operands   */

/* arithmetic of the TRAINING step (dr_config.dtype); inference always runs the fp32-parity split-fp16 engine */
#define DR_DTYPE_F32    0    /* split-fp16 (3 tensor passes) or FFMA: every gradient inside the fp32 parity bounds          */
#define DR_DTYPE_BF16   1    /* single-pass bf16 operands, fp32 accumulate, bf16-stored activations (BASELINE configs[2],[4]) */

typedef struct dr_model dr_model;

/* Mirrors QuantileRNN.__init__(input_size, num_metrics, hidden_layer_size=128, num_layers=1,
 * bidirectional=True, quantiles=(.05,.5,.95), dropout=.5)  — qrnn.py:7-8.
 * H, num_layers and bidirectional are fixed to the reference defaults. */
typedef struct dr_config {
    int32_t F;              /* input_size                                                   */
    int32_t M;              /* num_metrics (global, >= 2: the reference crashes at 1)       */
    int32_t H;              /* must be 128                                                  */
    int32_t Q;              /* must be 3                                                    */
    float   quantiles[8];   /* first Q used                                                 */
    float   dropout_p;      /* training only                                                */
    int32_t engine;         /* DR_ENGINE_*                                                  */
    int32_t device;         /* CUDA ordinal                                                 */
    int32_t rank, world;    /* expert shard: this handle owns experts [rank*M/world, ...)   */
    int32_t dtype;          /* DR_DTYPE_*: arithmetic of dr_train_step          */
} dr_config;

/* ---- lifecycle: replaces QuantileRNN(...).to(device)  (estimate.py:60) ---- */
int  dr_create(const dr_config* cfg, dr_model** out);
void dr_destroy(dr_model* m);
const char* dr_last_error(const dr_model* m);
int  dr_version(void);
/* 1 if this build contains the given DR_ENGINE_* recurrence engine */
int  dr_has_engine(int32_t engine);

/* optional: run on a caller-owned CUDA stream. use_caller_stream == 0 restores the handle's own
 * (non-blocking) stream; != 0 adopts cuda_stream (a cudaStream_t; NULL = the legacy default stream) */
int  dr_set_stream(dr_model* m, void* cuda_stream, int32_t use_caller_stream);
/* experts owned by this handle: [*lo, *hi) */
int  dr_local_experts(const dr_model* m, int32_t* lo, int32_t* hi);

/* ---- weights: replaces load_state_dict()/state_dict() (the reference never saves) ----
 * blob is the FULL model (M experts); the handle keeps only its shard. dr_get_weights writes
 * the local shard into its slot of a full-size blob and leaves the rest untouched. */
int  dr_load_weights(dr_model* m, const float* host_blob, size_t n_floats);
int  dr_get_weights (dr_model* m, float* host_blob, size_t n_floats);

/* ---- forward: replaces QuantileRNN.forward in eval mode (qrnn.py:28-56; estimate.py:91) ----
 * x [B,T,F] fp32 -> out [B,T,M,Q] fp32.  Requires world == 1. */
int  dr_forward    (dr_model* m, const float* x_host, int32_t B, int32_t T, float* out_host);
int  dr_forward_dev(dr_model* m, const float* x_dev,  int32_t B, int32_t T, float* out_dev);

/* ---- forward straight from the raw traffic series (on-device windowing) ----
 * Replaces sliding_window (utils.py:4-5) + the strided evaluation loop (estimate.py:85-91): window k is
 * series[k*stride : k*stride + W], for starts 0, stride, ... < N - W (like the reference, the last full window
 * is dropped).  series [N,F] -> out [dr_series_windows(N,W,stride), W, M, Q].  The W-fold blow-up of the
 * windowed tensor never exists: the operand packers read the series with a window stride. */
int  dr_series_windows(int32_t N, int32_t W, int32_t stride);
int  dr_forward_series    (dr_model* m, const float* series_host, int32_t N, int32_t W, int32_t stride, float* out_host);
int  dr_forward_series_dev(dr_model* m, const float* series_dev,  int32_t N, int32_t W, int32_t stride, float* out_dev);

/* ---- optional output transform fused into the head kernel ----
 * out = max(out, clamp_min) * scale[m] + offset[m]   — estimate.py:96 (clamp on the normalised forecast) and
 * estimate.py:101-102 (undo normalization_minmax: scale = max-min, offset = min).  scale/offset: host [M].
 * Pass NULLs to switch it off (the default). */
int  dr_set_output_transform(dr_model* m, const float* scale_host, const float* offset_host, float clamp_min);

/* ---- expert-sharded forward, device pointers, async ----
 *   1. dr_forward_local_dev : local bi-GRUs.  Writes S_dev = sum over LOCAL experts of their GRU
 *      outputs (dr_s_elems(B,T) floats, stored k-group major [T][2H/4][round_up(B,128)][4]; the
 *      layout is opaque to the caller, an element-wise all-reduce is all it needs), and
 *      out_local_dev[B,T,M_loc,Q] = own-expert part of the heads (FFMA engine; the tcgen05 engine keeps
 *      its partials in a library workspace until step 3 writes out_local_dev).
 *   2. caller all-reduces (sum) S_dev across ranks (torch.distributed / NCCL).
 *   3. dr_forward_heads_dev : adds the cross-expert-mean term and bias into out_local_dev.
 *   4. caller all-gathers out_local into gathered[world][B,T,M_loc,Q];
 *      dr_interleave_dev reorders it to the reference layout out[B,T,M,Q] (qrnn.py:55).
 *   Pairing: a heads call (3 or 3'+4') finds the own-expert partials of its chunk by the S_dev pointer it is given —
 *   pass the SAME S_dev buffer to step 1 and step 3.  Up to 4 local phases may be in flight (different S_dev buffers,
 *   any streams, any interleaving); a heads call with an S_dev no live local phase produced returns DR_ESTATE. */
int64_t dr_s_elems(int32_t B, int32_t T);     /* floats in S_dev for a [B,T] call */
int  dr_forward_local_dev(dr_model* m, const float* x_dev, int32_t B, int32_t T,
                          float* S_dev, float* out_local_dev);
int  dr_forward_heads_dev(dr_model* m, const float* S_dev, int32_t B, int32_t T,
                          float* out_local_dev);
/*   3'+4' fused (tcgen05 engine): instead of steps 3-4, dr_forward_heads_p2p_dev writes this rank's forecasts straight
 *      into EVERY rank's full tensor out[Bfull,T,M,Q] (reference layout): out_ptrs[w] is rank w's buffer, peer-mapped
 *      into this process (CUDA IPC / torch symmetric memory); rows row0..row0+B-1.  The stores to peers travel over
 *      NVLink from inside the head kernel — no all-gather, no interleave.  The caller barriers across ranks before
 *      (buffers free) and after (writes landed). */
int  dr_forward_heads_p2p_dev(dr_model* m, const float* S_dev, int32_t B, int32_t T,
                              void* const* out_ptrs, int32_t n_ptrs, int64_t row0);
/*   4'' copy-engine variant of step 4: after dr_forward_heads_dev, dr_scatter_forecasts_dev places this rank's
 *      out_local [B,T,M_loc,Q] into columns [rank*M_loc, ...) of EVERY rank's full tensor (rows row0..) with strided
 *      2-D peer copies — all-gather and interleave done by the DMA engines, no SM time. Same barrier rules as 3'+4'. */
int  dr_scatter_forecasts_dev(dr_model* m, const float* out_local_dev, int32_t B, int32_t T,
                              void* const* out_ptrs, int32_t n_ptrs, int64_t row0);
int  dr_interleave_dev   (dr_model* m, const float* gathered_dev, int32_t B, int32_t T,
                          float* out_dev);

/* ---- expert-sharded forward as ONE call per rank (no NCCL, no host-side collectives; csrc/dr_comm.cu) ----
 * The cross-expert mean (qrnn.py:46-52) needs the sum S over ALL experts.  Every rank owns an "arena" of device memory that
 * its peers map; the recurrence kernel runs as one launch and flags finished 256-window tiles, copy engines move each tile's
 * partial S to the peers, the head kernel adds the partials in rank order (bit-identical on every rank) and the forecast
 * columns are placed into every rank's stacked tensor [B,T,M,Q] (qrnn.py:55) by strided 2-D peer copies.
 *   dr_comm_init   : size this rank's arena for calls up to [Bmax, T]; returns its CUDA IPC handle (64 bytes; for one
 *                    process per GPU) and its device pointer (for several handles inside one process).  Collective in the
 *                    sense that every rank must have returned from it before any rank calls dr_comm_attach.
 *                    The arena is ONE allocation per (process, device), made by the first dr_comm_init on that device and
 *                    kept until the process exits (CUDA IPC rule: an exported allocation must not be freed while an importer
 *                    may still map it).  Later handles on the same device reuse it and the peer mappings; it must fit the
 *                    largest of them — create that handle first or set DR_COMM_ARENA_GB — else DR_ENOMEM.
 *   dr_comm_attach : map the peers' arenas — `ipc_handles` = world x 64 bytes in rank order (own entry ignored), or
 *                    `arena_ptrs` = world device pointers of the same process (peer access is enabled here).  Peers mapped by
 *                    an earlier handle of this process are reused.
 *   dr_comm_detach : drain this handle's sharded forwards (all of its internal streams).  Teardown order across ranks: every
 *                    rank detaches, the host synchronises the ranks, then every rank calls dr_destroy (or dr_comm_init
 *                    again) — so that no peer is still copying into this rank when its handle goes away.  dr_destroy
 *                    detaches by itself, which is enough for several handles inside one process.
 *   Handles that share a device share the arena: drain one handle's sharded forwards (dr_comm_detach or a stream
 *   synchronise on every rank) before another handle on that device issues one.  A sharded forward that fails part-way on one
 *   rank (any code other than DR_OK) leaves its peers waiting for signals that will not come: tear the group down.
 *   dr_forward_sharded_dev : x_dev [B,T,F] (replicated) -> *out_dev = the stacked forecasts [B,T,M,Q] in this rank's arena.
 *                    Asynchronous on the handle's stream (dr_set_stream).  The tensor stays valid until the second-next
 *                    sharded forward on this handle (two tensors alternate); copy it if it must live longer.
 *   dr_forward_sharded     : host entry point.  x_host [B,T,F]; this rank's forecast columns are written to
 *                    out_host[b, t, rank*M/world ..., :] where out_host is the FULL [B,T,M,Q] tensor (row pitch M*Q floats;
 *                    ranks of one host may share it), chunk by chunk while later chunks still compute; returns when they are
 *                    there.  *out_dev (nullable) as above.  tcgen05 engine (input_size <= 64) only. */
int64_t dr_comm_arena_bytes(const dr_model* m, int32_t Bmax, int32_t T);
int  dr_comm_init  (dr_model* m, int32_t Bmax, int32_t T, void* ipc_handle_out /* 64 bytes, nullable */, void** arena_ptr_out /* nullable */);
int  dr_comm_attach(dr_model* m, const void* ipc_handles /* nullable */, void* const* arena_ptrs /* nullable */);
int  dr_comm_detach(dr_model* m);
int  dr_forward_sharded_dev(dr_model* m, const float* x_dev, int32_t B, int32_t T, float** out_dev);
int  dr_forward_sharded    (dr_model* m, const float* x_host, int32_t B, int32_t T, float* out_host, float** out_dev);
/*   Two forwards in flight: dr_forward_sharded_dev makes the handle's stream wait for its result before it returns control of
 *   that stream, so the next forward's recurrence starts only after this forward's exchange tail.  A serving loop that has
 *   the next batch ready issues it first and consumes the previous result afterwards:
 *     dr_forward_sharded_issue_dev(batch n+1) -> *ticket;   dr_forward_sharded_wait(ticket of batch n);   use result n
 *   _issue_dev enqueues everything but does not touch the handle's stream after reading x; _wait makes the handle's stream
 *   wait for that forward (one of the last two). */
int  dr_forward_sharded_issue_dev(dr_model* m, const float* x_dev, int32_t B, int32_t T, float** out_dev, int32_t* ticket);
int  dr_forward_sharded_wait     (dr_model* m, int32_t ticket);

/* ---- loss: replaces QuantileRNN.quantile_loss (qrnn.py:58-67) ----
 * out [B,T,M,Q], y [B,T,M] -> scalar pinball loss (mean over M of mean over B,T of sum over Q) */
int  dr_quantile_loss    (dr_model* m, const float* out_host, const float* y_host,
                          int32_t B, int32_t T, float* loss_host);
int  dr_quantile_loss_dev(dr_model* m, const float* out_dev, const float* y_dev,
                          int32_t B, int32_t T, float* loss_dev);

/* ---- training step: replaces one iteration of estimate.py:67-74 ----
 *   outputs = model(inputs)  [train mode: dropout(p) on the GRU outputs, qrnn.py:43]
 *   loss = model.quantile_loss(outputs, labels); optimizer.zero_grad(); loss.backward(); optimizer.step()
 * with torch.optim.Adam(lr) defaults (estimate.py:61); Adam state lives in the handle.
 * x [B,T,F], y [B,T,M].  dropout_mask (nullable): replayed keep-mask, uint8 [M,B,T,2H] in the
 * reference's rnn_out element order (parity tests); NULL -> counter-based RNG keyed by `seed`.
 * dr_train_step(_dev) need world == 1 (sharded handles: dr_train_begin_dev / dr_train_advance below).  With engine AUTO / TC
 * the two recurrences and the weight-gradient reductions run on tcgen05 (split-fp16, fp32 accumulate), with engine FFMA the
 * whole step is exact fp32 on the CUDA cores.  dr_get_grads returns the gradients of the last step (blob order). */
int  dr_train_step    (dr_model* m, const float* x_host, const float* y_host, int32_t B, int32_t T,
                       const uint8_t* dropout_mask_host, uint64_t seed, float lr, float* loss_host);
int  dr_train_step_dev(dr_model* m, const float* x_dev, const float* y_dev, int32_t B, int32_t T,
                       const uint8_t* dropout_mask_dev, uint64_t seed, float lr, float* loss_dev,
                       float* out_dev /* [B,T,M,Q] train-mode forecasts */);
int  dr_get_grads     (dr_model* m, float* host_blob, size_t n_floats);
/* Large batches run as micro-batches (exact: gradients add).  By default the size comes from the free device memory; on
 * expert-sharded handles every rank must use the SAME size (the cross-rank sums are per micro-batch): set it explicitly.
 * windows == 0 restores the default.  The bf16 engine rounds up to whole 128-window tiles. */
int  dr_train_set_microbatch(dr_model* m, int32_t windows);
/* Expert-sharded training (world > 1): the step is a resumable state machine.  dr_train_begin_dev takes the replicated
 * x [B,T,F], this rank's label columns y [B,T,M_loc], the GLOBAL replayed mask (or NULL + seed) and out_dev
 * [B,T,M_loc,Q]; dr_train_advance runs kernels until *kind == 0 (step done: local gradients applied by Adam) or
 * *kind == 1: all-reduce(sum) `*count` elements at device pointer `*ptr` (*dtype 0 = fp32, 1 = fp64) across the
 * ranks on the handle's stream, then call dr_train_advance again.  Requests: S after each micro-batch forward, the loss
 * scalar, the head adjoint before each micro-batch backward.  No gradient all-reduce exists: weights are expert-local. */
int  dr_train_begin_dev(dr_model* m, const float* x_dev, const float* y_local_dev, int32_t B, int32_t T,
                        const uint8_t* dropout_mask_dev, uint64_t seed, float lr, float* loss_dev, float* out_local_dev);
int  dr_train_advance  (dr_model* m, int32_t* kind, void** ptr, int64_t* count, int32_t* dtype);

/* ---- test/diagnostic access to prepared tensors (not on the hot path) ----
 * what: "mask" [M_loc,F] (qrnn.py:34), "S" (dr_s_elems floats, layout above) of the last
 * dr_forward/dr_forward_dev, "ct", "bias4". Returns DR_EINVAL for unknown names. */
int  dr_debug_read(dr_model* m, const char* what, float* host_buf, size_t n_floats);
/* per-kernel device timing (CUDA events on the launching stream). dr_profile(m,1) clears the
 * record and makes every forward (up to 256) record events around the recurrence kernel and the
 * head kernel; dr_profile_read synchronises and returns how many forwards were recorded and the
 * SUM of their kernel durations in milliseconds. */
int  dr_profile(dr_model* m, int32_t enable);
int  dr_profile_read(dr_model* m, int32_t* n_forwards, float* gru_ms_sum, float* head_ms_sum);
/* hardware probe of the tensor-core building blocks (one tcgen05 GEMM tile; see
 * csrc/dr_tc_probe.cu). variant bit0: A from TMEM, bit1: cta_group::2. Host pointers. */
int  dr_tc_probe(int32_t variant, const void* a_host, size_t a_bytes, const void* b_host, size_t b_bytes,
                 int32_t N, int32_t K, int32_t flags, float* d_out_host);
/* same for operands in the MN-major SW128 canonical layout (reduced index slowest), the form the bf16 weight-gradient
 * kernel consumes: D[128 x N] = sum_k A[k][m] B[k][n].  params: 8 x uint32 {a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep,
 * a_mn, b_mn} (bytes / major bits); a, b: ready shared-memory images.  See csrc/dr_tc_probe.cu. */
int  dr_tc_probe_mn(const void* a_host, size_t a_bytes, const void* b_host, size_t b_bytes, int32_t N, int32_t K,
                    const uint32_t* params, float* d_out_host);
/* number of kernels this handle has launched since creation (bench.py's gpu_launches) */
int64_t dr_launch_count(const dr_model* m);
/* name of the GRU engine the last forward used: "ffma" or "tcgen05" */
const char* dr_last_engine(const dr_model* m);

#ifdef __cplusplus
}
#endif
#endif /* DEEPREST_B200_H */
