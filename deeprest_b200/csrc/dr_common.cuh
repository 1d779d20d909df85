// Internal declarations shared by the sm_100a kernels and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string>
#include "../../include/deeprest_b200.h"

#define DR_GATES 3                 // (r, z, n) — torch.nn.GRU order, qrnn.py:24
#define DR_2H (2 * DR_H)
#define DR_PROF_MAX 256
#define DR_KC 16                   // K-chunk of the FFMA weight stream; Fp is a multiple of it

// ---- per-expert offsets inside the reference-order blob (mirror of layout.py) ----
struct DrBlobOffsets {
    int mask_w1, mask_b1, mask_w2, mask_b2;
    int w_ih[2], w_hh[2], b_ih[2], b_hh[2];   // [0]=forward, [1]=reverse direction
    int head_w, head_b;
    int per_expert;
};

__host__ __device__ inline DrBlobOffsets dr_blob_offsets(int F) {
    DrBlobOffsets o;
    int off = 0;
    o.mask_w1 = off; off += DR_H;
    o.mask_b1 = off; off += DR_H;
    o.mask_w2 = off; off += F * DR_H;
    o.mask_b2 = off; off += F;
    for (int d = 0; d < 2; ++d) {
        o.w_ih[d] = off; off += 3 * DR_H * F;
        o.w_hh[d] = off; off += 3 * DR_H * DR_H;
        o.b_ih[d] = off; off += 3 * DR_H;
        o.b_hh[d] = off; off += 3 * DR_H;
    }
    o.head_w = off; off += DR_Q * 4 * DR_H;
    o.head_b = off; off += DR_Q;
    o.per_expert = off;
    return o;
}

struct dr_model {
    dr_config cfg;
    int e_lo, e_hi, M_loc;
    int Fp;                         // F padded up to a multiple of DR_KC (zero weights/inputs)
    DrBlobOffsets off;
    bool loaded;

    // device state
    float* d_blob;                  // local shard, reference order          [M_loc * per_expert]
    float* d_mask;                  // softmax feature mask, qrnn.py:34       [M_loc, F]
    float* d_wf;                    // FFMA weight stream                     [M_loc,2,2,KT,3,64]
    float* d_bias4;                 // (b_ir+b_hr, b_iz+b_hz, b_in, b_hn)     [M_loc,2,4,H]
    float* d_ct;                    // own-expert head coeff C - A/(M-1)      [M_loc,2,Q,H]
    float* d_abar;                  // mean-term head coeff A/(M-1)           [M_loc*Q, 2H]
    float* d_hb;                    // head bias                              [M_loc*Q]
    float* d_wihm;                  // mask-folded input weights W_ih*diag(mask)   [2][M_loc][3H][F] (training GEMM)
    float* d_grad;                  // gradients of the last train step, reference blob order [M_loc*per_expert]
    float* d_adam_m; float* d_adam_v; int64_t adam_step;
    void*  train_ws;                // training workspace (dr_train.cu)
    int    train_mb;                // micro-batch override in windows (dr_train_set_microbatch), 0 = sized from free memory
    void*  d_dropmask; size_t dropmask_cap;
    __nv_bfloat16* d_wtc;           // tcgen05 weight image (hi/lo bf16)      see dr_gru_tc.cu
    size_t wtc_bytes;

    // workspace, grown on demand
    // operand-image workspaces rotate through DR_WS_SLOTS slots (one per dr_forward_local_dev call), so that up to
    // DR_WS_SLOTS forwards may be in flight on different streams (the chunked multi-GPU pipeline); d_xT / d_xtc
    // point at the slot of the call being issued
    float* ws_xT[4]; size_t ws_xT_cap[4]; void* ws_xtc[4]; size_t ws_xtc_cap[4]; int ws_slot;
    float* ws_p[4]; size_t ws_p_cap[4];
    const float* ws_key[4]; bool ws_tc[4];   // the S_dev a slot was issued for, and whether that call ran on the tcgen05 engine:
                                             // dr_forward_heads_dev finds its partials by the S pointer it is given
    uint8_t* d_himg;                // K2 weight images for the tcgen05 head GEMM (dr_head_tc.cu)
    float* d_p; size_t p_cap; bool p_live;   // own-expert head partials of the call being issued / consumed (tcgen05 engine)
    float* d_xT;   size_t xT_cap;   // x transposed to [T, Fp, Bp]            (FFMA engine)
    void*  d_xtc;  size_t xtc_cap;  // x split to bf16 hi/lo [T, Bp, Fp]      (tcgen05 engine)
    void*  d_xtc_tr; size_t xtc_tr_cap;   // x image of the training micro-batch (own buffer: d_xtc above aliases a ring slot)
    void*  d_whT;    size_t whT_cap;      // W_hh^T split-fp16 images for the tensor-core backward recurrence (dr_gru_bwd_tc.cu)
    float* d_S;    size_t S_cap;    // [T][2H/4][Bp][4]
    float* d_out;  size_t out_cap;  // staging for host entry points
    float* d_xin;  size_t xin_cap;  // staging for host entry points
    float* d_loss;                  // 1 float + partials
    float* d_y;    size_t y_cap;

    cudaStream_t stream;
    cudaStream_t own_stream;
    long long x_bstride;            // 0 = dense windows [B,T,F]; else floats between window starts (series mode, N1)
    float* d_dn; bool dn_on; float dn_clamp;   // optional output transform: scale[M_loc] | offset[M_loc] (N2)
    unsigned int* tile_count; unsigned int* tile_flag; unsigned int tile_value;   // per-tile completion signal of the next K1 launch (dr_comm.cu), or null
    void* comm;                     // expert-sharded forward state (DrComm, dr_comm.cu)
    int tc_xdrop;                   // precision probe: drop one split term of the x-part (dr_debug_read "tc_xdrop1"/"tc_xdrop2"/"tc_xdrop0")
    unsigned long long* d_tc_dbg;   // optional cycle breakdown of the tcgen05 kernel (dr_debug_read "tc_timing")
    cudaStream_t copy_stream;       // H2D/D2H of the pipelined host entry point
    cudaStream_t stream2;           // second compute stream of the pipelined host entry point
    cudaEvent_t ev_pipe[10];
    float* ws_S[4]; size_t ws_S_cap[4];
    int64_t launches;
    bool profile;
    int prof_n;                     // forwards recorded since dr_profile(m,1)
    cudaEvent_t* ev;                // [DR_PROF_MAX][4]: gru start/stop, head start/stop
    int sm_count;
    const char* last_engine;
    std::string err;
};

int dr_fail(dr_model* m, int code, const std::string& msg);
int dr_cuda_fail(dr_model* m, cudaError_t e, const char* what);

#define DR_CUDA(m, call)                                                        \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) return dr_cuda_fail((m), _e, #call);             \
    } while (0)

// profile events of the forward being recorded (nullptr when profiling is off or the ring is full)
inline cudaEvent_t* dr_prof_slot(dr_model* m) {
    return (m->profile && m->ev && m->prof_n < DR_PROF_MAX) ? m->ev + 4 * m->prof_n : nullptr;
}

// S (cross-expert sum) is stored k-group major: [T][2H/4][dr_s_rows(B)][4] floats
inline int dr_s_rows(int B) { return (B + 127) / 128 * 128; }
inline size_t dr_s_floats(int B, int T) { return (size_t)T * DR_2H * dr_s_rows(B); }

// grows *ptr to at least `bytes` (device). Contents are NOT preserved.
int dr_reserve(dr_model* m, void** ptr, size_t* cap, size_t bytes);

// ---- kernels (host launchers) ----
// dr_prep.cu
int dr_launch_prep(dr_model* m);
// dr_gru_ffma.cu
int dr_ffma_rows_per_thread(int B);
int dr_launch_xT(dr_model* m, const float* x_dev, int B, int T, int Bp);
int dr_launch_gru_ffma(dr_model* m, int B, int T, int Bp, float* S_dev, float* out_local_dev);
// dr_gru_tc.cu
bool dr_tc_built();
bool dr_tc_supported(const dr_model* m, int B, int T);
int dr_tc_prep_weights(dr_model* m);
int dr_launch_gru_tc(dr_model* m, const float* x_dev, int B, int T, float* S_dev, float* out_local_dev);
// training forward of one micro-batch on the tensor-core engine: saves (r,z,n), q, h per step in dr_train.cu's layout
int dr_launch_gru_tc_train(dr_model* m, const float* x_dev, int Bm, int T, float* rzn, float* q, float* hs, long long dir_stride_rows,
                           int lane_major);
// weight-gradient reductions C[z][m][n] (+)= sum_k A[z][k][m] B[z][k][n] on the tensor-core engine (dr_wgrad_tc.cu)
bool dr_wgrad_tc_ok(int M, int N, int K);
int dr_grad_scale_log2(float inv_n);
int dr_launch_wgrad_tc(dr_model* m, int ndir, const float* const* A, long long lda, long long bsA, const int* acol3 /* nullable */,
                       const float* const* B, long long ldb, long long bsB, float* const* C, long long ldc, long long bsC,
                       int M, int N, int K, int batch, int a_scale_log2, int b_scale_log2, int accumulate);
// backward recurrence of one micro-batch on the tensor-core engine (dr_gru_bwd_tc.cu): gate adjoints + dh chain, in place
int dr_launch_gru_bwd_tc(dr_model* m, const float* rzn, const float* q, const float* hs, const float* dhout, float* g4,
                         long long dir_rows, long long dho_dir_rows, int Bm, int T, float inv_n, int in_lane_major);
// bf16 training engine (dr_gru_tc16.cu, dr_gru_bwd16.cu, dr_wgrad16.cu; layout in dr_t16.cuh)
size_t dr_t16_wimg_bytes(int M_loc);
size_t dr_t16_whT_bytes(int M_loc);
int dr_t16_pack_weights(dr_model* m, uint8_t* wimg);
int dr_t16_pack_whT(dr_model* m, uint8_t* img);
int dr_t16_pack_x(dr_model* m, const float* x_mb, int Bm, int T, uint8_t* ximg);
int dr_launch_gru_tc16(dr_model* m, const uint8_t* wimg, const uint8_t* ximg, int Bm, int T, float* S, float* P,
                       uint8_t* gate, uint8_t* himg, const uint8_t* mask, uint64_t seed, int b0, int Bfull);
int dr_launch_gru_bwd16(dr_model* m, const uint8_t* whT, uint8_t* gate, const uint8_t* himg, const float* dy, const float* gbar,
                        int Bm, int T, const uint8_t* mask, uint64_t seed, int b0, int Bfull);
int dr_launch_wgrad16(dr_model* m, const uint8_t* gate, const uint8_t* himg, const uint8_t* ximg, const uint8_t* zero,
                      float* P, int Bm, int T);
// dr_train.cu
int dr_train_step_impl(dr_model* m, const float* x, const float* y, int B, int T, const uint8_t* mask, uint64_t seed,
                       float lr, float* loss_dev, float* out_dev);
int dr_train_begin_impl(dr_model* m, const float* x, const float* y, int B, int T, const uint8_t* mask, uint64_t seed,
                        float lr, float* loss_dev, float* out_dev);
int dr_train_advance_impl(dr_model* m, int* kind, void** ptr, long long* count, int* dtype);
void dr_train_free(dr_model* m);
// dr_head_tc.cu
int dr_head_tc_prep(dr_model* m);
int dr_launch_heads_tc(dr_model* m, const float* S_dev, int B, int T, float* out_local_dev);
int dr_launch_heads_tc_dst(dr_model* m, const float* S_dev, int B, int T, void* const* dst_ptrs, int n_dst, long long row0);
// head kernel over a chunk of a larger batch with the cross-expert sum given as `nsrc` partial sums (expert-sharded forward):
// source w covers the chunk's windows at src[w] + layout [T][64][src_rows[w]][4], first window src_b0[w]; partials are added in
// source order (rank order: bit-identical on every rank).  P is the full batch's partial workspace (p_tiles 128-window tiles
// per step), the chunk starts at tile p_tile0.
int dr_launch_heads_tc_multi(dr_model* m, const float* const* src, const int* src_rows, const int* src_b0, int nsrc,
                             const float* P, int p_tiles, int p_tile0, int B, int T, float* out_local);
void dr_comm_free(dr_model* m);
// dr_head.cu
int dr_launch_heads(dr_model* m, const float* S_dev, int B, int T, float* out_local_dev);
int dr_launch_interleave(dr_model* m, const float* gathered, int B, int T, float* out);
int dr_launch_loss(dr_model* m, const float* out_dev, const float* y_dev, int B, int T, float* loss_dev);

// ---- small device helpers ----
__device__ __forceinline__ float dr_sigmoid(float v) {
    // 1/(1+exp(-v)); ex2.approx + rcp.approx keep the abs error ~1e-7 (well under the 1e-6 atol)
    return __fdividef(1.0f, 1.0f + __expf(-v));
}
__device__ __forceinline__ float dr_tanh(float v) {
    // tanh(v) = 1 - 2/(1+exp(2v)); saturates cleanly to +-1 for |v| large
    return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * v));
}
__device__ __forceinline__ void dr_red_add_v4(float* addr, float a, float b, float c, float d) {
    // vectorised fire-and-forget fp32 reduction (sm_90+): one L2 atomic transaction per 16 B
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void dr_red_add(float* addr, float a) {
    asm volatile("red.global.add.f32 [%0], %1;" :: "l"(addr), "f"(a) : "memory");
}
