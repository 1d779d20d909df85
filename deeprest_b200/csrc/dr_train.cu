// Training step of the estimator (estimate.py:67-74): forward in train mode (dropout on the GRU
// outputs, qrnn.py:43), pinball loss (qrnn.py:58-67), the backward that torch autograd derives
// from qrnn.py:28-67 (adjoints listed in SURVEY §8a "Backward"), and Adam (estimate.py:61).
//
// Two engines behind the same state machine and the same activation layout (TIME-MAJOR rows, row = (e*T + t)*Bm + b):
//  * tensor-core engine (default, cfg.engine != FFMA): the forward recurrence is the inference kernel instantiated to
//    save (r,z,n), q, h per step (dr_gru_tc.cu, F <= 64); the reverse-time chain (gate adjoints + dh_{t-1} = dh*z +
//    dgh W_hh) is one persistent tcgen05 kernel (dr_gru_bwd_tc.cu); the two weight-gradient reductions over all (t,b)
//    rows are tcgen05 GEMMs whose fp32 sources are converted to split-fp16 operand images in flight (dr_wgrad_tc.cu).
//    Split-fp16 (hi+lo, 3 tensor passes, fp32 accumulate) keeps every gradient tensor within the fp32 parity bounds of
//    tests/test_gpu_train.py; gradient operands are scaled by an exact power of two chosen from 1/(M*B*T).
//  * FFMA engine (cfg.engine == FFMA): exact fp32 on the CUDA cores — batched GEMMs (dr_bgemm_kernel) + elementwise
//    kernels, one launch per time step and direction for both recurrences.  Cross-check of the other engine.
// Dropout, the cross-expert sum, heads, loss, head/bias/mask gradients and Adam are fp32 elementwise kernels in both.
// Memory: 4.5 KB per expert-window-step-direction, so large batches are processed in micro-batches of Bm windows sized
// from the free HBM (exact: windows are independent given the cross-expert sums, and gradients add).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "dr_common.cuh"
#include "dr_t16.cuh"

namespace {

constexpr int GTM = 128, GTN = 64, GK = 16;          // CTA tile 128 x 64, K-chunk 16, 256 threads x (8 x 4) outputs
constexpr int GLDA = GTM + 4, GLDB = GTN + 4;

// C[z](m,n) = beta*C[z](m,n) + sum_k A[z](m,k) * B[z](k,n), arbitrary element strides
struct Gemm {
    const float* A; const float* B; float* C;
    int M, N, K;
    long sam, sak, sbk, sbn, scm, scn;
    long bsA, bsB, bsC;
    float beta;
};

// Register-tiled fp32 GEMM with a register-prefetched, double-buffered shared-memory pipeline.  Loads walk whichever of
// (m|n) and k is the unit-stride dimension of each operand, so all five uses (input projection, per-step recurrent
// products, dgh·W_hh, and the two weight-gradient reductions over (t,b)) stay coalesced without a transpose pass.
__global__ void __launch_bounds__(256) dr_bgemm_kernel(Gemm g) {
    __shared__ __align__(16) float As[2][GK][GLDA];
    __shared__ __align__(16) float Bs[2][GK][GLDB];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;        // ty: 8 rows each, tx: 4 cols each
    const int m0 = blockIdx.y * GTM, n0 = blockIdx.x * GTN;
    const float* A = g.A + (size_t)blockIdx.z * g.bsA;
    const float* B = g.B + (size_t)blockIdx.z * g.bsB;
    float* C = g.C + (size_t)blockIdx.z * g.bsC;
    const bool a_m_fast = g.sam <= g.sak, b_n_fast = g.sbn <= g.sbk;
    float acc[8][4] = {};
    float ra[8], rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                  // A tile: 128 x 16 = 2048 elements, 8 per thread
            int idx = tid + i * 256;
            int m = a_m_fast ? idx % GTM : idx / GK, k = a_m_fast ? idx / GTM : idx % GK;
            ra[i] = (m0 + m < g.M && k0 + k < g.K) ? A[(long)(m0 + m) * g.sam + (long)(k0 + k) * g.sak] : 0.0f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {                                  // B tile: 16 x 64 = 1024 elements, 4 per thread
            int idx = tid + i * 256;
            int n = b_n_fast ? idx % GTN : idx / GK, k = b_n_fast ? idx / GTN : idx % GK;
            rb[i] = (n0 + n < g.N && k0 + k < g.K) ? B[(long)(k0 + k) * g.sbk + (long)(n0 + n) * g.sbn] : 0.0f;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int idx = tid + i * 256;
            int m = a_m_fast ? idx % GTM : idx / GK, k = a_m_fast ? idx / GTM : idx % GK;
            As[buf][k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int idx = tid + i * 256;
            int n = b_n_fast ? idx % GTN : idx / GK, k = b_n_fast ? idx / GTN : idx % GK;
            Bs[buf][k][n] = rb[i];
        }
    };
    fetch(0);
    stash(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = 0; k0 < g.K; k0 += GK) {
        const bool more = k0 + GK < g.K;
        if (more) fetch(k0 + GK);                                      // global loads in flight during the FFMAs
#pragma unroll
        for (int kk = 0; kk < GK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (more) {
            stash(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m0 + ty * 8 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float* c = C + (long)m * g.scm + (long)n * g.scn;
            *c = (g.beta == 0.0f) ? acc[i][j] : g.beta * *c + acc[i][j];
        }
    }
}

// ---- dropout keep decision: replayed mask (parity) or counter-based hash (production) -------------
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// idx = ((e_glob*B + b)*T + t)*2H + k   — the reference's rnn_out element order [M,B,T,2H]
__device__ __forceinline__ float keep_scale(const uint8_t* mask, uint64_t seed, size_t idx, float p, float inv_keep) {
    if (mask) return mask[idx] ? inv_keep : 0.0f;
    if (p <= 0.0f) return 1.0f;
    float u = (float)(mix64(seed * 0x9E3779B97F4A7C15ull + idx) >> 40) * (1.0f / 16777216.0f);
    return (u >= p) ? inv_keep : 0.0f;
}

// x [B,T,F] (rows b0..b0+Bm) -> xt [(t*Bm+b)][F]
__global__ void dr_time_major_kernel(const float* __restrict__ x, float* __restrict__ xt, int b0, int Bm, int T, int F) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * Bm * F;
    if (i >= total) return;
    int f = (int)(i % F); size_t r = i / F;
    int b = (int)(r % Bm); int t = (int)(r / Bm);
    xt[i] = x[((size_t)(b0 + b) * T + t) * F + f];
}

// gi[e][(t,b)][3H] += b_ih   (after the input-projection GEMM)
__global__ void dr_add_bias_kernel(float* __restrict__ gi, const float* __restrict__ blob, int off_b, int pe, size_t rows_per_e, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c = (int)(i % (3 * DR_H)); size_t r = i / (3 * DR_H);
    int e = (int)(r / rows_per_e);
    gi[i] += blob[(size_t)e * pe + off_b + c];
}

// one forward time step for every local expert of one direction (GRU equations)
//   gi : [e][(t,b)][3H] (x projection + b_ih)      gh: [e][b][3H] (h_{t-1} W_hh^T, no bias)
//   saves rzn[e][(t,b)][3H] = (r,z,n), q[e][(t,b)][H] = W_hn h + b_hn, hs[e][(t,b)][H] = h_t
__global__ void dr_gate_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ hprev,
                                   const float* __restrict__ blob, int off_bhh, int pe,
                                   float* __restrict__ rzn, float* __restrict__ q, float* __restrict__ hs,
                                   int t, int Bm, int T, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int j = (int)(i % DR_H); size_t r = i / DR_H;
    int b = (int)(r % Bm); int e = (int)(r / Bm);
    size_t row = ((size_t)e * T + t) * Bm + b;
    const float* g = gi + row * 3 * DR_H;
    const float* h3 = gh + ((size_t)e * Bm + b) * 3 * DR_H;
    const float* bh = blob + (size_t)e * pe + off_bhh;
    float rr = 1.0f / (1.0f + expf(-(g[j] + h3[j] + bh[j])));
    float zz = 1.0f / (1.0f + expf(-(g[DR_H + j] + h3[DR_H + j] + bh[DR_H + j])));
    float qq = h3[2 * DR_H + j] + bh[2 * DR_H + j];
    float nn = tanhf(g[2 * DR_H + j] + rr * qq);
    float hp = hprev ? hprev[((size_t)e * T * Bm + b) * DR_H + j] : 0.0f;   // hprev points at hs[e=0][(t_prev,0)]
    float hn = __fadd_rn(__fmul_rn(__fsub_rn(hp, nn), zz), nn);
    rzn[row * 3 * DR_H + j] = rr; rzn[row * 3 * DR_H + DR_H + j] = zz; rzn[row * 3 * DR_H + 2 * DR_H + j] = nn;
    q[row * DR_H + j] = qq;
    hs[row * DR_H + j] = hn;
}

// S[(t,b)][2H] = sum over local experts of dropout(h)   (deterministic: no atomics)
__global__ void dr_sum_experts_kernel(const float* __restrict__ hs0, const float* __restrict__ hs1, const uint8_t* __restrict__ mask, uint64_t seed, float p,
                                      float* __restrict__ S, int M_loc, int e_lo, int B, int b0, int Bm, int T, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int k = (int)(i % DR_2H); size_t r = i / DR_2H;
    int b = (int)(r % Bm); int t = (int)(r / Bm);
    int d = k / DR_H, j = k % DR_H;
    float inv_keep = 1.0f / (1.0f - p);
    float acc = 0.0f;
    for (int e = 0; e < M_loc; ++e) {
        float h = (d ? hs1 : hs0)[(((size_t)e * T + t) * Bm + b) * DR_H + j];
        size_t midx = (((size_t)(e_lo + e) * B + b0 + b) * T + t) * DR_2H + k;
        acc += h * keep_scale(mask, seed, midx, p, inv_keep);
    }
    S[i] = acc;
}

// out[b,t,e,q] = sum_k Ct[e][d][q][j] r~_e + sum_k Abar[e][q][k] S + hb ; one warp per (t,b,e)
__global__ void dr_head_fwd_kernel(const float* __restrict__ hs0, const float* __restrict__ hs1, const float* __restrict__ S, const uint8_t* __restrict__ mask,
                                   uint64_t seed, float p, const float* __restrict__ ct, const float* __restrict__ abar,
                                   const float* __restrict__ hb, float* __restrict__ out,
                                   int M_loc, int e_lo, int B, int b0, int Bm, int T) {
    size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    size_t total = (size_t)T * Bm * M_loc;
    if (w >= total) return;
    int e = (int)(w % M_loc); size_t r = w / M_loc;
    int b = (int)(r % Bm); int t = (int)(r / Bm);
    float inv_keep = 1.0f / (1.0f - p);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane; k < DR_2H; k += 32) {
        int d = k / DR_H, j = k % DR_H;
        float h = (d ? hs1 : hs0)[(((size_t)e * T + t) * Bm + b) * DR_H + j];
        size_t midx = (((size_t)(e_lo + e) * B + b0 + b) * T + t) * DR_2H + k;
        float rt = h * keep_scale(mask, seed, midx, p, inv_keep);
        float s = S[((size_t)t * Bm + b) * DR_2H + k];
        const float* c = ct + ((size_t)(e * 2 + d) * DR_Q) * DR_H + j;
        const float* a = abar + ((size_t)e * DR_Q) * DR_2H + k;
        a0 += c[0] * rt + a[0] * s;
        a1 += c[DR_H] * rt + a[DR_2H] * s;
        a2 += c[2 * DR_H] * rt + a[2 * DR_2H] * s;
    }
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, o);
        a1 += __shfl_xor_sync(0xffffffffu, a1, o);
        a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    }
    if (lane == 0) {
        float* o = out + (((size_t)(b0 + b) * T + t) * M_loc + e) * DR_Q;
        o[0] = a0 + hb[e * DR_Q]; o[1] = a1 + hb[e * DR_Q + 1]; o[2] = a2 + hb[e * DR_Q + 2];
    }
}

// dL/dout (qrnn.py:58-67 through torch.max's tie rule) and the loss partial sum
// tm_B > 0: dy is written TIME-major, dy[(t*tm_B + b)][M_loc][Q] (bf16 engine: the windows of a step are contiguous for the
// backward kernels); out / y stay in the reference's [b][t][m] order.
__global__ void dr_loss_grad_kernel(const float* __restrict__ out, const float* __restrict__ y, float* __restrict__ dy,
                                    size_t n_rm, float q0, float q1, float q2, float inv_n, double* acc,
                                    int tm_B = 0, int tm_T = 0, int tm_M = 0) {
    float local = 0.0f;
    const float qs[3] = {q0, q1, q2};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rm; i += (size_t)gridDim.x * blockDim.x) {
        float yy = y[i];
        size_t o = i;
        if (tm_B > 0) {
            const size_t e_ = i % tm_M, bt = i / tm_M;
            const size_t t_ = bt % tm_T, b_ = bt / tm_T;
            o = (t_ * tm_B + b_) * tm_M + e_;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float e = yy - out[i * DR_Q + k];
            local += fmaxf((qs[k] - 1.0f) * e, qs[k] * e);
            float g = (e < 0.0f) ? (1.0f - qs[k]) : (e > 0.0f) ? -qs[k] : (0.5f - qs[k]);
            dy[o * DR_Q + k] = g * inv_n;
        }
    }
    __shared__ double red[256];
    red[threadIdx.x] = (double)local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) { if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) atomicAdd(acc, red[0]);
}

// Gbar[(t,b)][k] = sum_{e,q} Abar[e][q][k] dy[b,t,e,q]
__global__ void dr_gbar_kernel(const float* __restrict__ dy, const float* __restrict__ abar, float* __restrict__ gbar,
                               int M_loc, int b0, int Bm, int T, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int k = (int)(i % DR_2H); size_t r = i / DR_2H;
    int b = (int)(r % Bm); int t = (int)(r / Bm);
    const float* d = dy + ((size_t)(b0 + b) * T + t) * M_loc * DR_Q;
    float acc = 0.0f;
    for (int eq = 0; eq < M_loc * DR_Q; ++eq) acc = fmaf(abar[(size_t)eq * DR_2H + k], d[eq], acc);
    gbar[i] = acc;
}

// d(h_t) that enters the GRU backward: dropout adjoint of (Ct^T dy + Gbar)     dhout[d][e][(t,b)][H]
__global__ void dr_dhout_kernel(const float* __restrict__ dy, const float* __restrict__ gbar, const float* __restrict__ ct,
                                const uint8_t* __restrict__ mask, uint64_t seed, float p, float* __restrict__ dhout,
                                int M_loc, int e_lo, int B, int b0, int Bm, int T, size_t total, int lane_major) {
    size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;           // one thread = 4 consecutive hidden units
    if (i4 * 4 >= total) return;
    // row-major output [d][e][t][b][H] (i4 walks it in order), or lane-major [d][e][t][H/4][b][4] for the tensor-core
    // backward kernel (windows fastest: the stores stay coalesced, the strided gbar reads come from L2)
    int j, b; size_t r;
    if (lane_major) { b = (int)(i4 % Bm); r = i4 / Bm; j = (int)(r % (DR_H / 4)) * 4; r /= DR_H / 4; }
    else            { j = (int)(i4 % (DR_H / 4)) * 4; r = i4 / (DR_H / 4); b = (int)(r % Bm); r /= Bm; }
    int t = (int)(r % T); r /= T;
    int e = (int)(r % M_loc); int d = (int)(r / M_loc);
    int k = d * DR_H + j;
    const float* dd = dy + (((size_t)(b0 + b) * T + t) * M_loc + e) * DR_Q;
    const float* c = ct + ((size_t)(e * 2 + d) * DR_Q) * DR_H + j;
    const float d0 = dd[0], d1 = dd[1], d2 = dd[2];
    const float4 c0 = *reinterpret_cast<const float4*>(c), c1 = *reinterpret_cast<const float4*>(c + DR_H), c2 = *reinterpret_cast<const float4*>(c + 2 * DR_H);
    const float4 gb = *reinterpret_cast<const float4*>(gbar + ((size_t)t * Bm + b) * DR_2H + k);
    const size_t midx = (((size_t)(e_lo + e) * B + b0 + b) * T + t) * DR_2H + k;
    const float ik = 1.0f / (1.0f - p);
    float4 o;                                                            // same operation order per element as before
    o.x = (c0.x * d0 + c1.x * d1 + c2.x * d2 + gb.x) * keep_scale(mask, seed, midx, p, ik);
    o.y = (c0.y * d0 + c1.y * d1 + c2.y * d2 + gb.y) * keep_scale(mask, seed, midx + 1, p, ik);
    o.z = (c0.z * d0 + c1.z * d1 + c2.z * d2 + gb.z) * keep_scale(mask, seed, midx + 2, p, ik);
    o.w = (c0.w * d0 + c1.w * d1 + c2.w * d2 + gb.w) * keep_scale(mask, seed, midx + 3, p, ik);
    *reinterpret_cast<float4*>(dhout + i4 * 4) = o;
}

// head weight gradients: U[e][q][k] = sum_{t,b} dy r~ ; V = sum dy S ; db = sum dy.   grid (M_loc, chunks), 256 threads = k
__global__ void dr_head_grad_kernel(const float* __restrict__ hs0, const float* __restrict__ hs1, const float* __restrict__ S, const float* __restrict__ dy,
                                    const uint8_t* __restrict__ mask, uint64_t seed, float p, float* __restrict__ gblob,
                                    int off_hw, int off_hb, int pe, float inv_m1,
                                    int M_loc, int e_lo, int B, int b0, int Bm, int T, int rows_per_chunk) {
    int e = blockIdx.x, k = threadIdx.x;
    int d = k / DR_H, j = k % DR_H;
    size_t r0 = (size_t)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
    size_t rows = (size_t)T * Bm;
    if (r1 > rows) r1 = rows;
    float inv_keep = 1.0f / (1.0f - p);
    float u[3] = {0, 0, 0}, v[3] = {0, 0, 0}, sb[3] = {0, 0, 0};
    for (size_t r = r0; r < r1; ++r) {
        int t = (int)(r / Bm), b = (int)(r % Bm);
        float h = (d ? hs1 : hs0)[(((size_t)e * T + t) * Bm + b) * DR_H + j];
        size_t midx = (((size_t)(e_lo + e) * B + b0 + b) * T + t) * DR_2H + k;
        float rt = h * keep_scale(mask, seed, midx, p, inv_keep);
        float s = S[r * DR_2H + k];
        const float* dd = dy + (((size_t)(b0 + b) * T + t) * M_loc + e) * DR_Q;
#pragma unroll
        for (int q = 0; q < 3; ++q) { u[q] = fmaf(dd[q], rt, u[q]); v[q] = fmaf(dd[q], s, v[q]); if (k == 0) sb[q] += dd[q]; }
    }
    float* hw = gblob + (size_t)e * pe + off_hw;      // [Q][4H]: cols 0..2H-1 = A (mean part), 2H.. = C (own part)
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        atomicAdd(hw + (size_t)q * 4 * DR_H + DR_2H + k, u[q]);
        atomicAdd(hw + (size_t)q * 4 * DR_H + k, (v[q] - u[q]) * inv_m1);
        if (k == 0) atomicAdd(gblob + (size_t)e * pe + off_hb + q, sb[q]);
    }
}

// one backward time step for one direction: gate adjoints in place (rzn -> dgh, gi -> dgi), dh carry *= z
__global__ void dr_gate_bwd_kernel(float* __restrict__ rzn, float* __restrict__ gi, const float* __restrict__ q,
                                   const float* __restrict__ hprev, const float* __restrict__ dhout, float* __restrict__ dhc,
                                   int t, int Bm, int T, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int j = (int)(i % DR_H); size_t r = i / DR_H;
    int b = (int)(r % Bm); int e = (int)(r / Bm);
    size_t row = ((size_t)e * T + t) * Bm + b;
    float rr = rzn[row * 3 * DR_H + j], zz = rzn[row * 3 * DR_H + DR_H + j], nn = rzn[row * 3 * DR_H + 2 * DR_H + j];
    float qq = q[row * DR_H + j];
    float hp = hprev ? hprev[((size_t)e * T * Bm + b) * DR_H + j] : 0.0f;
    float dh = dhc[((size_t)e * Bm + b) * DR_H + j] + dhout[row * DR_H + j];
    float dn = dh * (1.0f - zz);
    float dz = dh * (hp - nn);
    float dan = dn * (1.0f - nn * nn);
    float dr = dan * qq;
    float dq = dan * rr;
    float daz = dz * zz * (1.0f - zz);
    float dar = dr * rr * (1.0f - rr);
    gi[row * 3 * DR_H + j] = dar; gi[row * 3 * DR_H + DR_H + j] = daz; gi[row * 3 * DR_H + 2 * DR_H + j] = dan;     // dgi
    rzn[row * 3 * DR_H + j] = dar; rzn[row * 3 * DR_H + DR_H + j] = daz; rzn[row * 3 * DR_H + 2 * DR_H + j] = dq;  // dgh
    dhc[((size_t)e * Bm + b) * DR_H + j] = dh * zz;        // + dgh W_hh is added by the GEMM that follows
}

// column sums: dst[e*pe + off + c] += sum over rows of src[e][row][c]   grid (M_loc, chunks), 3H threads
__global__ void dr_colsum_kernel(const float* __restrict__ src, float* __restrict__ gblob, int off, int pe,
                                 size_t rows, int rows_per_chunk) {
    int e = blockIdx.x, c = threadIdx.x;
    size_t r0 = (size_t)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
    if (r1 > rows) r1 = rows;
    float acc = 0.0f;
    for (size_t r = r0; r < r1; ++r) acc += src[((size_t)e * rows + r) * 3 * DR_H + c];
    atomicAdd(gblob + (size_t)e * pe + off + c, acc);
}

// tensor-core path: g4[e][row][4H] = (da_r, da_z, da_n, dq);  db_ih += sums of columns 0..3H-1,  db_hh += (da_r, da_z, dq)
__global__ void dr_colsum4_kernel(const float* __restrict__ g4, float* __restrict__ gblob, int off_bih, int off_bhh, int pe,
                                  size_t rows, int rows_per_chunk) {
    int e = blockIdx.x, c = threadIdx.x;                                  // 4H threads
    size_t r0 = (size_t)blockIdx.y * rows_per_chunk, r1 = r0 + rows_per_chunk;
    if (r1 > rows) r1 = rows;
    float acc = 0.0f;
    for (size_t r = r0; r < r1; ++r) acc += g4[((size_t)e * rows + r) * 4 * DR_H + c];
    float* gb = gblob + (size_t)e * pe;
    if (c < 3 * DR_H) atomicAdd(gb + off_bih + c, acc);
    if (c < 2 * DR_H) atomicAdd(gb + off_bhh + c, acc);
    else if (c >= 3 * DR_H) atomicAdd(gb + off_bhh + c - DR_H, acc);
}

// from P[e][3H][F] = sum dgi (x) x :  dW_ih += P * mask ;  dmask[e][f] += sum_i W_ih[i][f] P[i][f]
__global__ void dr_wih_grad_kernel(const float* __restrict__ P, const float* __restrict__ blob, const float* __restrict__ mask,
                                   float* __restrict__ gblob, float* __restrict__ dmask, int off_wih, int pe, int F) {
    int e = blockIdx.x;
    for (int f = threadIdx.x; f < F; f += blockDim.x) {
        float mk = mask[(size_t)e * F + f], acc = 0.0f;
        for (int i = 0; i < 3 * DR_H; ++i) {
            float pv = P[((size_t)e * 3 * DR_H + i) * F + f];
            gblob[(size_t)e * pe + off_wih + (size_t)i * F + f] += pv * mk;
            acc = fmaf(blob[(size_t)e * pe + off_wih + (size_t)i * F + f], pv, acc);
        }
        dmask[(size_t)e * F + f] += acc;
    }
}

// softmax / Linear / ReLU / Linear adjoints of the mask MLP (qrnn.py:34); one block per expert, H threads
__global__ void dr_mask_bwd_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F, const float* __restrict__ mask,
                                   const float* __restrict__ dmask, float* __restrict__ gblob) {
    extern __shared__ float sm[];            // dlogit[F]
    int e = blockIdx.x, tid = threadIdx.x;
    const float* ex = blob + (size_t)e * off.per_expert;
    float* gx = gblob + (size_t)e * off.per_expert;
    float dot = 0.0f;
    for (int f = 0; f < F; ++f) dot += mask[(size_t)e * F + f] * dmask[(size_t)e * F + f];
    for (int f = tid; f < F; f += blockDim.x) sm[f] = mask[(size_t)e * F + f] * (dmask[(size_t)e * F + f] - dot);
    __syncthreads();
    float pre = ex[off.mask_w1 + tid] + ex[off.mask_b1 + tid];
    float hid = fmaxf(pre, 0.0f);
    float dhid = 0.0f;
    for (int f = 0; f < F; ++f) {
        gx[off.mask_w2 + (size_t)f * DR_H + tid] += sm[f] * hid;
        dhid = fmaf(ex[off.mask_w2 + (size_t)f * DR_H + tid], sm[f], dhid);
    }
    if (pre <= 0.0f) dhid = 0.0f;
    gx[off.mask_w1 + tid] += dhid;
    gx[off.mask_b1 + tid] += dhid;
    for (int f = tid; f < F; f += blockDim.x) gx[off.mask_b2 + f] += sm[f];
}

// torch.optim.Adam defaults, in torch's operation order (estimate.py:61,74)
__global__ void dr_adam_kernel(float* __restrict__ w, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                               size_t n, float lr_over_bc1, float inv_sqrt_bc2, float b1, float b2, float eps) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float gi = g[i];
    float mi = b1 * m[i] + (1.0f - b1) * gi;
    float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    w[i] = w[i] - lr_over_bc1 * (mi / denom);
}

__global__ void dr_finish_loss_kernel(const double* acc, double inv_n, float* loss) { *loss = (float)(*acc * inv_n); }

// bf16 engine: head weight gradients from the h images.  U[e][q][k] = sum_{t,b} dy r~ ; V = sum dy S ; db = sum dy
// (dC = U, dA = (V - U)/(M-1); qrnn.py:46-54 differentiated).  grid (M_loc, chunks of time steps); a thread owns 4 consecutive
// columns k of [fwd | rev] (one 8-byte h load, one float4 of S in its k-group-major layout, one dropout hash) and every 4th
// window; the loops run over (step, 128-window tile, window) so that every address is a base plus a constant stride.
__global__ void __launch_bounds__(256) dr_head_grad16_kernel(const uint8_t* __restrict__ himg, const float* __restrict__ S, const float* __restrict__ dy,
                                                              drt16::Drop drop, float* __restrict__ gblob, int off_hw, int off_hb, int pe, float inv_m1,
                                                              int M_loc, int e_lo, int Bfull, int b0, int Bm, int T, int steps_per_chunk) {
    __shared__ float red[3][2][3][256];               // [slot 1..3][u|v][q][k]
    __shared__ float redb[4][3];
    const int e = blockIdx.x, kg = threadIdx.x & 63, slot = threadIdx.x >> 6;
    const int k0 = kg * 4, d = k0 >> 7, j0 = k0 & 127;
    const int ntiles = (Bm + 127) >> 7, Bp = ntiles * 128;
    const int t0 = blockIdx.y * steps_per_chunk, t1 = min(T, t0 + steps_per_chunk);
    float u[3][4], v[3][4], sb[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int q = 0; q < 3; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) { u[q][i] = 0.f; v[q][i] = 0.f; }
    const size_t col_off = (size_t)(j0 >> 6) * drt16::kColBlk + (size_t)(j0 & 7) * 2;
    const int chunk = (j0 & 63) >> 3;
    const size_t dy_bstride = (size_t)M_loc * DR_Q;
    const size_t drop_bstride = (size_t)T * DR_2H;
    for (int t = t0; t < t1; ++t) {
        for (int tile = 0; tile < ntiles; ++tile) {
            const uint8_t* hblk = himg + drt16::blk_index(d, e, t, tile, M_loc, T, ntiles) * drt16::kHImg + col_off;
            const float* sblk = S + (((size_t)t * 64 + kg) * Bp + tile * 128) * 4;
            const int nb = min(128, Bm - tile * 128);
            const float* dblk = dy + ((size_t)t * Bm + tile * 128) * M_loc * DR_Q + (size_t)e * DR_Q;     // dy is time-major
            const size_t dbase = (((size_t)(e_lo + e) * Bfull + b0 + tile * 128) * T + t) * DR_2H + k0;
#pragma unroll 2
            for (int r = slot; r < nb; r += 4) {
                const uint2 hw2 = __ldg(reinterpret_cast<const uint2*>(hblk + drt16::img_off(r, chunk)));
                const float4 s4 = __ldg(reinterpret_cast<const float4*>(sblk + (size_t)r * 4));
                const float* dd = dblk + (size_t)r * dy_bstride;
                const float g0 = __ldg(dd), g1 = __ldg(dd + 1), g2 = __ldg(dd + 2);
                const uint32_t kb = drt16::keep4(drop, dbase + (size_t)r * drop_bstride);
                const float2 h01 = drt16::unpack_bf2(hw2.x), h23 = drt16::unpack_bf2(hw2.y);
                const float rt[4] = {(kb & 1u) ? h01.x * drop.inv_keep : 0.f, (kb & 2u) ? h01.y * drop.inv_keep : 0.f,
                                     (kb & 4u) ? h23.x * drop.inv_keep : 0.f, (kb & 8u) ? h23.y * drop.inv_keep : 0.f};
                const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
                const float gq[3] = {g0, g1, g2};
#pragma unroll
                for (int q = 0; q < 3; ++q) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { u[q][i] = fmaf(gq[q], rt[i], u[q][i]); v[q][i] = fmaf(gq[q], sv[i], v[q][i]); }
                    if (kg == 0) sb[q] += gq[q];
                }
            }
        }
    }
    if (slot > 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) { red[slot - 1][0][q][k0 + i] = u[q][i]; red[slot - 1][1][q][k0 + i] = v[q][i]; }
    }
    if (kg == 0) { redb[slot][0] = sb[0]; redb[slot][1] = sb[1]; redb[slot][2] = sb[2]; }
    __syncthreads();
    if (slot == 0) {
        float* hw = gblob + (size_t)e * pe + off_hw;      // [Q][4H]: cols 0..2H-1 = A (mean part), 2H.. = C (own part)
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int k = k0 + i;
                const float uu = u[q][i] + red[0][0][q][k] + red[1][0][q][k] + red[2][0][q][k];
                const float vv = v[q][i] + red[0][1][q][k] + red[1][1][q][k] + red[2][1][q][k];
                atomicAdd(hw + (size_t)q * 4 * DR_H + DR_2H + k, uu);
                atomicAdd(hw + (size_t)q * 4 * DR_H + k, (vv - uu) * inv_m1);
            }
            if (kg == 0) atomicAdd(gblob + (size_t)e * pe + off_hb + q, redb[0][q] + redb[1][q] + redb[2][q] + redb[3][q]);
        }
    }
}

inline unsigned nblk(size_t n, int t = 256) { return (unsigned)((n + t - 1) / t); }

int gemm(dr_model* m, const Gemm& g, int batch) {
    if (g.M <= 0 || g.N <= 0 || batch <= 0) return DR_OK;
    dim3 grid((g.N + GTN - 1) / GTN, (g.M + GTM - 1) / GTM, batch);
    dr_bgemm_kernel<<<grid, 256, 0, m->stream>>>(g);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// The step is a resumable state machine: dr_train_advance runs kernels until the step is done or — on an
// expert-sharded handle — a cross-rank sum is needed (S after the forward of a micro-batch, the loss scalar, the
// head adjoint G-bar before the backward of a micro-batch).  The host performs that all-reduce on the returned
// device buffer and calls advance again (torch.distributed in the Python host; NCCL in a C/Go host).  With
// world == 1 no request is ever produced and dr_train_step simply drives the machine to completion.
enum { TS_IDLE = 0, TS_MB_BEGIN, TS_AFTER_S, TS_AFTER_LOSS, TS_AFTER_G, TS_FINISH,
       T16_MB_BEGIN, T16_AFTER_S, T16_AFTER_G, T16_AFTER_LOSS };     // bf16 engine: one pass per micro-batch

struct dr_train_ws {
    float *xt, *gi, *rzn, *q, *hs, *dhout, *gh, *dhc, *S, *gbar, *dy, *P, *dmask;
    size_t cap_rows; int cap_B; int cap_T;
    // bf16 engine (dr_config.dtype == DR_DTYPE_BF16): operand images instead of fp32 activation rows (dr_t16.cuh)
    uint8_t *w16, *whT16, *x16, *gate16, *h16, *zero16;
    float *S16, *P16, *dy16, *gbar16, *Px16;
    int cap16_tiles, cap16_T;
    // state of the step in flight
    int stage, pass, mb, n_mb, Bm, B, T;
    const float *x, *y; const uint8_t* mask; uint64_t seed; float lr; float* loss_dev; float* out_dev;
};

static int ws_alloc(dr_model* m, float** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    cudaError_t e = cudaMalloc((void**)p, (n ? n : 4) * sizeof(float));
    if (e != cudaSuccess) return dr_cuda_fail(m, e, "cudaMalloc(training workspace)");
    return DR_OK;
}

// ---------------------------------------------------------------------------------------------
// bf16 engine (dr_config.dtype == DR_DTYPE_BF16): single-pass bf16 tensor-core kernels, bf16 activation images
// (csrc/dr_gru_tc16.cu, dr_gru_bwd16.cu, dr_wgrad16.cu, layout in dr_t16.cuh).  One pass per micro-batch of whole
// 128-window tiles:   x image -> train-mode forward (dropout, S, head partials, saved images) -> [S all-reduce] -> heads
// (the inference K2) -> pinball-loss gradient of these windows -> G-bar GEMM -> [G-bar all-reduce] -> reverse chain ->
// weight-gradient GEMM -> head gradients.   The loss gradient of a window needs only that window's forecasts (the 1/(M·B·T)
// of the mean is a constant), so no forward is ever recomputed; the loss VALUE is summed across micro-batches.
static int ws16_alloc(dr_model* m, void** p, size_t bytes) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e != cudaSuccess) return dr_cuda_fail(m, e, "cudaMalloc(bf16 training workspace)");
    return DR_OK;
}

static int train16_begin(dr_model* m, dr_train_ws* ws, int B, int T) {
    const int F = m->cfg.F, Ml = m->M_loc;
    if (F > 64) return dr_fail(m, DR_EUNSUPPORTED, "bf16 training engine: input_size must be <= 64 (one 64-wide K block of x)");
    if (!m->d_himg && Ml) return dr_fail(m, DR_ESTATE, "bf16 training engine: head images missing (load weights first)");
    const int tiles_all = (B + 127) / 128;
    const size_t ngrp = (size_t)(Ml * DR_Q + 15) / 16;
    // bytes per 128-window tile of a micro-batch
    const size_t per_tile = (size_t)2 * Ml * T * (drt16::kGateImg + drt16::kHImg) + (size_t)T * drt16::kColBlk +
                            (size_t)T * DR_2H * 128 * 4 /* S */ + (size_t)T * ngrp * 4 * 16 * 128 * 4 /* P */ +
                            (size_t)128 * T * Ml * DR_Q * 4 /* dy */ + (size_t)128 * T * DR_2H * 4 /* gbar */;
    size_t budget = (size_t)24 << 30;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            size_t held = (size_t)ws->cap16_tiles * per_tile;
            budget = std::max<size_t>((size_t)4 << 30, (free_b + held) * 6 / 10);
        }
    }
    int mb_tiles = (int)std::min<size_t>((size_t)tiles_all, std::max<size_t>(1, budget / std::max<size_t>(per_tile, 1)));
    if (const char* ov = getenv("DR_TRAIN_MICROBATCH")) { int v = atoi(ov); if (v >= 1) mb_tiles = std::min(tiles_all, (v + 127) / 128); }   // test hook
    if (m->train_mb >= 1) mb_tiles = std::min(tiles_all, (m->train_mb + 127) / 128);
    if (ws->cap16_tiles >= mb_tiles && ws->cap16_T == T) mb_tiles = std::min(tiles_all, ws->cap16_tiles);   // keep what is already there
    else {
        // allocate; if the device refuses (memory, or address space when peers are mapped) fall back to fewer tiles per micro-batch
        for (;;) {
            ws->cap16_tiles = 0; ws->cap16_T = 0;
            const size_t nt = (size_t)mb_tiles;
            int rc;
            if (!((rc = ws16_alloc(m, (void**)&ws->w16, dr_t16_wimg_bytes(Ml))) || (rc = ws16_alloc(m, (void**)&ws->whT16, dr_t16_whT_bytes(Ml))) ||
                  (rc = ws16_alloc(m, (void**)&ws->x16, (size_t)T * nt * drt16::kColBlk)) ||
                  (rc = ws16_alloc(m, (void**)&ws->gate16, (size_t)2 * Ml * T * nt * drt16::kGateImg)) ||
                  (rc = ws16_alloc(m, (void**)&ws->h16, (size_t)2 * Ml * T * nt * drt16::kHImg)) ||
                  (rc = ws16_alloc(m, (void**)&ws->zero16, drt16::kColBlk)) ||
                  (rc = ws16_alloc(m, (void**)&ws->S16, (size_t)T * DR_2H * nt * 128 * sizeof(float))) ||
                  (rc = ws16_alloc(m, (void**)&ws->P16, (size_t)T * nt * ngrp * 4 * 16 * 128 * sizeof(float))) ||
                  (rc = ws16_alloc(m, (void**)&ws->dy16, (size_t)nt * 128 * T * Ml * DR_Q * sizeof(float))) ||
                  (rc = ws16_alloc(m, (void**)&ws->gbar16, (size_t)nt * 128 * T * DR_2H * sizeof(float))) ||
                  (rc = ws16_alloc(m, (void**)&ws->Px16, (size_t)2 * Ml * 3 * DR_H * F * sizeof(float))) ||
                  (rc = ws_alloc(m, &ws->dmask, (size_t)Ml * F))))
                break;
            cudaGetLastError();
            void** all16[] = {(void**)&ws->w16, (void**)&ws->whT16, (void**)&ws->x16, (void**)&ws->gate16, (void**)&ws->h16, (void**)&ws->zero16,
                              (void**)&ws->S16, (void**)&ws->P16, (void**)&ws->dy16, (void**)&ws->gbar16, (void**)&ws->Px16};
            for (void** pp : all16) if (*pp) { cudaFree(*pp); *pp = nullptr; }
            if (mb_tiles == 1) return rc;
            mb_tiles = (mb_tiles + 1) / 2;
        }
        DR_CUDA(m, cudaMemsetAsync(ws->zero16, 0, drt16::kColBlk, m->stream));
        ws->cap16_tiles = mb_tiles; ws->cap16_T = T;
    }
    const int Bm = std::min(B, mb_tiles * 128);
    DR_CUDA(m, cudaMemsetAsync(ws->Px16, 0, (size_t)2 * Ml * 3 * DR_H * F * sizeof(float), m->stream));
    DR_CUDA(m, cudaMemsetAsync(ws->dmask, 0, (size_t)Ml * F * sizeof(float), m->stream));
    int rc = dr_t16_pack_weights(m, ws->w16);
    if (rc) return rc;
    if ((rc = dr_t16_pack_whT(m, ws->whT16))) return rc;
    ws->stage = T16_MB_BEGIN; ws->pass = 0; ws->mb = 0; ws->Bm = Bm; ws->n_mb = (B + Bm - 1) / Bm;
    return DR_OK;
}

int dr_train_begin_impl(dr_model* m, const float* x, const float* y, int B, int T, const uint8_t* mask, uint64_t seed,
                        float lr, float* loss_dev, float* out_dev) {
    const int F = m->cfg.F, Ml = m->M_loc, pe = m->off.per_expert;
    if (m->cfg.dropout_p >= 1.0f) return dr_fail(m, DR_EINVAL, "dropout_p must be < 1");
    // micro-batch size from a memory budget (4.5 KB per expert-window-step-direction, see header)
    size_t per_window = (size_t)2 * Ml * T * (3 * DR_H + 4 * DR_H + DR_H * 3) * sizeof(float);   // rzn, gi/g4, q+hs+dhout
    // budget: half of the HBM that is free right now plus what this workspace already holds, at least 8 GB (a 180 GB
    // B200 gives ~85 GB: config-2-sized micro-batches of 256+ windows stay in one piece, so nothing is recomputed and
    // the per-step GEMMs of the backward chain run on full 128-row tiles)
    size_t budget = (size_t)24 << 30;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
            const dr_train_ws* cur = reinterpret_cast<const dr_train_ws*>(m->train_ws);
            size_t held = cur ? (size_t)2 * Ml * cur->cap_rows * (3 * DR_H + 4 * DR_H + DR_H * 3) * sizeof(float) : 0;
            budget = std::max<size_t>((size_t)8 << 30, (free_b + held) / 2);
        }
    }
    int Bm = (int)std::min<size_t>((size_t)B, std::max<size_t>(1, budget / std::max<size_t>(per_window, 1)));
    if (const char* ov = getenv("DR_TRAIN_MICROBATCH")) { int v = atoi(ov); if (v >= 1) Bm = std::min(B, v); }   // test hook
    if (m->train_mb >= 1) Bm = std::min(B, m->train_mb);
    dr_train_ws* ws = reinterpret_cast<dr_train_ws*>(m->train_ws);
    if (!ws) { ws = new dr_train_ws(); memset(ws, 0, sizeof(*ws)); m->train_ws = ws; }
    if (ws->stage != TS_IDLE) return dr_fail(m, DR_ESTATE, "a training step is already in flight on this handle");
    const bool bf16 = m->cfg.dtype == DR_DTYPE_BF16;
    if (bf16 && m->cfg.engine == DR_ENGINE_FFMA)
        return dr_fail(m, DR_EINVAL, "dtype bf16 selects the single-pass tensor-core training engine; it cannot be combined with engine FFMA");
    size_t rows = (size_t)T * Bm, E2 = (size_t)2 * Ml;
    if (bf16) rows = 0;                                 // the fp32 activation rows are not used
    if (!bf16 && (ws->cap_rows < rows || ws->cap_B < B || ws->cap_T != T)) {
        int rc;
        ws->cap_rows = 0; ws->cap_B = 0; ws->cap_T = 0;      // a failed allocation below leaves the workspace marked empty
        if ((rc = ws_alloc(m, &ws->xt, rows * F)) || (rc = ws_alloc(m, &ws->gi, E2 * rows * 4 * DR_H)) ||
            (rc = ws_alloc(m, &ws->rzn, E2 * rows * 3 * DR_H)) || (rc = ws_alloc(m, &ws->q, E2 * rows * DR_H)) ||
            (rc = ws_alloc(m, &ws->hs, E2 * rows * DR_H)) || (rc = ws_alloc(m, &ws->dhout, E2 * rows * DR_H)) ||
            (rc = ws_alloc(m, &ws->gh, (size_t)Ml * Bm * 3 * DR_H)) || (rc = ws_alloc(m, &ws->dhc, (size_t)Ml * Bm * DR_H)) ||
            (rc = ws_alloc(m, &ws->S, rows * DR_2H)) || (rc = ws_alloc(m, &ws->gbar, rows * DR_2H)) ||
            (rc = ws_alloc(m, &ws->dy, (size_t)B * T * Ml * DR_Q)) || (rc = ws_alloc(m, &ws->P, (size_t)2 * Ml * 3 * DR_H * F)) ||
            (rc = ws_alloc(m, &ws->dmask, (size_t)Ml * F)))
            return rc;
        ws->cap_rows = rows; ws->cap_B = B; ws->cap_T = T;
    }
    size_t nblob = (size_t)Ml * pe;
    if (!m->d_grad) {
        DR_CUDA(m, cudaMalloc((void**)&m->d_grad, std::max<size_t>(nblob, 1) * sizeof(float)));
        DR_CUDA(m, cudaMalloc((void**)&m->d_adam_m, std::max<size_t>(nblob, 1) * sizeof(float)));
        DR_CUDA(m, cudaMalloc((void**)&m->d_adam_v, std::max<size_t>(nblob, 1) * sizeof(float)));
        DR_CUDA(m, cudaMemsetAsync(m->d_adam_m, 0, nblob * sizeof(float), m->stream));
        DR_CUDA(m, cudaMemsetAsync(m->d_adam_v, 0, nblob * sizeof(float), m->stream));
        m->adam_step = 0;
    }
    DR_CUDA(m, cudaMemsetAsync(m->d_grad, 0, nblob * sizeof(float), m->stream));
    if (!bf16) DR_CUDA(m, cudaMemsetAsync(ws->dmask, 0, (size_t)Ml * F * sizeof(float), m->stream));
    DR_CUDA(m, cudaMemsetAsync(m->d_loss, 0, sizeof(double), m->stream));
    ws->stage = TS_MB_BEGIN; ws->pass = 0; ws->mb = 0; ws->Bm = Bm; ws->n_mb = (B + Bm - 1) / Bm; ws->B = B; ws->T = T;
    if (bf16) {
        int rc = train16_begin(m, ws, B, T);
        if (rc) { ws->stage = TS_IDLE; return rc; }
    }
    // the dropout draw is a pure function of (seed, element): mix the optimizer step in so that a loop calling
    // train_step with a constant seed still draws a fresh mask every iteration (nn.Dropout does, qrnn.py:43)
    ws->x = x; ws->y = y; ws->mask = mask; ws->seed = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(m->adam_step + 1); ws->lr = lr; ws->loss_dev = loss_dev; ws->out_dev = out_dev;
    return DR_OK;
}

// forward of micro-batch mb up to the local partial of S
static int train_forward_mb(dr_model* m, dr_train_ws* ws, int b0, int bm) {
    const int F = m->cfg.F, Ml = m->M_loc, pe = m->off.per_expert, T = ws->T, B = ws->B;
    const float p = m->cfg.dropout_p;
    cudaStream_t st = m->stream;
    const size_t r = (size_t)T * bm;
    const size_t ed_stride = (size_t)Ml * ws->cap_rows;
    dr_time_major_kernel<<<nblk(r * F), 256, 0, st>>>(ws->x, ws->xt, b0, bm, T, F);
    // Tensor-core engine: the whole recurrence of the micro-batch (both directions, all experts) is ONE launch of the
    // tcgen05 kernel of the inference path, instantiated to save (r,z,n), q and h per step (csrc/dr_gru_tc.cu).
    const bool tc_fwd = m->cfg.engine != DR_ENGINE_FFMA && dr_tc_supported(m, bm, T) && m->d_wtc != nullptr;   // == train_tc_forward()
    if (tc_fwd) {
        // rzn and q window-contiguous (lane-major): only the tensor-core backward kernel reads them
        int rc = dr_launch_gru_tc_train(m, ws->x + (size_t)b0 * T * F, bm, T, ws->rzn, ws->q, ws->hs, (long long)ed_stride, 1);
        if (rc) return rc;
    }
    for (int d = 0; d < 2 && !tc_fwd; ++d) {
        float* gi = ws->gi + d * ed_stride * 3 * DR_H;
        float* rzn = ws->rzn + d * ed_stride * 3 * DR_H;
        float* q = ws->q + d * ed_stride * DR_H;
        float* hs = ws->hs + d * ed_stride * DR_H;
        // gi = x (W_ih diag(mask))^T + b_ih for all steps at once (the folded weights come from K0: d_wihm)
        Gemm g{ws->xt, m->d_wihm + (size_t)d * Ml * 3 * DR_H * F, gi, (int)r, 3 * DR_H, F,
               F, 1, 1, F, 3 * DR_H, 1, 0, (long)3 * DR_H * F, (long)T * bm * 3 * DR_H, 0.0f};
        int rc = gemm(m, g, Ml);
        if (rc) return rc;
        size_t tot = (size_t)Ml * r * 3 * DR_H;
        if (tot) dr_add_bias_kernel<<<nblk(tot), 256, 0, st>>>(gi, m->d_blob, m->off.b_ih[d], pe, r, tot);
        for (int s = 0; s < T; ++s) {
            const int t = d ? (T - 1 - s) : s, tp = d ? t + 1 : t - 1;
            const float* hprev = (s == 0) ? nullptr : hs + (size_t)tp * bm * DR_H;
            if (s == 0) {
                DR_CUDA(m, cudaMemsetAsync(ws->gh, 0, (size_t)Ml * bm * 3 * DR_H * sizeof(float), st));
            } else {
                Gemm gh{hprev, m->d_blob + m->off.w_hh[d], ws->gh, bm, 3 * DR_H, DR_H,
                        DR_H, 1, 1, DR_H, 3 * DR_H, 1, (long)T * bm * DR_H, (long)pe, (long)bm * 3 * DR_H, 0.0f};
                rc = gemm(m, gh, Ml);
                if (rc) return rc;
            }
            size_t tg = (size_t)Ml * bm * DR_H;
            if (tg) dr_gate_fwd_kernel<<<nblk(tg), 256, 0, st>>>(gi, ws->gh, hprev, m->d_blob, m->off.b_hh[d], pe, rzn, q, hs, t, bm, T, tg);
        }
    }
    size_t ts = r * DR_2H;
    dr_sum_experts_kernel<<<nblk(ts), 256, 0, st>>>(ws->hs, ws->hs + ed_stride * DR_H, ws->mask, ws->seed, p, ws->S, Ml, m->e_lo, B, b0, bm, T, ts);
    DR_CUDA(m, cudaGetLastError());
    m->launches += tc_fwd ? 2 : 2 + 2 * (2 + 2 * T);
    m->last_engine = tc_fwd ? "tcgen05" : "ffma";
    return DR_OK;
}

// the forward recurrence of this micro-batch runs (ran) on the tensor-core kernel: rzn and q are lane-major then
static bool train_tc_forward(const dr_model* m, int bm, int T) {
    return m->cfg.engine != DR_ENGINE_FFMA && dr_tc_supported(m, bm, T) && m->d_wtc != nullptr;
}

// backward of micro-batch mb (G-bar already complete in ws->gbar)
static int train_backward_mb(dr_model* m, dr_train_ws* ws, int b0, int bm) {
    const int F = m->cfg.F, Ml = m->M_loc, M = m->cfg.M, pe = m->off.per_expert, T = ws->T, B = ws->B;
    const float p = m->cfg.dropout_p;
    cudaStream_t st = m->stream;
    const size_t r = (size_t)T * bm;
    const size_t ed_stride = (size_t)Ml * ws->cap_rows;
    size_t td = (size_t)2 * Ml * r * DR_H;
    const bool tc_bwd = m->cfg.engine != DR_ENGINE_FFMA && Ml > 0;
    // dhout [d][e][t] blocks of bm windows x H: row-major for the FFMA chain, window-contiguous for the tensor-core kernel
    if (td) dr_dhout_kernel<<<nblk(td / 4), 256, 0, st>>>(ws->dy, ws->gbar, m->d_ct, ws->mask, ws->seed, p, ws->dhout, Ml, m->e_lo, B, b0, bm, T, td,
                                                           tc_bwd ? 1 : 0);
    if (Ml) {
        int chunk = 2048;
        dim3 grid(Ml, (unsigned)((r + chunk - 1) / chunk));
        dr_head_grad_kernel<<<grid, DR_2H, 0, st>>>(ws->hs, ws->hs + ed_stride * DR_H, ws->S, ws->dy, ws->mask, ws->seed, p, m->d_grad,
                                                     m->off.head_w, m->off.head_b, pe, 1.0f / (float)(M - 1), Ml, m->e_lo, B, b0, bm, T, chunk);
    }
    DR_CUDA(m, cudaGetLastError());
    m->launches += 2;
    const size_t skip = (size_t)bm;                             // dW_hh skips the step whose h_prev is the zero initial state
    const size_t p_dir = (size_t)Ml * 3 * DR_H * F;             // ws->P holds one [Ml][3H][F] block per direction

    if (tc_bwd) {
        // ---------------- tensor-core engine ----------------
        // (1) the reverse-time chain (gate adjoints + dh_{t-1} = dh*z + dgh W_hh) of both directions and all experts is ONE
        //     persistent tcgen05 kernel (csrc/dr_gru_bwd_tc.cu); it leaves g4 = (da_r, da_z, da_n, dq) per row in ws->gi
        const float inv_n = 1.0f / ((float)M * (float)B * (float)T);
        int rc = dr_launch_gru_bwd_tc(m, ws->rzn, ws->q, ws->hs, ws->dhout, ws->gi, (long long)ed_stride, (long long)((size_t)Ml * r), bm, T,
                                      inv_n, train_tc_forward(m, bm, T) ? 1 : 0);
        if (rc) return rc;
        // (2) the reductions over all (t,b) rows as split-fp16 tcgen05 GEMMs (csrc/dr_wgrad_tc.cu), both directions per grid;
        //     the gradient operand is scaled by the same power of two as in (1), h / x by 2^4.  A = columns of g4:
        //     dW_hh uses (da_r, da_z, dq) = column blocks {0, H, 3H},  P = dgi^T x uses (da_r, da_z, da_n) = {0, H, 2H}
        const int ka = dr_grad_scale_log2(inv_n);
        const long long g4e = (long long)T * bm * 4 * DR_H;                             // floats per expert in g4
        const bool p_tc = dr_wgrad_tc_ok(3 * DR_H, F, (int)r);
        const float *Aw[2], *Bw[2], *Ap[2], *Bx[2];
        float *Cw[2], *Cp[2];
        for (int d = 0; d < 2; ++d) {
            const float* g4d = ws->gi + d * ed_stride * 4 * DR_H;
            Aw[d] = g4d + (d ? 0 : skip * 4 * DR_H);                                     // rows t>=1 (fwd) / t<=T-2 (rev)
            Bw[d] = ws->hs + d * ed_stride * DR_H + (d ? skip * DR_H : 0);              // h_{t-1} (fwd) / h_{t+1} (rev)
            Cw[d] = m->d_grad + m->off.w_hh[d];
            Ap[d] = g4d; Bx[d] = ws->xt; Cp[d] = ws->P + d * p_dir;
        }
        const int cols_whh[3] = {0, DR_H, 3 * DR_H}, cols_p[3] = {0, DR_H, 2 * DR_H};
        if (T > 1) {
            rc = dr_launch_wgrad_tc(m, 2, Aw, 4 * DR_H, g4e, cols_whh, Bw, DR_H, (long long)T * bm * DR_H, Cw, DR_H, (long long)pe,
                                    3 * DR_H, DR_H, (int)(r - skip), Ml, ka, 4, 1);
            if (rc) return rc;
        }
        if (p_tc) {
            rc = dr_launch_wgrad_tc(m, 2, Ap, 4 * DR_H, g4e, cols_p, Bx, F, 0, Cp, F, (long long)3 * DR_H * F, 3 * DR_H, F, (int)r, Ml, ka, 4, 0);
            if (rc) return rc;
        }
        for (int d = 0; d < 2; ++d) {
            const float* g4d = ws->gi + d * ed_stride * 4 * DR_H;
            float* Pd = ws->P + d * p_dir;
            if (!p_tc) {                                        // F > 256: fp32 GEMM on the dgi columns of g4 (row stride 4H)
                Gemm gp{g4d, ws->xt, Pd, 3 * DR_H, F, (int)r, 1, 4 * DR_H, F, 1, F, 1, (long)g4e, 0, (long)3 * DR_H * F, 0.0f};
                rc = gemm(m, gp, Ml);
                if (rc) return rc;
            }
            dr_wih_grad_kernel<<<Ml, 128, 0, st>>>(Pd, m->d_blob, m->d_mask, m->d_grad, ws->dmask, m->off.w_ih[d], pe, F);
            int chunk = 1024;
            dim3 grid(Ml, (unsigned)((r + chunk - 1) / chunk));
            dr_colsum4_kernel<<<grid, 4 * DR_H, 0, st>>>(g4d, m->d_grad, m->off.b_ih[d], m->off.b_hh[d], pe, r, chunk);
            DR_CUDA(m, cudaGetLastError());
            m->launches += 2;
        }
        return DR_OK;
    }

    // ---------------- FFMA engine: exact fp32, one gate kernel + one GEMM per time step and direction ----------------
    for (int d = 0; d < 2 && Ml; ++d) {
        float* gi = ws->gi + d * ed_stride * 3 * DR_H;
        float* rzn = ws->rzn + d * ed_stride * 3 * DR_H;
        float* q = ws->q + d * ed_stride * DR_H;
        float* hs = ws->hs + d * ed_stride * DR_H;
        float* dho = ws->dhout + (size_t)d * Ml * r * DR_H;
        DR_CUDA(m, cudaMemsetAsync(ws->dhc, 0, (size_t)Ml * bm * DR_H * sizeof(float), st));
        for (int s = T - 1; s >= 0; --s) {                      // reverse of the forward processing order
            const int t = d ? (T - 1 - s) : s, tp = d ? t + 1 : t - 1;
            const float* hprev = (s == 0) ? nullptr : hs + (size_t)tp * bm * DR_H;
            size_t tg = (size_t)Ml * bm * DR_H;
            dr_gate_bwd_kernel<<<nblk(tg), 256, 0, st>>>(rzn, gi, q, hprev, dho, ws->dhc, t, bm, T, tg);
            if (s > 0) {   // dh_{prev} = dh*z + dgh W_hh
                Gemm gd{rzn + (size_t)t * bm * 3 * DR_H, m->d_blob + m->off.w_hh[d], ws->dhc, bm, DR_H, 3 * DR_H,
                        3 * DR_H, 1, DR_H, 1, DR_H, 1, (long)T * bm * 3 * DR_H, (long)pe, (long)bm * DR_H, 1.0f};
                int rc = gemm(m, gd, Ml);
                if (rc) return rc;
            }
        }
        // weight gradients as GEMMs over all (t,b) rows of the micro-batch
        if (T > 1) {
            const float* A = rzn + (d ? 0 : skip * 3 * DR_H);     // dgh rows for t>=1 (fwd) / t<=T-2 (rev)
            const float* Bp = hs + (d ? skip * DR_H : 0);         // h_{t-1} (fwd) / h_{t+1} (rev)
            Gemm gw{A, Bp, m->d_grad + m->off.w_hh[d], 3 * DR_H, DR_H, (int)(r - skip),
                    1, 3 * DR_H, DR_H, 1, DR_H, 1, (long)T * bm * 3 * DR_H, (long)T * bm * DR_H, (long)pe, 1.0f};
            int rc = gemm(m, gw, Ml);
            if (rc) return rc;
        }
        Gemm gp{gi, ws->xt, ws->P, 3 * DR_H, F, (int)r, 1, 3 * DR_H, F, 1, F, 1,
                (long)T * bm * 3 * DR_H, 0, (long)3 * DR_H * F, 0.0f};
        int rc = gemm(m, gp, Ml);
        if (rc) return rc;
        dr_wih_grad_kernel<<<Ml, 128, 0, st>>>(ws->P, m->d_blob, m->d_mask, m->d_grad, ws->dmask, m->off.w_ih[d], pe, F);
        int chunk = 1024;
        dim3 grid(Ml, (unsigned)((r + chunk - 1) / chunk));
        dr_colsum_kernel<<<grid, 3 * DR_H, 0, st>>>(rzn, m->d_grad, m->off.b_hh[d], pe, r, chunk);
        dr_colsum_kernel<<<grid, 3 * DR_H, 0, st>>>(gi, m->d_grad, m->off.b_ih[d], pe, r, chunk);
        DR_CUDA(m, cudaGetLastError());
        m->launches += 3 + 2 * T;
    }
    return DR_OK;
}

// kind: 0 = step finished, 1 = all-reduce(sum) `count` elements at `ptr` (dtype 0 = fp32, 1 = fp64) across the ranks, then call again
static int train_advance_inner(dr_model* m, int* kind, void** ptr, long long* count, int* dtype);

// any failure inside a step (a CUDA error after a launch, an allocation) abandons the step: the handle goes back to
// TS_IDLE so that the next dr_train_begin is accepted instead of reporting "a training step is already in flight"
int dr_train_advance_impl(dr_model* m, int* kind, void** ptr, long long* count, int* dtype) {
    int rc = train_advance_inner(m, kind, ptr, count, dtype);
    if (rc != DR_OK) {
        dr_train_ws* ws = reinterpret_cast<dr_train_ws*>(m->train_ws);
        if (ws) ws->stage = TS_IDLE;
    }
    return rc;
}

static int train_advance_inner(dr_model* m, int* kind, void** ptr, long long* count, int* dtype) {
    dr_train_ws* ws = reinterpret_cast<dr_train_ws*>(m->train_ws);
    if (!ws || ws->stage == TS_IDLE) return dr_fail(m, DR_ESTATE, "dr_train_advance without dr_train_begin");
    const int Ml = m->M_loc, M = m->cfg.M, T = ws->T, B = ws->B, F = m->cfg.F, pe = m->off.per_expert;
    const float p = m->cfg.dropout_p;
    const bool sharded = m->cfg.world > 1;
    cudaStream_t st = m->stream;
    const float inv_n = 1.0f / ((float)M * (float)B * (float)T);
    double* acc = reinterpret_cast<double*>(m->d_loss);
    *kind = 0; *ptr = nullptr; *count = 0; *dtype = 0;
    for (;;) {
        const int b0 = ws->mb * ws->Bm, bm = std::min(ws->Bm, B - b0);
        const size_t r = (size_t)T * bm;
        const size_t ed_stride = (size_t)Ml * ws->cap_rows;
        switch (ws->stage) {
        case TS_MB_BEGIN: {
            // pass 0 runs every forward (the loss needs all windows); pass 1 recomputes a micro-batch's activations
            // only when there is more than one micro-batch (otherwise they are still in the workspace)
            const bool need_fwd = (ws->pass == 0) || (ws->n_mb > 1);
            ws->stage = TS_AFTER_S;
            if (need_fwd) {
                int rc = train_forward_mb(m, ws, b0, bm);
                if (rc) { ws->stage = TS_IDLE; return rc; }
                if (sharded) { *kind = 1; *ptr = ws->S; *count = (long long)(r * DR_2H); *dtype = 0; return DR_OK; }
            }
            break;
        }
        case TS_AFTER_S: {
            if (ws->pass == 0) {
                size_t tw = r * Ml * 32;
                if (tw) dr_head_fwd_kernel<<<nblk(tw), 256, 0, st>>>(ws->hs, ws->hs + ed_stride * DR_H, ws->S, ws->mask, ws->seed, p, m->d_ct,
                                                                      m->d_abar, m->d_hb, ws->out_dev, Ml, m->e_lo, B, b0, bm, T);
                m->launches += 1;
                if (ws->mb + 1 < ws->n_mb) { ws->mb += 1; ws->stage = TS_MB_BEGIN; break; }
                size_t n_rm = (size_t)B * T * Ml;              // local metrics; the mean runs over the GLOBAL M*B*T (inv_n)
                unsigned blocks = std::max(1u, std::min<unsigned>(nblk(n_rm), 148u * 8u));
                dr_loss_grad_kernel<<<blocks, 256, 0, st>>>(ws->out_dev, ws->y, ws->dy, n_rm, m->cfg.quantiles[0], m->cfg.quantiles[1],
                                                             m->cfg.quantiles[2], inv_n, acc);
                DR_CUDA(m, cudaGetLastError());
                m->launches += 1;
                ws->stage = TS_AFTER_LOSS;
                if (sharded) { *kind = 1; *ptr = acc; *count = 1; *dtype = 1; return DR_OK; }
            } else {
                size_t ts = r * DR_2H;
                dr_gbar_kernel<<<nblk(ts), 256, 0, st>>>(ws->dy, m->d_abar, ws->gbar, Ml, b0, bm, T, ts);
                DR_CUDA(m, cudaGetLastError());
                m->launches += 1;
                ws->stage = TS_AFTER_G;
                if (sharded) { *kind = 1; *ptr = ws->gbar; *count = (long long)ts; *dtype = 0; return DR_OK; }
            }
            break;
        }
        case TS_AFTER_LOSS:
            dr_finish_loss_kernel<<<1, 1, 0, st>>>(acc, (double)inv_n, ws->loss_dev);
            m->launches += 1;
            ws->pass = 1; ws->mb = 0; ws->stage = TS_MB_BEGIN;
            break;
        case TS_AFTER_G: {
            int rc = train_backward_mb(m, ws, b0, bm);
            if (rc) { ws->stage = TS_IDLE; return rc; }
            if (ws->mb + 1 < ws->n_mb) { ws->mb += 1; ws->stage = TS_MB_BEGIN; } else ws->stage = TS_FINISH;
            break;
        }
        // ---------------- bf16 engine: one pass per micro-batch ----------------
        case T16_MB_BEGIN: {
            int rc = dr_t16_pack_x(m, ws->x + (size_t)b0 * T * F, bm, T, ws->x16);
            if (rc) return rc;
            rc = dr_launch_gru_tc16(m, ws->w16, ws->x16, bm, T, ws->S16, ws->P16, ws->gate16, ws->h16, ws->mask, ws->seed, b0, B);
            if (rc) return rc;
            m->last_engine = "tcgen05-bf16";
            ws->stage = T16_AFTER_S;
            if (sharded) { *kind = 1; *ptr = ws->S16; *count = (long long)dr_s_floats(bm, T); *dtype = 0; return DR_OK; }
            break;
        }
        case T16_AFTER_S: {
            // heads: the inference head kernel on the (complete) S and this micro-batch's own-expert partials; the optional
            // inference output transform is not part of the training graph
            float* out_mb = ws->out_dev + (size_t)b0 * T * Ml * DR_Q;
            float* keep_p = m->d_p; const bool keep_dn = m->dn_on;
            m->d_p = ws->P16; m->dn_on = false;
            int rc = dr_launch_heads_tc(m, ws->S16, bm, T, out_mb);
            m->d_p = keep_p; m->dn_on = keep_dn;
            if (rc) return rc;
            const size_t n_rm = (size_t)bm * T * Ml;
            if (n_rm) {
                unsigned blocks = std::max(1u, std::min<unsigned>(nblk(n_rm), 148u * 8u));
                dr_loss_grad_kernel<<<blocks, 256, 0, st>>>(out_mb, ws->y + (size_t)b0 * T * Ml, ws->dy16, n_rm, m->cfg.quantiles[0],
                                                             m->cfg.quantiles[1], m->cfg.quantiles[2], inv_n, acc, bm, T, Ml);   // dy time-major
                DR_CUDA(m, cudaGetLastError());
                m->launches += 1;
            }
            // G-bar[(t,b)][k] = sum_{e,q} Abar[e][q][k] dy[(t,b),e,q]: one fp32 GEMM [T*bm x 3M_loc] x [3M_loc x 2H] (rows time-major)
            if (Ml) {
                Gemm gg{ws->dy16, m->d_abar, ws->gbar16, (int)((size_t)bm * T), DR_2H, Ml * DR_Q,
                        (long)Ml * DR_Q, 1, DR_2H, 1, DR_2H, 1, 0, 0, 0, 0.0f};
                rc = gemm(m, gg, 1);
                if (rc) return rc;
            } else {
                DR_CUDA(m, cudaMemsetAsync(ws->gbar16, 0, (size_t)bm * T * DR_2H * sizeof(float), st));
            }
            ws->stage = T16_AFTER_G;
            if (sharded) { *kind = 1; *ptr = ws->gbar16; *count = (long long)((size_t)bm * T * DR_2H); *dtype = 0; return DR_OK; }
            break;
        }
        case T16_AFTER_G: {
            int rc = dr_launch_gru_bwd16(m, ws->whT16, ws->gate16, ws->h16, ws->dy16, ws->gbar16, bm, T, ws->mask, ws->seed, b0, B);
            if (rc) return rc;
            rc = dr_launch_wgrad16(m, ws->gate16, ws->h16, ws->x16, ws->zero16, ws->Px16, bm, T);
            if (rc) return rc;
            if (Ml) {
                const int chunk = std::max(1, 2048 / std::max(bm, 1));              // time steps per CTA (~2048 rows)
                dim3 grid(Ml, (unsigned)((T + chunk - 1) / chunk));
                drt16::Drop dp;
                dp.mask = ws->mask; dp.seed = ws->seed; dp.inv_keep = 1.0f / (1.0f - p); dp.thr16 = (uint32_t)(p * 65536.0f + 0.5f);
                dr_head_grad16_kernel<<<grid, 256, 0, st>>>(ws->h16, ws->S16, ws->dy16, dp, m->d_grad, m->off.head_w, m->off.head_b, pe,
                                                            1.0f / (float)(M - 1), Ml, m->e_lo, B, b0, bm, T, chunk);
                DR_CUDA(m, cudaGetLastError());
                m->launches += 1;
            }
            if (ws->mb + 1 < ws->n_mb) { ws->mb += 1; ws->stage = T16_MB_BEGIN; break; }
            ws->stage = T16_AFTER_LOSS;
            if (sharded) { *kind = 1; *ptr = acc; *count = 1; *dtype = 1; return DR_OK; }
            break;
        }
        case T16_AFTER_LOSS: {
            dr_finish_loss_kernel<<<1, 1, 0, st>>>(acc, (double)inv_n, ws->loss_dev);
            for (int d = 0; d < 2 && Ml; ++d)
                dr_wih_grad_kernel<<<Ml, 128, 0, st>>>(ws->Px16 + (size_t)d * Ml * 3 * DR_H * F, m->d_blob, m->d_mask, m->d_grad, ws->dmask,
                                                       m->off.w_ih[d], pe, F);
            DR_CUDA(m, cudaGetLastError());
            m->launches += 3;
            ws->stage = TS_FINISH;
            break;
        }
        case TS_FINISH: {
            size_t nblob = (size_t)Ml * pe;
            if (Ml) dr_mask_bwd_kernel<<<Ml, DR_H, F * sizeof(float), st>>>(m->d_blob, m->off, F, m->d_mask, ws->dmask, m->d_grad);
            m->adam_step += 1;                      // torch.optim.Adam defaults (estimate.py:61)
            const double b1 = 0.9, b2 = 0.999;
            double bc1 = 1.0 - pow(b1, (double)m->adam_step), bc2 = 1.0 - pow(b2, (double)m->adam_step);
            if (nblob) dr_adam_kernel<<<nblk(nblob), 256, 0, st>>>(m->d_blob, m->d_grad, m->d_adam_m, m->d_adam_v, nblob,
                                                                   (float)(ws->lr / bc1), (float)(1.0 / sqrt(bc2)), (float)b1, (float)b2, 1e-8f);
            DR_CUDA(m, cudaGetLastError());
            m->launches += 2;
            ws->stage = TS_IDLE;
            int rc = dr_launch_prep(m);             // the inference images follow the new weights
            if (rc) return rc;
            return dr_tc_prep_weights(m);
        }
        default:
            ws->stage = TS_IDLE;
            return dr_fail(m, DR_ESTATE, "corrupt training state");
        }
    }
}

int dr_train_step_impl(dr_model* m, const float* x, const float* y, int B, int T, const uint8_t* mask, uint64_t seed,
                       float lr, float* loss_dev, float* out_dev) {
    if (m->cfg.world != 1)
        return dr_fail(m, DR_ESTATE, "dr_train_step needs world == 1; sharded handles drive dr_train_begin_dev / dr_train_advance");
    int rc = dr_train_begin_impl(m, x, y, B, T, mask, seed, lr, loss_dev, out_dev);
    if (rc) return rc;
    int kind; void* ptr; long long count; int dtype;
    return dr_train_advance_impl(m, &kind, &ptr, &count, &dtype);
}

extern "C" {

int dr_train_step_dev(dr_model* m, const float* x_dev, const float* y_dev, int32_t B, int32_t T,
                      const uint8_t* dropout_mask_dev, uint64_t seed, float lr, float* loss_dev, float* out_dev) {
    if (!m) return DR_EINVAL;
    if (!x_dev || !y_dev || !loss_dev || !out_dev || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_train_step_dev: bad argument");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "train step before dr_load_weights");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return dr_train_step_impl(m, x_dev, y_dev, B, T, dropout_mask_dev, seed, lr, loss_dev, out_dev);
}

int dr_train_step(dr_model* m, const float* x, const float* y, int32_t B, int32_t T,
                  const uint8_t* dropout_mask, uint64_t seed, float lr, float* loss_out) {
    if (!m) return DR_EINVAL;
    if (!x || !y || !loss_out || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_train_step: bad argument");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "train step before dr_load_weights");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    size_t nx = (size_t)B * T * m->cfg.F, ny = (size_t)B * T * m->cfg.M, no = ny * DR_Q, nm = (size_t)m->cfg.M * B * T * DR_2H;
    int rc;
    if ((rc = dr_reserve(m, (void**)&m->d_xin, &m->xin_cap, nx * sizeof(float)))) return rc;
    if ((rc = dr_reserve(m, (void**)&m->d_y, &m->y_cap, ny * sizeof(float)))) return rc;
    if ((rc = dr_reserve(m, (void**)&m->d_out, &m->out_cap, no * sizeof(float)))) return rc;
    DR_CUDA(m, cudaMemcpyAsync(m->d_xin, x, nx * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    DR_CUDA(m, cudaMemcpyAsync(m->d_y, y, ny * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    uint8_t* dmask = nullptr;
    if (dropout_mask) {
        if ((rc = dr_reserve(m, (void**)&m->d_dropmask, &m->dropmask_cap, nm))) return rc;
        dmask = reinterpret_cast<uint8_t*>(m->d_dropmask);
        DR_CUDA(m, cudaMemcpyAsync(dmask, dropout_mask, nm, cudaMemcpyHostToDevice, m->stream));
    }
    rc = dr_train_step_impl(m, m->d_xin, m->d_y, B, T, dmask, seed, lr, m->d_loss + 4, m->d_out);
    if (rc) return rc;
    DR_CUDA(m, cudaMemcpyAsync(loss_out, m->d_loss + 4, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    return DR_OK;
}

int dr_train_begin_dev(dr_model* m, const float* x_dev, const float* y_dev, int32_t B, int32_t T,
                       const uint8_t* dropout_mask_dev, uint64_t seed, float lr, float* loss_dev, float* out_dev) {
    if (!m) return DR_EINVAL;
    if (!x_dev || !y_dev || !loss_dev || !out_dev || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_train_begin_dev: bad argument");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "train step before dr_load_weights");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return dr_train_begin_impl(m, x_dev, y_dev, B, T, dropout_mask_dev, seed, lr, loss_dev, out_dev);
}

int dr_train_advance(dr_model* m, int32_t* kind, void** ptr, int64_t* count, int32_t* dtype) {
    if (!m || !kind || !ptr || !count || !dtype) return DR_EINVAL;
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    int k = 0, dt = 0; void* p = nullptr; long long c = 0;
    int rc = dr_train_advance_impl(m, &k, &p, &c, &dt);
    *kind = k; *ptr = p; *count = c; *dtype = dt;
    return rc;
}

int dr_train_set_microbatch(dr_model* m, int32_t windows) {
    if (!m || windows < 0) return DR_EINVAL;
    m->train_mb = windows;
    return DR_OK;
}

int dr_get_grads(dr_model* m, float* host_blob, size_t n) {
    if (!m) return DR_EINVAL;
    size_t pe = (size_t)m->off.per_expert;
    if (!host_blob || n != pe * m->cfg.M) return dr_fail(m, DR_EINVAL, "dr_get_grads: wrong blob size");
    if (!m->d_grad) return dr_fail(m, DR_ESTATE, "dr_get_grads before any train step");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    DR_CUDA(m, cudaMemcpyAsync(host_blob + (size_t)m->e_lo * pe, m->d_grad, (size_t)m->M_loc * pe * sizeof(float),
                               cudaMemcpyDeviceToHost, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    return DR_OK;
}

}  // extern "C"

void dr_train_free(dr_model* m) {
    dr_train_ws* ws = reinterpret_cast<dr_train_ws*>(m->train_ws);
    if (!ws) return;
    float* ptrs[] = {ws->xt, ws->gi, ws->rzn, ws->q, ws->hs, ws->dhout, ws->gh, ws->dhc, ws->S, ws->gbar, ws->dy, ws->P, ws->dmask};
    for (float* p : ptrs) if (p) cudaFree(p);
    void* p16[] = {ws->w16, ws->whT16, ws->x16, ws->gate16, ws->h16, ws->zero16, ws->S16, ws->P16, ws->dy16, ws->gbar16, ws->Px16};
    for (void* p : p16) if (p) cudaFree(p);
    delete ws;
    m->train_ws = nullptr;
}
