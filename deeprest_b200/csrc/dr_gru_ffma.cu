// K1 (fp32 engine) — fused bidirectional-GRU recurrence on the CUDA cores.
//
// Replaces, for every local expert and both directions at once, the reference's per-expert
//   mask -> x*mask -> permute -> nn.GRU -> permute            (qrnn.py:33-42)
// plus the data movement of the cross-expert mean/concat/head  (qrnn.py:46-54), which is folded
// algebraically:  every chain adds its hidden state into the cross-expert
// sum S[b,t,:] and adds its own-expert head term (C_i - A_i/(M-1))·h into out_local[b,t,i,:].
// rnn_out [M,B,T,2H] (38.6 GB at config 2) is never materialised.
//
// One CTA = one chain group: (batch tile of BT windows, direction, expert), 256 threads.
// Per time step the CTA evaluates the [BT x (Fp+H)] x [(Fp+H) x 3H] gate GEMM with fp32 FFMA:
//   A (x_t tile and h_{t-1}) lives in shared memory k-major, the 288 KB weight image streams
//   from L2 through a cp.async ring (it does not fit beside A), each thread owns an
//   RPT x 4 register tile of hidden units for the four accumulators (r, z, gi_n, gh_n).
// This is the exact-fp32 engine (parity reference for the tcgen05 engine and the small-batch path).
#include "dr_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kChunkFloats = DR_KC * 3 * 64;     // one K-chunk of one pass: 16 x (r,z,n) x 64

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

template <int RPT> struct Cfg {
    static constexpr int BT = 16 * RPT;
    static constexpr int NSTAGE = (RPT >= 8) ? 2 : 4;   // small tiles are latency bound: deeper ring
};

template <int RPT>
__device__ __forceinline__ void load_rows(const float* p, float (&a)[RPT]) {
    if constexpr (RPT == 1) { a[0] = p[0]; }
    else if constexpr (RPT == 2) { float2 v = *reinterpret_cast<const float2*>(p); a[0] = v.x; a[1] = v.y; }
    else {
#pragma unroll
        for (int i = 0; i < RPT; i += 4) {
            float4 v = *reinterpret_cast<const float4*>(p + i);
            a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
        }
    }
}

template <int RPT>
__global__ void __launch_bounds__(kThreads, 1)
dr_gru_ffma_kernel(const float* __restrict__ xT,     // [T][Fp][Bp]
                   const float* __restrict__ wf,     // [M_loc][2][2][KT][3][64]
                   const float* __restrict__ bias4,  // [M_loc][2][4][H]
                   const float* __restrict__ ct,     // [M_loc][2][Q][H]
                   float* __restrict__ S,            // [T][2H/4][BpS][4]  (k-group major: coalesced REDs)
                   float* __restrict__ out_local,    // [B][T][M_loc][Q]
                   int B, int T, int Fp, int Bp, int BpS, int M_loc) {
    constexpr int BT = Cfg<RPT>::BT;
    constexpr int NS = Cfg<RPT>::NSTAGE;
    extern __shared__ __align__(16) float smem[];
    const int KT = Fp + DR_H;
    const int NX = Fp / DR_KC;               // chunks that read the x tile
    const int NCH = KT / DR_KC;              // chunks per pass
    const int per_step = 2 * NCH;

    float* xs = smem;                         // [Fp][BT]
    float* hs = xs + (size_t)Fp * BT;         // [2][H][BT]
    float* ws = hs + 2 * DR_H * BT;           // [NS][16][3][64]
    float* bs = ws + NS * kChunkFloats;       // [4][H]
    float* cs = bs + 4 * DR_H;                // [Q][H]

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int r0 = ty * RPT;
    const int b0 = blockIdx.x * BT;
    const int dir = blockIdx.y;
    const int e = blockIdx.z;

    const float* wbase = wf + (size_t)(e * 2 + dir) * 2 * KT * 192;
    const long long total_chunks = (long long)T * per_step;

    auto issue_w = [&](long long idx) {       // chunk idx (global over all steps) -> ring slot
        if (idx < total_chunks) {
            int g = (int)(idx % per_step);    // g = p*NCH + c ; the image is laid out [p][k][3][64]
            const float* src = wbase + (size_t)g * kChunkFloats;
            float* dst = ws + (int)(idx % NS) * kChunkFloats;
            for (int i = tid; i < kChunkFloats / 4; i += kThreads) cp_async16(dst + i * 4, src + i * 4);
        }
    };
    auto issue_x = [&](int s) {               // x tile of step s -> xs
        int tt = dir ? (T - 1 - s) : s;
        const float* src = xT + ((size_t)tt * Fp) * Bp + b0;
        constexpr int V = BT / 4;             // float4 per feature row
        for (int i = tid; i < Fp * V; i += kThreads) {
            int f = i / V, v = i % V;
            cp_async16(xs + f * BT + v * 4, src + (size_t)f * Bp + v * 4);
        }
    };

    // ---- prologue ----
    for (int i = tid; i < 4 * DR_H; i += kThreads) bs[i] = bias4[(size_t)(e * 2 + dir) * 4 * DR_H + i];
    for (int i = tid; i < DR_Q * DR_H; i += kThreads) cs[i] = ct[(size_t)(e * 2 + dir) * DR_Q * DR_H + i];
    for (int i = tid; i < DR_H * BT; i += kThreads) hs[i] = 0.0f;      // h0 = 0, qrnn.py:39
    issue_x(0);
    for (int st = 0; st < NS - 1; ++st) { issue_w(st); cp_async_commit(); }

    long long consume = 0;
    for (int s = 0; s < T; ++s) {
        const int tt = dir ? (T - 1 - s) : s;
        const float* hcur = hs + (s & 1) * DR_H * BT;
        float* hnxt = hs + ((s & 1) ^ 1) * DR_H * BT;

#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
            float acc_r[RPT][4], acc_z[RPT][4], acc_i[RPT][4], acc_h[RPT][4];
#pragma unroll
            for (int i = 0; i < RPT; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc_r[i][j] = 0.f; acc_z[i][j] = 0.f; acc_i[i][j] = 0.f; acc_h[i][j] = 0.f; }

#pragma unroll 1
            for (int c = 0; c < NCH; ++c, ++consume) {
                cp_async_wait<NS - 2>();       // chunk `consume` (and anything older) has landed
                __syncthreads();               // ... for every thread; also: ring slot (consume-1) is free
                if (p == 1 && c == NX && s + 1 < T) issue_x(s + 1);   // xs is dead for this step now
                issue_w(consume + NS - 1);
                cp_async_commit();

                const float* wrow = ws + (int)(consume % NS) * kChunkFloats + tx * 4;
                if (c < NX) {
                    const float* arow = xs + (size_t)(c * DR_KC) * BT + r0;
#pragma unroll
                    for (int kk = 0; kk < DR_KC; ++kk) {
                        float a[RPT];
                        load_rows<RPT>(arow + kk * BT, a);
                        float4 wr = *reinterpret_cast<const float4*>(wrow + kk * 192);
                        float4 wz = *reinterpret_cast<const float4*>(wrow + kk * 192 + 64);
                        float4 wn = *reinterpret_cast<const float4*>(wrow + kk * 192 + 128);
                        const float r4[4] = {wr.x, wr.y, wr.z, wr.w};
                        const float z4[4] = {wz.x, wz.y, wz.z, wz.w};
                        const float n4[4] = {wn.x, wn.y, wn.z, wn.w};
#pragma unroll
                        for (int i = 0; i < RPT; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc_r[i][j] = fmaf(a[i], r4[j], acc_r[i][j]);
                                acc_z[i][j] = fmaf(a[i], z4[j], acc_z[i][j]);
                                acc_i[i][j] = fmaf(a[i], n4[j], acc_i[i][j]);
                            }
                    }
                } else {
                    const float* arow = hcur + (size_t)(c * DR_KC - Fp) * BT + r0;
#pragma unroll
                    for (int kk = 0; kk < DR_KC; ++kk) {
                        float a[RPT];
                        load_rows<RPT>(arow + kk * BT, a);
                        float4 wr = *reinterpret_cast<const float4*>(wrow + kk * 192);
                        float4 wz = *reinterpret_cast<const float4*>(wrow + kk * 192 + 64);
                        float4 wn = *reinterpret_cast<const float4*>(wrow + kk * 192 + 128);
                        const float r4[4] = {wr.x, wr.y, wr.z, wr.w};
                        const float z4[4] = {wz.x, wz.y, wz.z, wz.w};
                        const float n4[4] = {wn.x, wn.y, wn.z, wn.w};
#pragma unroll
                        for (int i = 0; i < RPT; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc_r[i][j] = fmaf(a[i], r4[j], acc_r[i][j]);
                                acc_z[i][j] = fmaf(a[i], z4[j], acc_z[i][j]);
                                acc_h[i][j] = fmaf(a[i], n4[j], acc_h[i][j]);
                            }
                    }
                }
            }

            // ---- gate epilogue for hidden units p*64 + tx*4 .. +3 (GRU equations) ----
            float hn[RPT][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int hid = p * 64 + tx * 4 + j;
                const float br = bs[hid], bz = bs[DR_H + hid], bin = bs[2 * DR_H + hid], bhn = bs[3 * DR_H + hid];
#pragma unroll
                for (int i = 0; i < RPT; ++i) {
                    float r = dr_sigmoid(acc_r[i][j] + br);
                    float z = dr_sigmoid(acc_z[i][j] + bz);
                    float n = dr_tanh(acc_i[i][j] + bin + r * (acc_h[i][j] + bhn));
                    float hold = hcur[hid * BT + r0 + i];
                    // h' = (1-z)*n + z*h, evaluated as torch's CPU cell does: (h - n)*z + n
                    float hnew = __fadd_rn(__fmul_rn(__fsub_rn(hold, n), z), n);
                    hnxt[hid * BT + r0 + i] = hnew;
                    hn[i][j] = hnew;
                }
            }
            // cross-expert sum S[b,t, dir*H + hid] += h   (the mean of qrnn.py:46-52, deferred)
#pragma unroll
            for (int i = 0; i < RPT; ++i) {
                int b = b0 + r0 + i;
                if (b < B)
                    dr_red_add_v4(S + (((size_t)tt * 64 + dir * 32 + p * 16 + tx) * BpS + b) * 4,
                                  hn[i][0], hn[i][1], hn[i][2], hn[i][3]);
            }
        }
        __syncthreads();                        // h_t complete in hnxt

        // ---- own-expert head term: out_local[b,t,e,q] += (C - A/(M-1))[q, dir half] · h_t ----
        {
            constexpr int PARTS = kThreads / BT;
            constexpr int KPER = DR_H / PARTS;
            const int row = tid % BT, part = tid / BT;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll 8
            for (int k = part * KPER; k < (part + 1) * KPER; ++k) {
                float hv = hnxt[k * BT + row];
                a0 = fmaf(cs[k], hv, a0);
                a1 = fmaf(cs[DR_H + k], hv, a1);
                a2 = fmaf(cs[2 * DR_H + k], hv, a2);
            }
            int b = b0 + row;
            if (b < B) {
                float* o = out_local + (((size_t)b * T + tt) * M_loc + e) * DR_Q;
                dr_red_add(o, a0); dr_red_add(o + 1, a1); dr_red_add(o + 2, a2);
            }
        }
    }
    cp_async_wait<0>();
}

// x [B,T,F] -> xT [T][Fp][Bp], zero padded in F and B
__global__ void dr_xT_kernel(const float* __restrict__ x, float* __restrict__ xT,
                             int B, int T, int F, int Fp, int Bp, long long xbs /* floats between window starts */) {
    __shared__ float tile[32][33];
    // grid: (ceil(Bp/32), ceil(Fp/32), T)
    int t = blockIdx.z;
    int bb = blockIdx.x * 32, ff = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int b = bb + i, f = ff + threadIdx.x;
        tile[i][threadIdx.x] = (b < B && f < F) ? x[(size_t)b * xbs + (size_t)t * F + f] : 0.0f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        int f = ff + i, b = bb + threadIdx.x;
        if (f < Fp && b < Bp) xT[((size_t)t * Fp + f) * Bp + b] = tile[threadIdx.x][i];
    }
}

template <int RPT>
int launch_one(dr_model* m, int B, int T, int Bp, float* S, float* out_local) {
    const int BpS = dr_s_rows(B);
    constexpr int BT = Cfg<RPT>::BT;
    int Fp = m->Fp;
    size_t smem = ((size_t)Fp * BT + 2 * DR_H * BT + Cfg<RPT>::NSTAGE * kChunkFloats + 4 * DR_H + DR_Q * DR_H) * sizeof(float);
    if (smem > 227 * 1024) return dr_fail(m, DR_EUNSUPPORTED, "FFMA engine: F too large for the shared-memory tile");
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_ffma_kernel<RPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(Bp / BT, 2, m->M_loc);
    dr_gru_ffma_kernel<RPT><<<grid, kThreads, smem, m->stream>>>(m->d_xT, m->d_wf, m->d_bias4, m->d_ct, S, out_local,
                                                                 B, T, Fp, Bp, BpS, m->M_loc);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

}  // namespace

int dr_ffma_rows_per_thread(int B) { return B > 64 ? 8 : B > 32 ? 4 : B > 16 ? 2 : 1; }

int dr_launch_xT(dr_model* m, const float* x_dev, int B, int T, int Bp) {
    dim3 grid((Bp + 31) / 32, (m->Fp + 31) / 32, T), block(32, 8);
    dr_xT_kernel<<<grid, block, 0, m->stream>>>(x_dev, m->d_xT, B, T, m->cfg.F, m->Fp, Bp,
                                                m->x_bstride ? m->x_bstride : (long long)T * m->cfg.F);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_gru_ffma(dr_model* m, int B, int T, int Bp, float* S, float* out_local) {
    switch (dr_ffma_rows_per_thread(B)) {
        case 8: return launch_one<8>(m, B, T, Bp, S, out_local);
        case 4: return launch_one<4>(m, B, T, Bp, S, out_local);
        case 2: return launch_one<2>(m, B, T, Bp, S, out_local);
        default: return launch_one<1>(m, B, T, Bp, S, out_local);
    }
}
