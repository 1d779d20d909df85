// K2 — cross-expert-mean head term, output layout, and the pinball loss.
//
//  dr_head_kernel      out_local[r, i*Q+q] += (A_i/(M-1))[q,:] · S[r,:] + b_i[q]     (qrnn.py:46-54, folded; FFMA engine —
//                      the tcgen05 engine's K2 is dr_head_tc.cu)
//  dr_interleave_kernel gathered[world][R][M_loc*Q] -> out[R][M*Q]                   (qrnn.py:55 layout)
//  dr_loss_*           pinball loss                                                  (qrnn.py:58-67)
#include "dr_common.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 32, LD = 68;

// grid (ceil(B/64), T, ceil(N/64)): a tile is 64 windows of one time step, so the k-group-major S
// ([T][64][Bp][4]) is read as contiguous 1 KB runs.
__global__ void __launch_bounds__(256)
dr_head_kernel(const float* __restrict__ S, const float* __restrict__ abar, const float* __restrict__ hb,
               float* __restrict__ out, int B, int T, int BpS, int N,
               const float* __restrict__ dn_scale, const float* __restrict__ dn_offset, float clamp_min) {
    __shared__ __align__(16) float As[TK][LD];
    __shared__ __align__(16) float Bs[TK][LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b0 = blockIdx.x * TM;
    const int t = blockIdx.y;
    const int c0 = blockIdx.z * TN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < DR_2H; k0 += TK) {
#pragma unroll
        for (int i = 0; i < (TM * TK / 4) / 256; ++i) {       // float4 loads: 64 rows x 8 k-groups
            int idx = tid + i * 256, r = idx % TM, kg = idx / TM;
            int b = b0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (b < B) v = *reinterpret_cast<const float4*>(S + (((size_t)t * 64 + k0 / 4 + kg) * BpS + b) * 4);
            As[kg * 4 + 0][r] = v.x; As[kg * 4 + 1][r] = v.y; As[kg * 4 + 2][r] = v.z; As[kg * 4 + 3][r] = v.w;
        }
#pragma unroll
        for (int i = 0; i < (TN * TK) / 256; ++i) {
            int idx = tid + i * 256, k = idx % TK, r = idx / TK;
            int col = c0 + r;
            Bs[k][r] = (col < N) ? abar[(size_t)col * DR_2H + k0 + k] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int b = b0 + ty * 4 + i;
        if (b >= B) continue;
        float* orow = out + ((size_t)b * T + t) * N;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int col = c0 + tx * 4 + j;
            if (col < N) {
                const float own = orow[col];          // the FFMA recurrence kernel REDs its own-expert term in place
                float v = own + acc[i][j] + hb[col];
                if (dn_scale) {       // estimate.py:96,102 — clamp the normalised forecast, then undo the min-max scaling
                    int e = col / DR_Q;
                    v = fmaxf(v, clamp_min) * dn_scale[e] + dn_offset[e];
                }
                orow[col] = v;
            }
        }
    }
}

__global__ void dr_interleave_kernel(const float* __restrict__ g, float* __restrict__ out,
                                     size_t R, int NL, int world) {
    size_t total = R * (size_t)NL * world;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        int c = (int)(i % NL); size_t r = i / NL;
        int w = (int)(r % world); size_t row = r / world;
        out[(row * world + w) * NL + c] = g[((size_t)w * R + row) * NL + c];
    }
}

__global__ void dr_loss_zero_kernel(double* acc) { *acc = 0.0; }

__global__ void dr_loss_kernel(const float* __restrict__ out, const float* __restrict__ y, size_t n_rm,
                               float q0, float q1, float q2, double* acc) {
    float local = 0.0f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_rm; i += (size_t)gridDim.x * blockDim.x) {
        float yy = y[i];
        const float* o = out + i * DR_Q;
        float e0 = yy - o[0], e1 = yy - o[1], e2 = yy - o[2];
        // rho_q(e) = max((q-1)e, q e)   qrnn.py:64
        local += fmaxf((q0 - 1.0f) * e0, q0 * e0) + fmaxf((q1 - 1.0f) * e1, q1 * e1) + fmaxf((q2 - 1.0f) * e2, q2 * e2);
    }
    __shared__ double red[256];
    red[threadIdx.x] = (double)local;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(acc, red[0]);
}

__global__ void dr_loss_final_kernel(const double* acc, double inv_n, float* loss) { *loss = (float)(*acc * inv_n); }

}  // namespace

int dr_launch_heads(dr_model* m, const float* S, int B, int T, float* out_local) {
    int N = m->M_loc * DR_Q;
    if (N == 0) return DR_OK;
    dim3 grid((B + TM - 1) / TM, T, (N + TN - 1) / TN);
    dr_head_kernel<<<grid, 256, 0, m->stream>>>(S, m->d_abar, m->d_hb, out_local, B, T, dr_s_rows(B), N,
                                                m->dn_on ? m->d_dn : nullptr, m->dn_on ? m->d_dn + m->M_loc : nullptr, m->dn_clamp);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_interleave(dr_model* m, const float* gathered, int B, int T, float* out) {
    size_t R = (size_t)B * T;
    int NL = m->M_loc * DR_Q;
    size_t total = R * NL * m->cfg.world;
    unsigned blocks = (unsigned)((total + 255) / 256);
    if (blocks > 148u * 16u) blocks = 148u * 16u;
    dr_interleave_kernel<<<blocks, 256, 0, m->stream>>>(gathered, out, R, NL, m->cfg.world);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_loss(dr_model* m, const float* out, const float* y, int B, int T, float* loss) {
    size_t n_rm = (size_t)B * T * m->cfg.M;
    double* acc = reinterpret_cast<double*>(m->d_loss);
    dr_loss_zero_kernel<<<1, 1, 0, m->stream>>>(acc);
    unsigned blocks = (unsigned)((n_rm + 255) / 256);
    if (blocks > 148u * 8u) blocks = 148u * 8u;
    dr_loss_kernel<<<blocks, 256, 0, m->stream>>>(out, y, n_rm, m->cfg.quantiles[0], m->cfg.quantiles[1],
                                                  m->cfg.quantiles[2], acc);
    dr_loss_final_kernel<<<1, 1, 0, m->stream>>>(acc, 1.0 / (double)n_rm, loss);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 3;
    return DR_OK;
}
