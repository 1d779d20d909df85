// K0 — one-time weight preparation (runs inside dr_load_weights).
//
//  * feature mask  softmax(W2·relu(W1·1+b1)+b2)                 qrnn.py:34  (input independent)
//  * mask folded into the input projection  W_ih' = W_ih·diag(mask)   (inference)
//  * weights re-laid as the k-major stream the FFMA recurrence kernel consumes
//  * gate biases pre-summed where the GRU equations allow it (r, z) and kept apart for n
//  * head split  W_i = [A_i | C_i]  (qrnn.py:53-54):  y_i = (A_i/(M-1))·S + (C_i - A_i/(M-1))·r_i + b_i
#include "dr_common.cuh"

// one block per local expert, DR_H threads
__global__ void dr_mask_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F,
                               float* __restrict__ mask) {
    extern __shared__ float sm[];          // hid[H] | logits[F]
    float* hid = sm;
    float* logit = sm + DR_H;
    const float* ex = blob + (size_t)blockIdx.x * off.per_expert;
    int tid = threadIdx.x;
    // Linear(1,H) on mask_init == 1, then ReLU
    hid[tid] = fmaxf(ex[off.mask_w1 + tid] * 1.0f + ex[off.mask_b1 + tid], 0.0f);
    __syncthreads();
    for (int f = tid; f < F; f += blockDim.x) {
        const float* w = ex + off.mask_w2 + (size_t)f * DR_H;
        float acc = 0.0f;
        for (int k = 0; k < DR_H; ++k) acc = fmaf(w[k], hid[k], acc);
        logit[f] = acc + ex[off.mask_b2 + f];
    }
    __syncthreads();
    // F is small (tens); every thread redoes the scalar max/sum — no cross-thread reduction needed
    float mx = -INFINITY;
    for (int f = 0; f < F; ++f) mx = fmaxf(mx, logit[f]);
    float sum = 0.0f;
    for (int f = 0; f < F; ++f) sum += expf(logit[f] - mx);
    for (int f = tid; f < F; f += blockDim.x)
        mask[(size_t)blockIdx.x * F + f] = expf(logit[f] - mx) / sum;
}

// FFMA stream: wf[e][d][p][k][g][jj],  k < Fp: W_ih[g*H + p*64 + jj][k]*mask[k] (0 for k >= F)
//                                       k >= Fp: W_hh[g*H + p*64 + jj][k - Fp]
__global__ void dr_pack_ffma_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F, int Fp,
                                    const float* __restrict__ mask, float* __restrict__ wf,
                                    size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int KT = Fp + DR_H;
    int jj = (int)(i % 64); size_t r = i / 64;
    int g = (int)(r % 3); r /= 3;
    int k = (int)(r % KT); r /= KT;
    int p = (int)(r % 2); r /= 2;
    int d = (int)(r % 2); r /= 2;
    int e = (int)r;
    const float* ex = blob + (size_t)e * off.per_expert;
    int row = g * DR_H + p * 64 + jj;
    float v;
    if (k < Fp) v = (k < F) ? ex[off.w_ih[d] + (size_t)row * F + k] * mask[(size_t)e * F + k] : 0.0f;
    else        v = ex[off.w_hh[d] + (size_t)row * DR_H + (k - Fp)];
    wf[i] = v;
}

// bias4[e][d][4][H]: 0: b_ir+b_hr   1: b_iz+b_hz   2: b_in   3: b_hn
// ct[e][d][q][H]   : C[q][d*H+j] - A[q][d*H+j]/(M-1)
// abar[e*Q+q][2H]  : A[q][:]/(M-1)             hb[e*Q+q]: head bias
__global__ void dr_pack_small_kernel(const float* __restrict__ blob, DrBlobOffsets off, int M_loc,
                                     float inv_m1, float* __restrict__ bias4, float* __restrict__ ct,
                                     float* __restrict__ abar, float* __restrict__ hb) {
    int e = blockIdx.x;
    const float* ex = blob + (size_t)e * off.per_expert;
    for (int i = threadIdx.x; i < 2 * 4 * DR_H; i += blockDim.x) {
        int j = i % DR_H, c = (i / DR_H) % 4, d = i / (4 * DR_H);
        const float* bi = ex + off.b_ih[d];
        const float* bh = ex + off.b_hh[d];
        float v = (c == 0) ? bi[j] + bh[j]
                : (c == 1) ? bi[DR_H + j] + bh[DR_H + j]
                : (c == 2) ? bi[2 * DR_H + j] : bh[2 * DR_H + j];
        bias4[(size_t)e * 2 * 4 * DR_H + i] = v;
    }
    const float* hw = ex + off.head_w;           // [Q][4H]: cols 0..2H-1 = A (mean of others), 2H.. = C (own)
    for (int i = threadIdx.x; i < 2 * DR_Q * DR_H; i += blockDim.x) {
        int j = i % DR_H, q = (i / DR_H) % DR_Q, d = i / (DR_Q * DR_H);
        float a = hw[(size_t)q * 4 * DR_H + d * DR_H + j];
        float c = hw[(size_t)q * 4 * DR_H + DR_2H + d * DR_H + j];
        ct[(size_t)e * 2 * DR_Q * DR_H + i] = c - a * inv_m1;
    }
    for (int i = threadIdx.x; i < DR_Q * DR_2H; i += blockDim.x) {
        int k = i % DR_2H, q = i / DR_2H;
        abar[((size_t)e * DR_Q + q) * DR_2H + k] = hw[(size_t)q * 4 * DR_H + k] * inv_m1;
    }
    if (threadIdx.x < DR_Q) hb[e * DR_Q + threadIdx.x] = ex[off.head_b + threadIdx.x];
}

// wihm[d][e][i][f] = W_ih[d][e][i][f] * mask[e][f]   (training-path input projection, same folding as inference)
__global__ void dr_pack_wihm_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F, int M_loc,
                                    const float* __restrict__ mask, float* __restrict__ wihm, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int f = (int)(i % F); size_t r = i / F;
    int row = (int)(r % (3 * DR_H)); r /= 3 * DR_H;
    int e = (int)(r % M_loc); int d = (int)(r / M_loc);
    wihm[i] = blob[(size_t)e * off.per_expert + off.w_ih[d] + (size_t)row * F + f] * mask[(size_t)e * F + f];
}

int dr_launch_prep(dr_model* m) {
    int F = m->cfg.F, Fp = m->Fp, Ml = m->M_loc;
    if (Ml == 0) return DR_OK;
    size_t smem = (DR_H + F) * sizeof(float);
    dr_mask_kernel<<<Ml, DR_H, smem, m->stream>>>(m->d_blob, m->off, F, m->d_mask);
    DR_CUDA(m, cudaGetLastError());
    size_t total = (size_t)Ml * 2 * 2 * (Fp + DR_H) * 3 * 64;
    unsigned blocks = (unsigned)((total + 255) / 256);
    dr_pack_ffma_kernel<<<blocks, 256, 0, m->stream>>>(m->d_blob, m->off, F, Fp, m->d_mask, m->d_wf, total);
    DR_CUDA(m, cudaGetLastError());
    {
        size_t tw = (size_t)2 * Ml * 3 * DR_H * F;
        dr_pack_wihm_kernel<<<(unsigned)((tw + 255) / 256), 256, 0, m->stream>>>(m->d_blob, m->off, F, Ml, m->d_mask, m->d_wihm, tw);
        DR_CUDA(m, cudaGetLastError());
    }
    float inv_m1 = 1.0f / (float)(m->cfg.M - 1);
    dr_pack_small_kernel<<<Ml, 256, 0, m->stream>>>(m->d_blob, m->off, Ml, inv_m1, m->d_bias4, m->d_ct,
                                                    m->d_abar, m->d_hb);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 4;
    return DR_OK;
}
