// K4c (bf16 training engine) — every weight-gradient reduction over the (t, window) rows of a micro-batch in ONE streaming
// tcgen05 kernel (SURVEY §8a "Backward": dW_hh += dgh (x) h_{t-1}, P = dgi^T x, db_ih = sum dgi, db_hh = sum dgh):
//
//     C[m][n] += sum_k A[k][m] * B[k][n]      k = window (the reduced index),  m = column of (da_r | da_z | da_n | dq),
//                                             n = column of [ h_prev (128) | x (64) | 1 (16, only column 0 is one) ]
//
// Both operands are the bf16 images the recurrence kernels wrote (dr_t16.cuh) consumed as MN-major SW128 operands: a
// chunk of 64 windows of one column block is 8 contiguous KB, so the producer thread moves operands with plain 1-D bulk
// copies — no conversion, no transposition, no second pass over the adjoints (the split-fp16 engine re-reads them twice
// and converts fp32 in flight).  The "ones" column turns the bias column sums into part of the same GEMM.
// One CTA = (expert, direction, half of the gate adjoints): hsel 0 -> {da_r: N = 208, da_n: N = 80 (x | 1)},
// hsel 1 -> {da_z: N = 208, dq: N = 208}; accumulators stay in TMEM (288 / 416 columns) over the whole micro-batch.
// Streaming kernel: 56 KB of operands per 64-window chunk against ~0.6 us of tensor time -> HBM bound by design.
#include "dr_t16.cuh"

using namespace drtc;
using namespace drt16;

namespace {

constexpr int kThreads = 192;                       // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue
constexpr int kStages = 3;
constexpr uint32_t kHalfBlk = 64 * 128;             // 64 windows of one column block = 8 KB (8 k-groups x 1 KB)
constexpr uint32_t kStageA = 4 * kHalfBlk;          // two adjoint tiles x two column blocks
constexpr uint32_t kStageB = 4 * kHalfBlk;          // h (2 blocks) | x | ones
constexpr uint32_t kStage = kStageA + kStageB;      // 64 KB
constexpr uint32_t kOffBar = kStages * kStage;
constexpr uint32_t kSmem = kOffBar + 128;
constexpr uint32_t kTxBytes = 7 * kHalfBlk;         // the ones block is filled once, not copied
enum WgBar { WG_FULL0 = 0, WG_EMPTY0 = kStages, WG_DFULL = 2 * kStages, WG_NUM };

struct Wg16Args {
    const uint8_t* gate;      // adjoint images [dir][e][t][tile] (da_r, da_z, da_n, dq)
    const uint8_t* himg;      // h images [dir][e][t][tile]
    const uint8_t* ximg;      // x images [t][tile]
    const uint8_t* zero;      // >= 8 KB of zeros (h_prev of the first processed step)
    float* grad;              // gradient blob of the shard (reference order)
    float* P;                 // [2 dir][M_loc][3H][F] accumulators of dgi^T x
    int off_whh[2], off_bih[2], off_bhh[2], per_expert;
    int T, M_loc, ntiles, F;
};

// MN-major SW128 descriptor: LBO = bytes between 64-column blocks, SBO = bytes between 8-window groups (pinned by
// tests/test_gpu_tc_probe.py::test_tcgen05_tile_mn_major)
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kThreads, 1) dr_wgrad16_kernel(Wg16Args g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hsel = blockIdx.x, e = blockIdx.y, dir = blockIdx.z;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WG_NUM);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    const int nchunks = g.T * g.ntiles * 2;
    const int n1 = hsel ? 208 : 80;                   // width of the second tile: dq x [h | x | 1]  or  da_n x [x | 1]

    if (tid == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(bar(WG_FULL0 + i), 1); mbar_init(bar(WG_EMPTY0 + i), 1); }
        mbar_init(bar(WG_DFULL), 1);
        fence_mbar_init();
    }
    // the "ones" block of every stage: logical column 0 of each window row = 1.0, everything else 0
    for (int i = tid; i < kStages * (int)(kHalfBlk / 16); i += kThreads) {
        const int st = i / (int)(kHalfBlk / 16), c16 = i % (int)(kHalfBlk / 16);
        const int k = c16 >> 3, phys = c16 & 7;       // window row inside the 64, physical 16-byte chunk
        const uint32_t first = ((phys ^ (k & 7)) == 0) ? 0x00003F80u : 0u;     // bf16 1.0 in the low half = column 0
        *reinterpret_cast<uint4*>(smem + (size_t)st * kStage + kStageA + 3 * kHalfBlk + (size_t)c16 * 16) = make_uint4(first, 0, 0, 0);
    }
    fence_proxy_async();
    if (warp == 1) { tmem_alloc<1>(smem_u32(tmem_slot), 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: 7 bulk copies of 8 KB per 64-window chunk =====================
        if (elect_one()) {
            const int cbA0 = hsel ? 2 : 0, cbA1 = hsel ? 6 : 4;      // first column block of the two adjoint tiles
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % kStages;
                if (c >= kStages) mbar_wait(bar(WG_EMPTY0 + st), (uint32_t)(((c / kStages) - 1) & 1));
                const int hw = c & 1, tile = (c >> 1) % g.ntiles, t = (c >> 1) / g.ntiles;
                const int tp = dir ? t + 1 : t - 1;                  // the step whose output is this step's h_prev
                const uint8_t* ga = g.gate + blk_index(dir, e, t, tile, g.M_loc, g.T, g.ntiles) * kGateImg + (size_t)hw * kHalfBlk;
                const bool edge = (tp < 0 || tp >= g.T);
                const uint8_t* hb = edge ? g.zero : g.himg + blk_index(dir, e, tp, tile, g.M_loc, g.T, g.ntiles) * kHImg + (size_t)hw * kHalfBlk;
                const uint8_t* xb = g.ximg + ((size_t)t * g.ntiles + tile) * kColBlk + (size_t)hw * kHalfBlk;
                const uint32_t sA = smem_u32(smem) + (uint32_t)st * kStage, sB = sA + kStageA;
                const uint32_t fb = bar(WG_FULL0 + st);
                mbar_expect_tx(fb, kTxBytes);
                bulk_g2s(sA, ga + (size_t)cbA0 * kColBlk, kHalfBlk, fb);
                bulk_g2s(sA + kHalfBlk, ga + (size_t)(cbA0 + 1) * kColBlk, kHalfBlk, fb);
                bulk_g2s(sA + 2 * kHalfBlk, ga + (size_t)cbA1 * kColBlk, kHalfBlk, fb);
                bulk_g2s(sA + 3 * kHalfBlk, ga + (size_t)(cbA1 + 1) * kColBlk, kHalfBlk, fb);
                bulk_g2s(sB, hb, kHalfBlk, fb);
                bulk_g2s(sB + kHalfBlk, edge ? hb : hb + kColBlk, kHalfBlk, fb);
                bulk_g2s(sB + 2 * kHalfBlk, xb, kHalfBlk, fb);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t mn = (1u << 15) | (1u << 16);             // A and B MN-major
            const uint32_t idesc0 = make_idesc_bf16(128, 208) | mn;
            const uint32_t idesc1 = make_idesc_bf16(128, n1) | mn;
            for (int c = 0; c < nchunks; ++c) {
                const int st = c % kStages;
                mbar_wait(bar(WG_FULL0 + st), (uint32_t)((c / kStages) & 1));
                tc_fence_after();
                const uint32_t sA = smem_u32(smem) + (uint32_t)st * kStage, sB = sA + kStageA;
                const uint64_t a0 = make_desc_mn(sA, kHalfBlk, 1024), a1 = make_desc_mn(sA + 2 * kHalfBlk, kHalfBlk, 1024);
                const uint64_t b0 = make_desc_mn(sB, kHalfBlk, 1024);
                const uint64_t b1 = make_desc_mn(sB + (hsel ? 0u : 2 * kHalfBlk), kHalfBlk, 1024);
#pragma unroll
                for (int k16 = 0; k16 < 4; ++k16) {                  // 16 windows = two 1 KB groups per step
                    mma_ss<1>(tbase, a0 + ((k16 * 2048) >> 4), b0 + ((k16 * 2048) >> 4), idesc0, (c | k16) ? 1u : 0u);
                    mma_ss<1>(tbase + 208, a1 + ((k16 * 2048) >> 4), b1 + ((k16 * 2048) >> 4), idesc1, (c | k16) ? 1u : 0u);
                }
                mma_commit_1(bar(WG_EMPTY0 + st));
            }
            mma_commit_1(bar(WG_DFULL));
        }
        __syncwarp();
    } else {
        // ===================== epilogue: accumulators -> gradient blob (+=) =====================
        mbar_wait(bar(WG_DFULL), 0);
        tc_fence_after();
        const int q4 = warp & 3;                                     // TMEM lane quarter this warp may read
        const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
        const int mrow = q4 * 32 + lane;                             // hidden unit of the gate
        float* ge = g.grad + (size_t)e * g.per_expert;
        float* Pe = g.P + ((size_t)dir * g.M_loc + e) * 3 * DR_H * g.F;
        auto add_cols = [&](uint32_t col0, int n, float* dst) {      // dst[0..n) += D[mrow][col0 .. col0+n)   (n % 16 == 0)
            for (int c0 = 0; c0 < n; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tbase + lane_base + col0 + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) dst[c0 + j] += __uint_as_float(v[j]);
            }
        };
        auto read_cols = [&](uint32_t col0, float (&out)[16]) {
            uint32_t v[16];
            tmem_ld16(tbase + lane_base + col0, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) out[j] = __uint_as_float(v[j]);
        };
        auto add_x = [&](uint32_t col0, int gate) {                   // P[gate*H + mrow][0..F) += 64 x columns
            float* dst = Pe + (size_t)(gate * DR_H + mrow) * g.F;
            for (int c0 = 0; c0 < 64; c0 += 16) {
                float v[16];
                read_cols(col0 + c0, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) if (c0 + j < g.F) dst[c0 + j] += v[j];
            }
        };
        float one[16];
        const int g0 = hsel;                                         // tile 0: da_r (gate 0) or da_z (gate 1)
        add_cols(0, 128, ge + g.off_whh[dir] + (size_t)(g0 * DR_H + mrow) * DR_H);
        add_x(128, g0);
        read_cols(192, one);
        ge[g.off_bih[dir] + g0 * DR_H + mrow] += one[0];
        ge[g.off_bhh[dir] + g0 * DR_H + mrow] += one[0];
        if (hsel == 0) {                                             // tile 1 = da_n: [x | 1]
            add_x(208, 2);
            read_cols(208 + 64, one);
            ge[g.off_bih[dir] + 2 * DR_H + mrow] += one[0];
        } else {                                                     // tile 1 = dq: [h | (x unused) | 1]
            add_cols(208, 128, ge + g.off_whh[dir] + (size_t)(2 * DR_H + mrow) * DR_H);
            read_cols(208 + 192, one);
            ge[g.off_bhh[dir] + 2 * DR_H + mrow] += one[0];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<1>(tbase, 512);
}

}  // namespace

int dr_launch_wgrad16(dr_model* m, const uint8_t* gate, const uint8_t* himg, const uint8_t* ximg, const uint8_t* zero,
                      float* P, int Bm, int T) {
    const int Ml = m->M_loc;
    if (Ml == 0 || Bm <= 0 || T <= 0) return DR_OK;
    Wg16Args g;
    g.gate = gate; g.himg = himg; g.ximg = ximg; g.zero = zero; g.grad = m->d_grad; g.P = P;
    for (int d = 0; d < 2; ++d) { g.off_whh[d] = m->off.w_hh[d]; g.off_bih[d] = m->off.b_ih[d]; g.off_bhh[d] = m->off.b_hh[d]; }
    g.per_expert = m->off.per_expert;
    g.T = T; g.M_loc = Ml; g.ntiles = (Bm + 127) / 128; g.F = m->cfg.F;
    DR_CUDA(m, cudaFuncSetAttribute(dr_wgrad16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    dim3 grid(2, Ml, 2);
    dr_wgrad16_kernel<<<grid, kThreads, kSmem, m->stream>>>(g);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
