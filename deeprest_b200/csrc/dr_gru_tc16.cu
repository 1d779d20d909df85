// K4a (bf16 training engine) — the train-mode forward of the bi-GRU (qrnn.py:33-43) on tcgen05, single-pass bf16 operands.
//
// Same recurrence, same epilogue structure and the same S / head-partial outputs as the inference kernel
// (dr_gru_tc.cu), with three differences that follow from bf16 operands:
//   * ONE tensor pass per product instead of three (hi only): 12 MMAs per hidden quarter instead of 36;
//   * the whole weight image of an expert-direction is 144 KB, so ONE CTA holds it: cta_group::1, M = 128 windows per
//     work item, no cluster, no remote barriers — a micro-batch of 128 windows fills a tile (the 2-CTA kernel needs 256);
//   * the epilogue applies dropout (qrnn.py:43) before the cross-expert sum and the own-expert head term, and saves what
//     the backward pass needs — (r, z, n, q = W_hn h + b_hn) and h_t — as bf16 operand images (dr_t16.cuh), one
//     128-byte row per thread and column block, written once.
//   shared memory : weights 4 quarters x {Wx, Wh k-block 0, Wh k-block 1} x [96 rows x 128 B]  = 144 KB (4 bulk copies),
//                   x tiles 2 stages x 16 KB (bulk copies of the x image), biases, head coefficients, barriers
//   TMEM (512 col): 2 gate buffers x 128 fp32 columns [gi_n | r | z | gh_n], 2 h-operand buffers x 64 columns (bf16 pairs)
//   warps         : 0-7 gate epilogue, 8 MMA issuer, 9 bulk-copy producer, 10-11 register donors
#include "dr_t16.cuh"

using namespace drtc;
using namespace drt16;

namespace {

constexpr int kThreads = 384;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kEpiWarps = 8, kMmaWarp = 8, kLoadWarp = 9;
constexpr uint32_t kWBlk = 96 * 128;                  // one B block: 96 rows (D columns of a quarter) x 64 K bf16 = 12 KB
constexpr uint32_t kWQuarter = 3 * kWBlk;             // Wx, Wh[k 0..63], Wh[k 64..127]
constexpr uint32_t kWImg = 4 * kWQuarter;             // 147456 B per expert-direction
constexpr uint32_t kXTile = kColBlk;                  // 128 windows x 64 features bf16
constexpr uint32_t kG0 = 0, kHA = 256, kHB = 320;     // TMEM columns
constexpr uint32_t kOffW = 0;
constexpr uint32_t kOffX = kWImg;
constexpr uint32_t kOffBias = kOffX + 2 * kXTile;
constexpr uint32_t kOffCt = kOffBias + 4 * DR_H * 4;
constexpr uint32_t kOffBar = kOffCt + DR_Q * DR_H * 4;
constexpr uint32_t kSmemBytes = kOffBar + 256;

enum Bar { GATE_FULL0 = 0, GATE_FULL1, GATE_FREE0, GATE_FREE1, H_READY0, H_READY1, H_READY2, H_READY3,
           X_FULL0, X_FULL1, X_FREE0, X_FREE1, W_FULL, NUM_BARS };

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void store_bhn(uint32_t taddr, const float* src) {
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    const float4 c = *reinterpret_cast<const float4*>(src + 8), d = *reinterpret_cast<const float4*>(src + 12);
    uint32_t v0[8] = {__float_as_uint(a.x), __float_as_uint(a.y), __float_as_uint(a.z), __float_as_uint(a.w),
                      __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w)};
    uint32_t v1[8] = {__float_as_uint(c.x), __float_as_uint(c.y), __float_as_uint(c.z), __float_as_uint(c.w),
                      __float_as_uint(d.x), __float_as_uint(d.y), __float_as_uint(d.z), __float_as_uint(d.w)};
    tmem_st8(taddr, v0);
    tmem_st8(taddr + 8, v1);
}

struct Fwd16Args {
    const uint8_t* wimg;      // [M_loc][2][kWImg]
    const uint8_t* ximg;      // [T][ntiles][kXTile]
    const float* bias4;       // [M_loc][2][4][H]
    const float* ct;          // [M_loc][2][Q][H]
    float* S;                 // [T][64][Bp][4]   (zeroed by the caller)
    float* P;                 // own-expert head partials, layout of dr_gru_tc.cu
    uint8_t* gate;            // gate images [dir][e][t][tile]
    uint8_t* himg;            // h images    [dir][e][t][tile]
    Drop drop;
    int B, T, Bp, M_loc, ntiles;
    int e_lo;                 // first global expert of this shard, b0 / Bfull: position of the micro-batch inside the full
    int b0, Bfull;            //   batch — both only address the dropout element index ((e*Bfull + b)*T + t)*2H + k
};

__global__ void __launch_bounds__(kThreads, 1) dr_gru_tc16_kernel(Fwd16Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x;                       // (tile, e, dir), tile-major
    const int tile = item / (2 * a.M_loc);
    const int e = (item % (2 * a.M_loc)) >> 1;
    const int dir = item & 1;
    const int T = a.T, B = a.B, Bp = a.Bp;

    float* bs = reinterpret_cast<float*>(smem + kOffBias);
    float* cs = reinterpret_cast<float*>(smem + kOffCt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    for (int i = tid; i < 4 * DR_H; i += kThreads) {
        float v = a.bias4[(size_t)(e * 2 + dir) * 4 * DR_H + i];
        int c = i / DR_H;
        bs[i] = (c < 2) ? -v * kLog2e : (c == 2) ? 2.0f * kLog2e * v : v;
    }
    for (int i = tid; i < DR_Q * DR_H; i += kThreads) cs[i] = a.ct[(size_t)(e * 2 + dir) * DR_Q * DR_H + i];
    if (tid == 0) {
        mbar_init(bar(GATE_FULL0), 1); mbar_init(bar(GATE_FULL1), 1);
        mbar_init(bar(GATE_FREE0), kEpiWarps); mbar_init(bar(GATE_FREE1), kEpiWarps);
        for (int i = 0; i < 4; ++i) mbar_init(bar(H_READY0 + i), kEpiWarps);
        mbar_init(bar(X_FULL0), 1); mbar_init(bar(X_FULL1), 1);
        mbar_init(bar(X_FREE0), 1); mbar_init(bar(X_FREE1), 1);
        mbar_init(bar(W_FULL), 1);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) { tmem_alloc<1>(smem_u32(tmem_slot), 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp < kEpiWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else                  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp < kEpiWarps) {
        // ======================= gate epilogue warps =======================
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                   // TMEM lane == window row of the tile
        const int b = tile * 128 + row;
        const bool live = b < B;
        {
            uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 32; c += 8) tmem_st8(tbase + lane_base + kHA + half * 32 + c, z8);   // h0 = 0 (qrnn.py:39)
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mbar_arrive(bar(H_READY0 + i));
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g)
            store_bhn(tbase + lane_base + kG0 + g * 128 + 96 + half * 16, bs + 3 * DR_H + g * 32 + half * 16);
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive(bar(GATE_FREE0)); mbar_arrive(bar(GATE_FREE1)); }

        float hreg[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) hreg[q][j] = 0.0f;
        uint32_t full_phase[2] = {0, 0};
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        float hn[16];
        // element index of (this expert, this window, step 0, column 0) in the reference's rnn_out order [M][B][T][2H]
        const size_t drop_base = (((size_t)(a.e_lo + e) * a.Bfull + (size_t)(a.b0 + (live ? b : 0))) * T) * DR_2H + (size_t)dir * DR_H;
        // dropout (qrnn.py:43) + cross-expert sum + own-expert head term of 16 hidden units of step ttp (see dr_gru_tc.cu::tail)
        auto tail = [&](int u0p, int ttp) {
            const uint32_t kb = keep16(a.drop, drop_base + (size_t)ttp * DR_2H + u0p);
            float rt[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) rt[j] = ((kb >> j) & 1u) ? hn[j] * a.drop.inv_keep : 0.0f;
            float p0[4] = {0.f, 0.f, 0.f, 0.f}, p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
                const float4 c0 = *reinterpret_cast<const float4*>(cs + u0p + j4);
                const float4 c1 = *reinterpret_cast<const float4*>(cs + DR_H + u0p + j4);
                const float4 c2 = *reinterpret_cast<const float4*>(cs + 2 * DR_H + u0p + j4);
                p0[0] = fmaf(c0.x, rt[j4], p0[0]); p0[1] = fmaf(c0.y, rt[j4 + 1], p0[1]); p0[2] = fmaf(c0.z, rt[j4 + 2], p0[2]); p0[3] = fmaf(c0.w, rt[j4 + 3], p0[3]);
                p1[0] = fmaf(c1.x, rt[j4], p1[0]); p1[1] = fmaf(c1.y, rt[j4 + 1], p1[1]); p1[2] = fmaf(c1.z, rt[j4 + 2], p1[2]); p1[3] = fmaf(c1.w, rt[j4 + 3], p1[3]);
                p2[0] = fmaf(c2.x, rt[j4], p2[0]); p2[1] = fmaf(c2.y, rt[j4 + 1], p2[1]); p2[2] = fmaf(c2.z, rt[j4 + 2], p2[2]); p2[3] = fmaf(c2.w, rt[j4 + 3], p2[3]);
            }
            o0 += (p0[0] + p0[1]) + (p0[2] + p0[3]);
            o1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
            o2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
            if (live) {
                float* sp = a.S + (((size_t)ttp * 64 + dir * 32 + u0p / 4) * Bp + b) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dr_red_add_v4(sp + (size_t)j * Bp * 4, rt[4 * j], rt[4 * j + 1], rt[4 * j + 2], rt[4 * j + 3]);
            }
        };
        auto flush = [&](int ttp) {
            if (live) {
                const int ngrp = (a.M_loc * DR_Q + 15) >> 4;
                const size_t slab = (((size_t)ttp * (Bp >> 7) + (b >> 7)) * ngrp) * 4 * 16 * 128;
                const int dh = dir * 2 + half;
                const int c = e * DR_Q;
                float* o = a.P + slab + (b & 127);
                o[((size_t)((c >> 4) * 4 + dh) * 16 + (c & 15)) * 128] = o0;
                o[((size_t)(((c + 1) >> 4) * 4 + dh) * 16 + ((c + 1) & 15)) * 128] = o1;
                o[((size_t)(((c + 2) >> 4) * 4 + dh) * 16 + ((c + 2) & 15)) * 128] = o2;
            }
            o0 = 0.f; o1 = 0.f; o2 = 0.f;
        };
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? (T - 1 - s) : s;
            const int tt_prev = dir ? (T - s) : (s - 1);
            const uint32_t hnext = (s & 1) ? kHA : kHB;
            const size_t blk = blk_index(dir, e, tt, tile, a.M_loc, T, a.ntiles);
            uint8_t* gimg = a.gate + blk * kGateImg;
            uint8_t* himg = a.himg + blk * kHImg;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int buf = q & 1;
                const int u0 = q * 32 + half * 16;
                const uint32_t G = tbase + lane_base + kG0 + buf * 128 + half * 16;
                mbar_wait(bar(GATE_FULL0 + buf), full_phase[buf]);
                full_phase[buf] ^= 1;
                tc_fence_after();
                uint32_t gi[16], gr[16], gz[16], gh[16];
                tmem_ld16(G + 0, gi); tmem_ld16(G + 32, gr); tmem_ld16(G + 64, gz); tmem_ld16(G + 96, gh);
                tc_wait_ld();
                {
                    const int qn = (q + 2) & 3;
                    store_bhn(G + 96, bs + 3 * DR_H + qn * 32 + half * 16);
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar(GATE_FREE0 + buf));
                }
                if (half == 1 && (s > 0 || q > 0)) {               // lagging warp: previous quarter's tail first
                    tail(((q + 3) & 3) * 32 + 16, q == 0 ? tt_prev : tt);
                    if (q == 0) flush(tt_prev);
                }
                uint32_t ph[8];
                uint32_t pr[8], pz[8], pn[8], pq[8];             // bf16 pairs for the saved images
#pragma unroll
                for (int j8 = 0; j8 < 16; j8 += 8) {
                    float cr[8], cz[8], cn[8];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const float4 aa = *reinterpret_cast<const float4*>(bs + u0 + j8 + 4 * v);
                        const float4 b4 = *reinterpret_cast<const float4*>(bs + DR_H + u0 + j8 + 4 * v);
                        const float4 c = *reinterpret_cast<const float4*>(bs + 2 * DR_H + u0 + j8 + 4 * v);
                        cr[4 * v] = aa.x; cr[4 * v + 1] = aa.y; cr[4 * v + 2] = aa.z; cr[4 * v + 3] = aa.w;
                        cz[4 * v] = b4.x; cz[4 * v + 1] = b4.y; cz[4 * v + 2] = b4.z; cz[4 * v + 3] = b4.w;
                        cn[4 * v] = c.x; cn[4 * v + 1] = c.y; cn[4 * v + 2] = c.z; cn[4 * v + 3] = c.w;
                    }
                    float er[8], ez[8], rr[8], zz[8], en[8], rv[8], nv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        er[i] = fminf(fmaf(__uint_as_float(gr[j8 + i]), -kLog2e, cr[i]), 30.0f);
                        ez[i] = fminf(fmaf(__uint_as_float(gz[j8 + i]), -kLog2e, cz[i]), 30.0f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { er[i] = ex2_approx(er[i]); ez[i] = ex2_approx(ez[i]); }
#pragma unroll
                    for (int i = 0; i < 8; ++i) { er[i] += 1.0f; ez[i] += 1.0f; rr[i] = er[i] * ez[i]; }
                    float iv2[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) iv2[i] = rcp_approx(rr[2 * i] * rr[2 * i + 1]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float inv = rr[i ^ 1] * iv2[i >> 1];          // 1/((1+er_i)(1+ez_i))
                        zz[i] = er[i] * inv;
                        rv[i] = ez[i] * inv;
                        const float t = fmaf(rv[i], __uint_as_float(gh[j8 + i]), __uint_as_float(gi[j8 + i]));
                        en[i] = fminf(fmaf(t, 2.0f * kLog2e, cn[i]), 30.0f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) en[i] = ex2_approx(en[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) en[i] += 1.0f;
                    float ivn[4];
#pragma unroll
                    for (int g4 = 0; g4 < 2; ++g4) {
                        const float pa = en[4 * g4] * en[4 * g4 + 1], pb = en[4 * g4 + 2] * en[4 * g4 + 3];
                        const float inv4 = rcp_approx(pa * pb);
                        ivn[2 * g4] = pb * inv4;
                        ivn[2 * g4 + 1] = pa * inv4;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const int j = j8 + i;
                        const float n0 = fmaf(-2.0f * en[i + 1], ivn[i >> 1], 1.0f);
                        const float n1 = fmaf(-2.0f * en[i], ivn[i >> 1], 1.0f);
                        const float a0 = __fadd_rn(__fmul_rn(__fsub_rn(hreg[q][j], n0), zz[i]), n0);
                        const float a1 = __fadd_rn(__fmul_rn(__fsub_rn(hreg[q][j + 1], n1), zz[i + 1]), n1);
                        hreg[q][j] = a0; hreg[q][j + 1] = a1;
                        hn[j] = a0; hn[j + 1] = a1;
                        nv[i] = n0; nv[i + 1] = n1;
                        ph[j >> 1] = pack_bf2(a0, a1);
                    }
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const int j = j8 + i;
                        pr[j >> 1] = pack_bf2(rv[i], rv[i + 1]);
                        pz[j >> 1] = pack_bf2(zz[i], zz[i + 1]);
                        pn[j >> 1] = pack_bf2(nv[i], nv[i + 1]);
                        pq[j >> 1] = pack_bf2(__uint_as_float(gh[j]), __uint_as_float(gh[j + 1]));
                    }
                }
                tmem_st8(tbase + lane_base + hnext + u0 / 2, ph);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(H_READY0 + q));
                // saved activations: 16 columns = two 16-byte chunks per gate.  Rows past the batch are written as zeros: the
                // weight-gradient GEMMs read whole 64-window chunks.
                {
                    const int cb = u0 >> 6, sc = (u0 & 63) >> 4;           // column block, 32-byte sector of the row (16 columns)
                    if (!live) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { pr[i] = 0u; pz[i] = 0u; pn[i] = 0u; pq[i] = 0u; ph[i] = 0u; }
                    }
                    uint8_t* g0 = gimg + (size_t)cb * kColBlk;
                    st_cols16(g0, row, sc, pr);                            // one 256-bit store per array (see dr_t16.cuh)
                    st_cols16(g0 + 2 * kColBlk, row, sc, pz);
                    st_cols16(g0 + 4 * kColBlk, row, sc, pn);
                    st_cols16(g0 + 6 * kColBlk, row, sc, pq);
                    st_cols16(himg + (size_t)cb * kColBlk, row, sc, ph);
                }
                if (half == 0) { tail(u0, tt); if (q == 3) flush(tt); }
            }
        }
        if (half == 1) { const int tl = dir ? 0 : (T - 1); tail(3 * 32 + 16, tl); flush(tl); }
    } else if (warp == kMmaWarp) {
        // ======================= MMA issuer =======================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, 96);
            const uint32_t w_s = smem_u32(smem + kOffW);
            const uint32_t x_s = smem_u32(smem + kOffX);
            mbar_wait(bar(W_FULL), 0);
            uint32_t free_bits = 0;
            for (int s = 0; s < T; ++s) {
                const uint32_t hcur = tbase + ((s & 1) ? kHB : kHA);
                const uint32_t xst = x_s + (s & 1) * kXTile;
                mbar_wait(bar(X_FULL0 + (s & 1)), (s >> 1) & 1);
                const uint64_t xdesc = make_desc_sw128(xst);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const int buf = q & 1;
                    const uint32_t G = tbase + kG0 + buf * 128;
                    const uint64_t wdesc = make_desc_sw128(w_s + q * kWQuarter);
                    mbar_wait(bar(GATE_FREE0 + buf), (free_bits >> buf) & 1u);
                    free_bits ^= 1u << buf;
                    tc_fence_after();
                    // x-part: D[:, 0:96] = x_t * [W_in | W_ir | W_iz]_q^T
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        mma_ss<1>(G, xdesc + ((k16 * 32) >> 4), wdesc + ((k16 * 32) >> 4), idesc, k16 ? 1u : 0u);
                    // h-part: D[:, 32:128] += h_{t-1} * [W_hr | W_hz | W_hn]_q^T, A from TMEM, K walked in quarters of 32
#pragma unroll
                    for (int kq = 0; kq < 4; ++kq) {
                        if (q == 0) { mbar_wait(bar(H_READY0 + kq), s & 1); tc_fence_after(); }
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const int ks = kq * 2 + j, kb = ks >> 2, k16 = ks & 3;
                            mma_ts<1>(G + 32, hcur + (kb * 64 + k16 * 16) / 2,
                                      wdesc + (((1 + kb) * kWBlk + k16 * 32) >> 4), idesc, 1u);
                        }
                    }
                    mma_commit_1(bar(GATE_FULL0 + buf));
                }
                mma_commit_1(bar(X_FREE0 + (s & 1)));
            }
        }
        __syncwarp();
    } else {
        // ======================= bulk-copy producer =======================
        if (warp == kLoadWarp && elect_one()) {
            const uint8_t* wsrc = a.wimg + (size_t)(e * 2 + dir) * kWImg;
            mbar_expect_tx(bar(W_FULL), kWImg);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                bulk_g2s(smem_u32(smem + kOffW) + i * kWQuarter, wsrc + (size_t)i * kWQuarter, kWQuarter, bar(W_FULL));
            for (int s = 0; s < T; ++s) {
                const int st = s & 1;
                const int tt = dir ? (T - 1 - s) : s;
                if (s >= 2) mbar_wait(bar(X_FREE0 + st), ((s >> 1) - 1) & 1);
                mbar_expect_tx(bar(X_FULL0 + st), kXTile);
                bulk_g2s(smem_u32(smem + kOffX) + st * kXTile, a.ximg + ((size_t)tt * a.ntiles + tile) * kXTile, kXTile, bar(X_FULL0 + st));
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) tmem_dealloc<1>(tbase, 512);
}

// weights: one thread per (e, d, q, block kk in 0..2 (Wx, Wh k-block 0, Wh k-block 1), row 0..95, chunk8)
// row = D column inside the 96-wide MMA: x-part rows [gi_n | r | z], h-part rows [r | z | gh_n] (see dr_gru_tc.cu)
__global__ void dr_tc16_pack_w_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F,
                                      const float* __restrict__ mask, uint8_t* __restrict__ wimg, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int chunk = (int)(i % 8); size_t r = i / 8;
    int row = (int)(r % 96); r /= 96;
    int kk = (int)(r % 3); r /= 3;
    int q = (int)(r % 4); r /= 4;
    int d = (int)(r % 2); r /= 2;
    int e = (int)r;
    const float* ex = blob + (size_t)e * off.per_expert;
    const int grp = row / 32, unit = q * 32 + row % 32;
    float v[8];
    if (kk == 0) {
        const int gate = (grp == 0) ? 2 : grp - 1;
        const float* w = ex + off.w_ih[d] + (size_t)(gate * DR_H + unit) * F;
        for (int j = 0; j < 8; ++j) {
            const int k = chunk * 8 + j;
            v[j] = (k < F) ? w[k] * mask[(size_t)e * F + k] : 0.0f;           // mask folded: W_ih' = W_ih diag(mask)
        }
    } else {
        const float* w = ex + off.w_hh[d] + (size_t)(grp * DR_H + unit) * DR_H + (kk - 1) * 64;
        for (int j = 0; j < 8; ++j) v[j] = w[chunk * 8 + j];
    }
    uint8_t* base = wimg + (size_t)(e * 2 + d) * kWImg + (size_t)q * kWQuarter + (size_t)kk * kWBlk;
    *reinterpret_cast<uint4*>(base + sw128_offset(row, chunk * 8)) =
        make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}

// x [B,T,F] fp32 (rows b0.. of the batch) -> x image [T][ntiles][128 windows x 64 features] bf16, zero padded
__global__ void dr_tc16_pack_x_kernel(const float* __restrict__ x, uint8_t* __restrict__ ximg, int B, int T, int F, int ntiles) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * ntiles * 128 * 8;
    if (i >= total) return;
    int chunk = (int)(i % 8); size_t r = i / 8;
    int row = (int)(r % 128); r /= 128;
    int tile = (int)(r % ntiles);
    int t = (int)(r / ntiles);
    int b = tile * 128 + row;
    float v[8];
    for (int j = 0; j < 8; ++j) {
        int f = chunk * 8 + j;
        v[j] = (b < B && f < F) ? x[((size_t)b * T + t) * F + f] : 0.0f;
    }
    *reinterpret_cast<uint4*>(ximg + ((size_t)t * ntiles + tile) * kXTile + img_off(row, chunk)) =
        make_uint4(pack_bf2(v[0], v[1]), pack_bf2(v[2], v[3]), pack_bf2(v[4], v[5]), pack_bf2(v[6], v[7]));
}

}  // namespace

size_t dr_t16_wimg_bytes(int M_loc) { return (size_t)M_loc * 2 * kWImg; }

int dr_t16_pack_weights(dr_model* m, uint8_t* wimg) {
    if (m->M_loc == 0) return DR_OK;
    size_t total = (size_t)m->M_loc * 2 * 4 * 3 * 96 * 8;
    dr_tc16_pack_w_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_blob, m->off, m->cfg.F, m->d_mask, wimg, total);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_t16_pack_x(dr_model* m, const float* x_mb, int Bm, int T, uint8_t* ximg) {
    const int ntiles = (Bm + 127) / 128;
    size_t total = (size_t)T * ntiles * 128 * 8;
    dr_tc16_pack_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(x_mb, ximg, Bm, T, m->cfg.F, ntiles);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

// train-mode forward of one micro-batch (Bm windows starting at b0 of the full batch of Bfull): S (zeroed here), head
// partials P, gate / h images
int dr_launch_gru_tc16(dr_model* m, const uint8_t* wimg, const uint8_t* ximg, int Bm, int T, float* S, float* P,
                       uint8_t* gate, uint8_t* himg, const uint8_t* mask, uint64_t seed, int b0, int Bfull) {
    if (m->M_loc == 0 || Bm <= 0 || T <= 0) return DR_OK;
    Fwd16Args a;
    a.wimg = wimg; a.ximg = ximg; a.bias4 = m->d_bias4; a.ct = m->d_ct; a.S = S; a.P = P; a.gate = gate; a.himg = himg;
    const float p = m->cfg.dropout_p;
    a.drop.mask = mask; a.drop.seed = seed; a.drop.inv_keep = 1.0f / (1.0f - p);
    a.drop.thr16 = (uint32_t)(p * 65536.0f + 0.5f);
    a.B = Bm; a.T = T; a.Bp = (Bm + 127) / 128 * 128; a.M_loc = m->M_loc; a.ntiles = (Bm + 127) / 128;
    a.e_lo = m->e_lo; a.b0 = b0; a.Bfull = Bfull;
    DR_CUDA(m, cudaMemsetAsync(S, 0, dr_s_floats(Bm, T) * sizeof(float), m->stream));
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_tc16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    dr_gru_tc16_kernel<<<m->M_loc * 2 * a.ntiles, kThreads, kSmemBytes, m->stream>>>(a);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
