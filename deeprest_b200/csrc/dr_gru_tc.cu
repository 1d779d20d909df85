// K1 (tensor-core engine) — fused bidirectional-GRU recurrence on tcgen05 / TMEM.
//
// Same contract as the FFMA engine (dr_gru_ffma.cu): for every local expert and both directions,
// run the GRU over T steps (qrnn.py:33-42), add h_t into the cross-expert sum S and store the
// own-expert head term into the partials workspace P for K2 (qrnn.py:46-54 folded).
//
// fp32 parity on 16-bit tensor cores: every operand is split v = hi + lo (two fp16, ~22 mantissa
// bits) and each product is formed as hi*hi + hi*lo + lo*hi with fp32 accumulation in TMEM
// ("3-pass"); the dropped lo*lo term is ~2^-24 relative.  fp16 rather than bf16 pairs: the operands
// are O(1) or smaller (h in (-1,1), normalised rates, weights), so fp16's 3 extra mantissa bits per
// piece buy 64x on the absolute error of a gate pre-activation; inputs are clamped to +-65504.
//
// One work item = (expert, direction, 256-window pair tile), run by a 2-CTA cluster with
// cta_group::2 MMAs (M = 256: 128 windows per CTA).  Why a pair: the split weight image of one
// expert-direction is 288 KB — it only fits on chip when each SM holds half of the B operand.
//   shared memory / CTA : weight image 144 KB (4 hidden-quarters x {Wx hi,lo ; Wh hi,lo x 2 K-blocks}),
//                         x tiles 2 stages x {hi,lo} x 16 KB, all pre-swizzled SW128 K-major images
//                         moved by 1-D bulk (TMA engine) copies.
//   TMEM / CTA (512 col): 2 gate buffers x 128 fp32 columns [gi_n | r | z | gh_n] x 32 hidden units,
//                         2 h-operand buffers x 128 columns (fp16 hi | lo, two per column): the
//                         recurrent A operand never touches shared memory (tcgen05.st -> MMA.TS).
//   warps               : 0-7 gate epilogue (TMEM lane quarter = w%4, hidden half = w/4), 8 MMA issuer (leader CTA),
//                         9 bulk-copy producer, 10-11 register donors (setmaxnreg: 216 regs for the epilogue warpgroups).
//   work order          : tile-major (all expert-directions of a 256-window tile before the next), so the slice of S the
//                         running clusters reduce into stays L2-resident.
// Per step and hidden-quarter q:  x-part  D[:,0:96]   = x_t  * [W_in|W_ir|W_iz]_q^T   (A from smem)
//                                 h-part  D[:,32:128] += h    * [W_hr|W_hz|W_hn]_q^T   (A from TMEM)
// so r and z accumulate both parts while gi_n / gh_n stay separate (n = tanh(gi_n + r*gh_n)).
#include "dr_common.cuh"
#include "dr_tc.cuh"

using namespace drtc;

namespace {

constexpr int kThreads = 384;          // 3 warpgroups: 2 x epilogue (warps 0-7), 1 x {MMA issuer, producer, 2 idle}
constexpr float kLog2e = 1.4426950408889634f;
constexpr int kEpiWarps = 8;
constexpr int kMmaWarp = 8, kLoadWarp = 9;
constexpr uint32_t kBlk = 48 * 128;                 // one B block: 48 rows x 64 K (fp16) = 6 KB
constexpr uint32_t kQuarterBytes = 6 * kBlk;        // Wx_hi, Wx_lo, Wh_hi[2], Wh_lo[2]
constexpr uint32_t kWBytes = 4 * kQuarterBytes;     // 147456 per CTA
constexpr uint32_t kXTile = 128 * 128;              // one x part: 128 rows x 64 features (fp16) = 16 KB
constexpr uint32_t kXStage = 2 * kXTile;            // hi + lo
// TMEM columns
constexpr uint32_t kG0 = 0, kHA = 256, kHB = 384;
// shared memory map
constexpr uint32_t kOffW = 0;
constexpr uint32_t kOffX = kWBytes;                         // 2 stages
constexpr uint32_t kOffBias = kOffX + 2 * kXStage;          // 4*H floats
constexpr uint32_t kOffCt = kOffBias + 4 * DR_H * 4;        // Q*H floats
constexpr uint32_t kOffBar = kOffCt + DR_Q * DR_H * 4;      // barriers
constexpr uint32_t kSmemBytes = kOffBar + 256;

enum Bar { GATE_FULL0 = 0, GATE_FULL1, GATE_FREE0, GATE_FREE1, H_READY0, H_READY1, H_READY2, H_READY3,
           X_FULL0, X_FULL1, X_FREE0, X_FREE1,
           X_LAND0, X_LAND1, W_LAND, W_READY, NUM_BARS };

__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 16 fp32 values from shared memory -> this warp's lanes x 16 TMEM columns (every lane stores the same row of constants)
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

// Training-mode outputs (kTrain): instead of reducing h into S and emitting head partials, the epilogue saves what the
// backward pass of csrc/dr_train.cu consumes, in its layout: row = (e*T + t)*B + b (B = the micro-batch of this launch),
//   rzn[dir][row][3H] = (r, z, n)    q[dir][row][H] = W_hn h_{t-1} + b_hn    hs[dir][row][H] = h_t
// with dir_stride_rows rows between the two directions.  Dropout, S and the heads are applied by the training kernels.
// lane_major != 0: rzn and q are written window-contiguous inside each (dir, e, t) block — element (column c, window b) at
//   block + ((c/4)*B + b)*4 + c%4 — so that this kernel's stores (thread = window) and the backward kernel's loads are
//   coalesced 512-byte runs per warp instead of 16 bytes per lane per row.  hs stays row-major (read by the head kernels).
struct TcTrainOut { float* rzn; float* q; float* hs; long long dir_stride_rows; int lane_major; };

// Completion signal per 256-window tile (expert-sharded forward, csrc/dr_comm.cu): the last work item of a tile to finish
// writes `value` into flag[tile] with system scope, after every thread has fenced its REDs into S and its stores into P.
// A stream memory-wait on that word (no SM involved) releases the DMA copies of this tile's partial S to the peers while
// the kernel — ONE launch for the whole batch — keeps running the next tiles.
struct TcTileSignal { unsigned int* count; unsigned int* flag; unsigned int value; };

template <bool kTiming, bool kTrain>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
dr_gru_tc_kernel(const uint8_t* __restrict__ wtc,     // [M_loc][2 dir][2 cta][kWBytes]
                 const uint8_t* __restrict__ xtc,     // [T][ntiles][2 cta][hi|lo][kXTile]
                 const float* __restrict__ bias4,     // [M_loc][2][4][H]
                 const float* __restrict__ ct,        // [M_loc][2][Q][H]
                 float* __restrict__ S,               // [T][64][Bp][4]
                 float* __restrict__ P,               // own-expert head partials [T][Bp/128][ceil(3M_loc/16)][dir*2+half][16][128]
                 int B, int T, int Bp, int M_loc, int ntiles,
                 unsigned long long* __restrict__ dbg /* nullable: cycle breakdown of work item 0 */,
                 TcTrainOut tr /* used only when kTrain */,
                 int xdrop /* 0: all three split terms of the x-part; 1 / 2: drop hi*lo / lo*hi (precision probe only) */,
                 TcTileSignal sig /* nullable counters: publish "all work items of this 256-window tile are done" */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t cta = cluster_ctarank();
    // Work items are ordered TILE-major: clusters that run at the same time work on the same 256 windows, so the slice
    // of S they all reduce into (75 MB at config 2) stays L2-resident while it accumulates instead of being written
    // back and re-fetched by every wave (ncu r01d: 27 GB of DRAM traffic per launch with the expert-major order).
    const int item = blockIdx.x >> 1;                 // (tile, e, dir)
    const int tile = item / (2 * M_loc);
    const int e = (item % (2 * M_loc)) >> 1;
    const int dir = item & 1;

    float* bs = reinterpret_cast<float*>(smem + kOffBias);
    float* cs = reinterpret_cast<float*>(smem + kOffCt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    // gate constants, pre-scaled for ex2:  [0] -(b_ir+b_hr)*log2e   [1] -(b_iz+b_hz)*log2e   [2] 2*log2e*b_in   [3] b_hn
    for (int i = tid; i < 4 * DR_H; i += kThreads) {
        float v = bias4[(size_t)(e * 2 + dir) * 4 * DR_H + i];
        int c = i / DR_H;
        bs[i] = (c < 2) ? -v * kLog2e : (c == 2) ? 2.0f * kLog2e * v : v;
    }
    for (int i = tid; i < DR_Q * DR_H; i += kThreads) cs[i] = ct[(size_t)(e * 2 + dir) * DR_Q * DR_H + i];
    if (tid == 0) {
        mbar_init(bar(GATE_FULL0), 1); mbar_init(bar(GATE_FULL1), 1);
        mbar_init(bar(GATE_FREE0), 2 * kEpiWarps); mbar_init(bar(GATE_FREE1), 2 * kEpiWarps);
        for (int i = 0; i < 4; ++i) mbar_init(bar(H_READY0 + i), 2 * kEpiWarps);
        mbar_init(bar(X_FULL0), 2); mbar_init(bar(X_FULL1), 2);
        mbar_init(bar(X_FREE0), 1); mbar_init(bar(X_FREE1), 1);
        mbar_init(bar(X_LAND0), 1); mbar_init(bar(X_LAND1), 1);
        mbar_init(bar(W_LAND), 1); mbar_init(bar(W_READY), 2);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) { tmem_alloc<2>(smem_u32(tmem_slot), 512); tmem_relinquish<2>(); }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    // Register re-partition (setmaxnreg, per warpgroup): the launch gives every thread 168 registers; the third
    // warpgroup (MMA issuer, producer, two idle warps) keeps 72 and the two epilogue warpgroups grow to 216 (2*128*216 + 128*72 = the 384*168 registers of the launch), which
    // removes the spills ptxas needed at 168 (ncu r01c: LDL long-scoreboard stalls inside the gate math).
    // (setmaxnreg is warpgroup-collective: every warp of a warpgroup executes the SAME instruction, hence before the
    //  per-warp role dispatch.)
    if (warp < kEpiWarps) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else                  asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp < kEpiWarps) {
        // ======================= gate epilogue warps =======================
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int half = warp >> 2;                               // which 16 of the quarter's 32 hidden units
        const int row = (warp & 3) * 32 + lane;                   // TMEM lane == window row in this CTA
        const int b = tile * 256 + (int)cta * 128 + row;
        const bool live = b < B;

        // ---- init: h0 = 0 (qrnn.py:39), gh_n accumulator columns = b_hn, then publish "previous step done"
        {
            uint32_t z8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
            for (int c = 0; c < 64; c += 8) tmem_st8(tbase + lane_base + kHA + half * 64 + c, z8);
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 4; ++i) mbar_arrive_remote(bar(H_READY0 + i), 0);
            }
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            // buffer g serves quarters g and g+2; the first use is quarter g
            store_bhn(tbase + lane_base + kG0 + g * 128 + 96 + half * 16, bs + 3 * DR_H + g * 32 + half * 16);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { mbar_arrive_remote(bar(GATE_FREE0), 0); mbar_arrive_remote(bar(GATE_FREE1), 0); }

        float hreg[4][16];
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 16; ++j) hreg[q][j] = 0.0f;

        const bool timing = kTiming && dbg != nullptr && item == 0 && cta == 0 && warp == 0 && lane == 0;
        long long t_wait[4] = {0, 0, 0, 0}, t_ld = 0, t_math = 0, t_tail = 0;
        const long long t_begin = clock64();
        uint32_t full_phase[2] = {0, 0};
        float o0 = 0.f, o1 = 0.f, o2 = 0.f;
        float hn[16];
        // Work that is off the recurrence's critical path — the own-expert head dot and the RED of h into S — for the 16
        // hidden units starting at u0p of time step ttp.  The two warps that share an SM sub-partition (half 0 / half 1) run
        // it at different points of the quarter so that one warp's MUFU-bound gate math overlaps the other's FMA/LSU work:
        // half 0 right after its gate math, half 1 one quarter later, just before its next gate math (clock64 breakdown:
        // in lock-step both warps contend for the 16-lane XU pipe during the math and leave it idle during this part).
        auto tail = [&](int u0p, int ttp) {
            float p0[4], p1[4], p2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { p0[i] = 0.f; p1[i] = 0.f; p2[i] = 0.f; }
#pragma unroll
            for (int j4 = 0; j4 < 16; j4 += 4) {
                const float4 c0 = *reinterpret_cast<const float4*>(cs + u0p + j4);
                const float4 c1 = *reinterpret_cast<const float4*>(cs + DR_H + u0p + j4);
                const float4 c2 = *reinterpret_cast<const float4*>(cs + 2 * DR_H + u0p + j4);
                p0[0] = fmaf(c0.x, hn[j4], p0[0]); p0[1] = fmaf(c0.y, hn[j4 + 1], p0[1]); p0[2] = fmaf(c0.z, hn[j4 + 2], p0[2]); p0[3] = fmaf(c0.w, hn[j4 + 3], p0[3]);
                p1[0] = fmaf(c1.x, hn[j4], p1[0]); p1[1] = fmaf(c1.y, hn[j4 + 1], p1[1]); p1[2] = fmaf(c1.z, hn[j4 + 2], p1[2]); p1[3] = fmaf(c1.w, hn[j4 + 3], p1[3]);
                p2[0] = fmaf(c2.x, hn[j4], p2[0]); p2[1] = fmaf(c2.y, hn[j4 + 1], p2[1]); p2[2] = fmaf(c2.z, hn[j4 + 2], p2[2]); p2[3] = fmaf(c2.w, hn[j4 + 3], p2[3]);
            }
            o0 += (p0[0] + p0[1]) + (p0[2] + p0[3]);
            o1 += (p1[0] + p1[1]) + (p1[2] + p1[3]);
            o2 += (p2[0] + p2[1]) + (p2[2] + p2[3]);
            if (live) {
                float* sp = S + (((size_t)ttp * 64 + dir * 32 + u0p / 4) * Bp + b) * 4;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    dr_red_add_v4(sp + (size_t)j * Bp * 4, hn[4 * j], hn[4 * j + 1], hn[4 * j + 2], hn[4 * j + 3]);
            }
        };
        // the step's own-expert head term of this warp (its 64 hidden units) is complete: one plain, coalesced store per
        // quantile into P (written exactly once per element; K2 sums the 2 directions x 2 halves).  Replaces 12-byte REDs
        // scattered into out[B,T,M,Q], which cost a DRAM read-modify-write per touch.
        auto flush = [&](int ttp) {
            if (live) {
                // P is laid out [t][128-window tile][16-column group][dir*2+half][16][128]: what the head kernel's CTA (t, tile)
                // needs for 16 output columns is one contiguous 32 KB slab (one bulk copy), and a warp stores 128 contiguous bytes
                const int ngrp = (M_loc * DR_Q + 15) >> 4;
                const size_t slab = (((size_t)ttp * (Bp >> 7) + (b >> 7)) * ngrp) * 4 * 16 * 128;
                const int dh = dir * 2 + half;
                const int c = e * DR_Q;
                float* o = P + slab + (b & 127);
                o[((size_t)((c >> 4) * 4 + dh) * 16 + (c & 15)) * 128] = o0;
                o[((size_t)(((c + 1) >> 4) * 4 + dh) * 16 + ((c + 1) & 15)) * 128] = o1;
                o[((size_t)(((c + 2) >> 4) * 4 + dh) * 16 + ((c + 2) & 15)) * 128] = o2;
            }
            o0 = 0.f; o1 = 0.f; o2 = 0.f;
        };
        for (int s = 0; s < T; ++s) {
            const int tt = dir ? (T - 1 - s) : s;
            const int tt_prev = dir ? (T - s) : (s - 1);
            const uint32_t hnext = (s & 1) ? kHA : kHB;           // step s reads (s&1 ? HB : HA), writes the other
            // training mode: this thread's row of the saved activations for step tt
            const size_t trow = kTrain ? (size_t)dir * (size_t)tr.dir_stride_rows + ((size_t)e * T + tt) * (size_t)B + (size_t)(live ? b : 0) : 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int buf = q & 1;
                const int u0 = q * 32 + half * 16;                // first hidden unit handled here
                const uint32_t G = tbase + lane_base + kG0 + buf * 128 + half * 16;
                const long long tq0 = timing ? clock64() : 0;
                mbar_wait(bar(GATE_FULL0 + buf), full_phase[buf]);
                full_phase[buf] ^= 1;
                tc_fence_after();
                const long long tq1 = timing ? clock64() : 0;
                uint32_t gi[16], gr[16], gz[16], gh[16];
                tmem_ld16(G + 0, gi); tmem_ld16(G + 32, gr); tmem_ld16(G + 64, gz); tmem_ld16(G + 96, gh);
                tc_wait_ld();
                {   // re-arm the gh_n accumulator columns with b_hn of the quarter that uses this buffer next
                    // (the h-part MMAs always accumulate), then hand the buffer back to the MMA warp
                    const int qn = (q + 2) & 3;
                    store_bhn(G + 96, bs + 3 * DR_H + qn * 32 + half * 16);
                    tc_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar(GATE_FREE0 + buf), 0);
                }
                const long long tq2 = timing ? clock64() : 0;
                if (!kTrain && half == 1 && (s > 0 || q > 0)) {   // lagging warp: previous quarter's tail first
                    tail(((q + 3) & 3) * 32 + 16, q == 0 ? tt_prev : tt);
                    if (q == 0) flush(tt_prev);
                }
                uint32_t phi[8], plo[8];
                // Stage-major over 8 cells at a time: each stage's instructions are independent of one another, so one warp
                // keeps the XU (MUFU) pipe and the FMA pipe busy by itself (ncu/clock64: the per-cell order was latency bound).
#pragma unroll
                for (int j8 = 0; j8 < 16; j8 += 8) {
                    float cr[8], cz[8], cn[8];
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const float4 a = *reinterpret_cast<const float4*>(bs + u0 + j8 + 4 * v);
                        const float4 b4 = *reinterpret_cast<const float4*>(bs + DR_H + u0 + j8 + 4 * v);
                        const float4 c = *reinterpret_cast<const float4*>(bs + 2 * DR_H + u0 + j8 + 4 * v);
                        cr[4 * v] = a.x; cr[4 * v + 1] = a.y; cr[4 * v + 2] = a.z; cr[4 * v + 3] = a.w;
                        cz[4 * v] = b4.x; cz[4 * v + 1] = b4.y; cz[4 * v + 2] = b4.z; cz[4 * v + 3] = b4.w;
                        cn[4 * v] = c.x; cn[4 * v + 1] = c.y; cn[4 * v + 2] = c.z; cn[4 * v + 3] = c.w;
                    }
                    float er[8], ez[8], rr[8], zz[8], pn[8];
                    // r, z = sigmoid(.): exp via ex2, and ONE reciprocal per two cells: 1/((1+er0)(1+ez0)(1+er1)(1+ez1)).
                    // Arguments are clamped at 2^30 so that four factors cannot overflow (sigmoid error < 1e-9 there).
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
                    // n = tanh(gi_n + b_in + r*(gh_n + b_hn)) = 1 - 2/(1 + exp(2t));  b_hn is already in gh
#pragma unroll
                    float rv[8], nv[8];                                       // kept only by the training variant
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float inv = rr[i ^ 1] * iv2[i >> 1];          // 1/((1+er_i)(1+ez_i))
                        zz[i] = er[i] * inv;
                        const float r_i = ez[i] * inv;
                        rv[i] = r_i;
                        const float t = fmaf(r_i, __uint_as_float(gh[j8 + i]), __uint_as_float(gi[j8 + i]));
                        pn[i] = fminf(fmaf(t, 2.0f * kLog2e, cn[i]), 30.0f);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) pn[i] = ex2_approx(pn[i]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) pn[i] += 1.0f;
                    float ivn[4];                                             // 1/(pn[2i]*pn[2i+1]) from ONE reciprocal per four cells
#pragma unroll
                    for (int g4 = 0; g4 < 2; ++g4) {
                        const float pa = pn[4 * g4] * pn[4 * g4 + 1], pb = pn[4 * g4 + 2] * pn[4 * g4 + 3];
                        const float inv4 = rcp_approx(pa * pb);
                        ivn[2 * g4] = pb * inv4;
                        ivn[2 * g4 + 1] = pa * inv4;
                    }
#pragma unroll
                    for (int i = 0; i < 8; i += 2) {
                        const int j = j8 + i;
                        const float n0 = fmaf(-2.0f * pn[i + 1], ivn[i >> 1], 1.0f);
                        const float n1 = fmaf(-2.0f * pn[i], ivn[i >> 1], 1.0f);
                        // h' = (1-z)*n + z*h, evaluated as torch's CPU cell does: (h - n)*z + n
                        const float a0 = __fadd_rn(__fmul_rn(__fsub_rn(hreg[q][j], n0), zz[i]), n0);
                        const float a1 = __fadd_rn(__fmul_rn(__fsub_rn(hreg[q][j + 1], n1), zz[i + 1]), n1);
                        hreg[q][j] = a0; hreg[q][j + 1] = a1;
                        hn[j] = a0; hn[j + 1] = a1;
                        nv[i] = n0; nv[i + 1] = n1;
                        // fp16 split of the pair with packed conversions (F2FP / HADD2.F32: no XU-pipe traffic)
                        __half2 hi2 = __floats2half2_rn(a0, a1);
                        float2 back = __half22float2(hi2);
                        __half2 lo2 = __floats2half2_rn(a0 - back.x, a1 - back.y);
                        phi[j >> 1] = *reinterpret_cast<uint32_t*>(&hi2);
                        plo[j >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
                    }
                    if (kTrain && live) {
                        float* ph = tr.hs + trow * DR_H + u0 + j8;
                        // (column c, this window): row-major  trow*ncols + c   |   lane-major  block + ((c/4)*B + b)*4
                        const size_t blk = trow - (size_t)b;                     // first row of the (dir, e, t) block
                        const size_t cs = tr.lane_major ? (size_t)B : 1;        // float4 slots between consecutive column groups
                        float* pr = tr.lane_major ? tr.rzn + blk * (3 * DR_H) + (size_t)b * 4 : tr.rzn + trow * (3 * DR_H);
                        float* pq = tr.lane_major ? tr.q + blk * DR_H + (size_t)b * 4 : tr.q + trow * DR_H;
#pragma unroll
                        for (int v = 0; v < 8; v += 4) {
                            const size_t c4 = (size_t)((u0 + j8 + v) >> 2);
                            *reinterpret_cast<float4*>(pr + c4 * cs * 4) = make_float4(rv[v], rv[v + 1], rv[v + 2], rv[v + 3]);
                            *reinterpret_cast<float4*>(pr + (c4 + DR_H / 4) * cs * 4) = make_float4(zz[v], zz[v + 1], zz[v + 2], zz[v + 3]);
                            *reinterpret_cast<float4*>(pr + (c4 + 2 * DR_H / 4) * cs * 4) = make_float4(nv[v], nv[v + 1], nv[v + 2], nv[v + 3]);
                            *reinterpret_cast<float4*>(pq + c4 * cs * 4) = make_float4(__uint_as_float(gh[j8 + v]), __uint_as_float(gh[j8 + v + 1]),
                                                                                       __uint_as_float(gh[j8 + v + 2]), __uint_as_float(gh[j8 + v + 3]));
                            *reinterpret_cast<float4*>(ph + v) = make_float4(hn[j8 + v], hn[j8 + v + 1], hn[j8 + v + 2], hn[j8 + v + 3]);
                        }
                    }
                }
                tmem_st8(tbase + lane_base + hnext + u0 / 2, phi);
                tmem_st8(tbase + lane_base + hnext + 64 + u0 / 2, plo);
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(bar(H_READY0 + q), 0);   // K columns [32q, 32q+32) of h_t are in TMEM
                const long long tq3 = timing ? clock64() : 0;
                if (!kTrain && half == 0) { tail(u0, tt); if (q == 3) flush(tt); }
                if (timing) {
                    const long long tq4 = clock64();
                    t_wait[q] += tq1 - tq0; t_ld += tq2 - tq1; t_math += tq3 - tq2; t_tail += tq4 - tq3;
                }
            }
        }
        if (!kTrain && half == 1) { const int tl = dir ? 0 : (T - 1); tail(3 * 32 + 16, tl); flush(tl); }
        if (timing) {
            dbg[0] = (unsigned long long)(clock64() - t_begin);
            for (int q = 0; q < 4; ++q) dbg[1 + q] = (unsigned long long)t_wait[q];
            dbg[5] = (unsigned long long)t_ld; dbg[6] = (unsigned long long)t_math; dbg[7] = (unsigned long long)t_tail;
        }
    } else if (warp == kMmaWarp) {
        // ======================= MMA issuer (leader CTA only) =======================
        if (cta == 0 && elect_one()) {
            const uint32_t idesc = make_idesc_f16(256, 96);
            const uint32_t w_s = smem_u32(smem + kOffW);
            const uint32_t x_s = smem_u32(smem + kOffX);
            mbar_wait_cluster(bar(W_READY), 0);
            const bool mt = kTiming && dbg != nullptr && item == 0;
            long long m_x = 0, m_free[4] = {0, 0, 0, 0}, m_h[4] = {0, 0, 0, 0};
            const long long m_begin = clock64();
            uint32_t free_bits = 0;                    // bit b = phase parity of GATE_FREE[b]
            for (int s = 0; s < T; ++s) {
                const uint32_t hcur = tbase + ((s & 1) ? kHB : kHA);
                const uint32_t xst = x_s + (s & 1) * kXStage;
                { const long long a = mt ? clock64() : 0;
                  mbar_wait_cluster(bar(X_FULL0 + (s & 1)), (s >> 1) & 1);
                  if (mt) m_x += clock64() - a; }
                const uint64_t xdesc = make_desc_sw128(xst);
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {
                    const int buf = q & 1;
                    const uint32_t G = tbase + kG0 + buf * 128;
                    const uint32_t wq = w_s + q * kQuarterBytes;
                    // descriptors: one base per operand tile; every other one is base + (byte offset >> 4) in the 14-bit
                    // start-address field (shared memory < 256 KB, so the field cannot carry out)
                    const uint64_t wdesc = make_desc_sw128(wq);
                    { const long long a = mt ? clock64() : 0;
                      mbar_wait_cluster(bar(GATE_FREE0 + buf), (free_bits >> buf) & 1u);
                      if (mt) m_free[q] += clock64() - a; }
                    free_bits ^= 1u << buf;
                    tc_fence_after();
                    // x-part: (hi,hi) (hi,lo) (lo,hi);  A parts at xst + {0, kXTile}, B parts at wq + {0, kBlk}
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        if (xdrop != 0 && term == xdrop) continue;  // xdrop is 0 (keep all), 1 or 2
                        const uint32_t ab = xst + (term == 2 ? kXTile : 0);
                        const uint32_t bb = wq + (term == 1 ? kBlk : 0);
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16)
                            mma_ss<2>(G, xdesc + ((ab - xst + k16 * 32) >> 4), wdesc + ((bb - wq + k16 * 32) >> 4), idesc,
                                      (term | k16) ? 1u : 0u);
                    }
                    // h-part: A from TMEM (hi at hcur, lo at hcur+64), B blocks Wh_hi[kb] at wq+2*kBlk, Wh_lo[kb] at wq+4*kBlk.
                    // K is walked in quarters of 32: quarter kq of h_{t-1} is published by the epilogue as soon as
                    // hidden quarter kq is done, so only the last 6 MMAs of the step's first quarter wait for the
                    // end of the previous step's epilogue.
#pragma unroll
                    for (int kq = 0; kq < 4; ++kq) {
                        if (q == 0) {
                            const long long a = mt ? clock64() : 0;
                            mbar_wait_cluster(bar(H_READY0 + kq), s & 1);
                            if (mt) m_h[kq] += clock64() - a;
                            tc_fence_after();
                        }
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t at = hcur + (term == 2 ? 64 : 0);
                            const uint32_t bb = wq + (term == 1 ? 4 : 2) * kBlk;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int ks = kq * 2 + j, kb = ks >> 2, k16 = ks & 3;
                                mma_ts<2>(G + 32, at + (kb * 64 + k16 * 16) / 2,
                                          wdesc + ((bb - wq + kb * kBlk + k16 * 32) >> 4), idesc, 1u);
                            }
                        }
                    }
                    mma_commit_2(bar(GATE_FULL0 + buf), 0x3);
                }
                mma_commit_2(bar(X_FREE0 + (s & 1)), 0x3);
            }
            if (mt) {
                dbg[8] = (unsigned long long)(clock64() - m_begin); dbg[9] = (unsigned long long)m_x;
                for (int q = 0; q < 4; ++q) { dbg[10 + q] = (unsigned long long)m_free[q]; dbg[14 + q] = (unsigned long long)m_h[q]; }
            }
        }
        __syncwarp();
    } else {
        // ======================= bulk-copy producer (warp 9; warps 10-11 only donate their registers) =======================
        if (warp == kLoadWarp && elect_one()) {
            const uint8_t* wsrc = wtc + ((size_t)(e * 2 + dir) * 2 + cta) * kWBytes;
            mbar_expect_tx(bar(W_LAND), kWBytes);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                bulk_g2s(smem_u32(smem + kOffW) + i * kQuarterBytes, wsrc + (size_t)i * kQuarterBytes, kQuarterBytes, bar(W_LAND));
            bool w_pending = true;
            for (int s = 0; s < T; ++s) {
                const int st = s & 1;
                const int tt = dir ? (T - 1 - s) : s;
                if (s >= 2) mbar_wait(bar(X_FREE0 + st), ((s >> 1) - 1) & 1);
                const uint8_t* xsrc = xtc + (((size_t)tt * ntiles + tile) * 2 + cta) * kXStage;
                mbar_expect_tx(bar(X_LAND0 + st), kXStage);
                bulk_g2s(smem_u32(smem + kOffX) + st * kXStage, xsrc, kXStage, bar(X_LAND0 + st));
                if (w_pending) { mbar_wait(bar(W_LAND), 0); mbar_arrive_cluster(bar(W_READY), 0); w_pending = false; }
                mbar_wait(bar(X_LAND0 + st), (s >> 1) & 1);
                mbar_arrive_cluster(bar(X_FULL0 + st), 0);
            }
        }
        __syncwarp();
    }

    if (sig.count != nullptr) __threadfence();           // this thread's REDs / stores are performed before the item is counted
    tc_fence_before();
    cluster_sync_all();
    if (warp == kMmaWarp) tmem_dealloc<2>(tbase, 512);
    if (sig.count != nullptr && cta == 0 && tid == 0) {
        const unsigned int done = atomicAdd(sig.count + tile, 1u) + 1u;
        if (done == (unsigned int)(2 * M_loc)) {
            sig.count[tile] = 0;                            // re-armed for the next launch (stream-ordered after this one)
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int*>(sig.flag + tile) = sig.value;
        }
    }
}

// ---- operand image builders -------------------------------------------------------------------

// weights: one thread per (e, d, cta, q, block-kind kk in 0..2 (Wx, Wh kb0, Wh kb1), row48, chunk8)
__global__ void dr_tc_pack_w_kernel(const float* __restrict__ blob, DrBlobOffsets off, int F,
                                    const float* __restrict__ mask, uint8_t* __restrict__ wtc, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int chunk = (int)(i % 8); size_t r = i / 8;
    int row = (int)(r % 48); r /= 48;
    int kk = (int)(r % 3); r /= 3;
    int q = (int)(r % 4); r /= 4;
    int c = (int)(r % 2); r /= 2;
    int d = (int)(r % 2); r /= 2;
    int e = (int)r;
    const float* ex = blob + (size_t)e * off.per_expert;
    int col = c * 48 + row;                 // D column inside the 96-wide MMA
    int grp = col / 32, unit = q * 32 + col % 32;
    float v[8];
    if (kk == 0) {                          // x-part rows: [gi_n | r | z]  -> torch gate index (2, 0, 1)
        int gate = (grp == 0) ? 2 : grp - 1;
        const float* w = ex + off.w_ih[d] + (size_t)(gate * DR_H + unit) * F;
        for (int j = 0; j < 8; ++j) {
            int k = chunk * 8 + j;
            // mask folded: W_ih' = W_ih diag(mask).  The folded weights are O(1e-3): scaled by 2^5 (and x by 2^-5, exact)
            // so that the fp16 lo parts of both operands stay clear of the subnormal range
            v[j] = (k < F) ? w[k] * mask[(size_t)e * F + k] * 32.0f : 0.0f;
        }
    } else {                                // h-part rows: [r | z | gh_n] -> torch gate index (0, 1, 2)
        const float* w = ex + off.w_hh[d] + (size_t)(grp * DR_H + unit) * DR_H + (kk - 1) * 64;
        for (int j = 0; j < 8; ++j) v[j] = w[chunk * 8 + j];
    }
    uint32_t hi[4], lo[4];
    for (int j = 0; j < 4; ++j) {
        __half h0, l0, h1, l1;
        split_f16(v[2 * j], h0, l0); split_f16(v[2 * j + 1], h1, l1);
        hi[j] = pack_h2(h0, h1);
        lo[j] = pack_h2(l0, l1);
    }
    uint8_t* base = wtc + ((size_t)(e * 2 + d) * 2 + c) * kWBytes + (size_t)q * kQuarterBytes;
    uint32_t hi_blk = (kk == 0) ? 0 : (kk == 1 ? 2 : 3);
    uint32_t lo_blk = (kk == 0) ? 1 : (kk == 1 ? 4 : 5);
    uint32_t o = sw128_offset(row, chunk * 8);
    *reinterpret_cast<uint4*>(base + hi_blk * kBlk + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + lo_blk * kBlk + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// x [B,T,F] fp32 -> xtc [T][ntiles][2][hi|lo][128 rows x 64 K] swizzled fp16 images (scaled by 2^-5); zero padded
__global__ void dr_tc_pack_x_kernel(const float* __restrict__ x, uint8_t* __restrict__ xtc,
                                    int B, int T, int F, int ntiles, long long xbs /* floats between window starts */) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)T * ntiles * 256 * 8;
    if (i >= total) return;
    int chunk = (int)(i % 8); size_t r = i / 8;
    int rb = (int)(r % 256); r /= 256;
    int tile = (int)(r % ntiles);
    int t = (int)(r / ntiles);
    int b = tile * 256 + rb;
    float v[8];
    for (int j = 0; j < 8; ++j) {
        int f = chunk * 8 + j;
        v[j] = (b < B && f < F) ? x[(size_t)b * xbs + (size_t)t * F + f] * 0.03125f : 0.0f;   // x * 2^-5, see dr_tc_pack_w_kernel
    }
    uint32_t hi[4], lo[4];
    for (int j = 0; j < 4; ++j) {
        __half h0, l0, h1, l1;
        split_f16(v[2 * j], h0, l0); split_f16(v[2 * j + 1], h1, l1);
        hi[j] = pack_h2(h0, h1);
        lo[j] = pack_h2(l0, l1);
    }
    int c = rb / 128, row = rb % 128;
    uint8_t* base = xtc + (((size_t)t * ntiles + tile) * 2 + c) * kXStage;
    uint32_t o = sw128_offset(row, chunk * 8);
    *reinterpret_cast<uint4*>(base + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + kXTile + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

bool dr_tc_built() { return true; }

bool dr_tc_supported(const dr_model* m, int B, int T) {
    (void)B; (void)T;
    return m->cfg.F <= 64;      // one 64-wide K block for the input projection (see header comment)
}

int dr_tc_prep_weights(dr_model* m) {
    if (m->cfg.F > 64 || m->M_loc == 0) return DR_OK;
    {
        int rc0 = dr_head_tc_prep(m);
        if (rc0 != DR_OK) return rc0;
    }
    size_t bytes = (size_t)m->M_loc * 2 * 2 * kWBytes;
    if (!m->d_wtc) {
        DR_CUDA(m, cudaMalloc((void**)&m->d_wtc, bytes));
        m->wtc_bytes = bytes;
    }
    size_t total = (size_t)m->M_loc * 2 * 2 * 4 * 3 * 48 * 8;
    unsigned blocks = (unsigned)((total + 255) / 256);
    dr_tc_pack_w_kernel<<<blocks, 256, 0, m->stream>>>(m->d_blob, m->off, m->cfg.F, m->d_mask,
                                                       reinterpret_cast<uint8_t*>(m->d_wtc), total);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_gru_tc(dr_model* m, const float* x, int B, int T, float* S, float* out_local) {
    (void)out_local;               // the head partials go to the workspace P; K2 writes out_local
    int ntiles = (B + 255) / 256;
    int Bp = (B + 127) / 128 * 128;
    {
        int rc0 = dr_reserve(m, (void**)&m->d_p, &m->p_cap, (size_t)T * (Bp / 128) * ((m->M_loc * DR_Q + 15) / 16) * 4 * 16 * 128 * sizeof(float));
        if (rc0 != DR_OK) return rc0;
    }
    size_t xbytes = (size_t)T * ntiles * 2 * kXStage;
    int rc = dr_reserve(m, &m->d_xtc, &m->xtc_cap, xbytes);
    if (rc != DR_OK) return rc;
    {
        size_t total = (size_t)T * ntiles * 256 * 8;
        unsigned blocks = (unsigned)((total + 255) / 256);
        dr_tc_pack_x_kernel<<<blocks, 256, 0, m->stream>>>(x, reinterpret_cast<uint8_t*>(m->d_xtc), B, T, m->cfg.F, ntiles,
                                                           m->x_bstride ? m->x_bstride : (long long)T * m->cfg.F);
        DR_CUDA(m, cudaGetLastError());
    }
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    int items = m->M_loc * 2 * ntiles;
    const TcTileSignal sig{m->tile_count, m->tile_flag, m->tile_value};
    cudaEvent_t* ev = dr_prof_slot(m);
    if (ev) DR_CUDA(m, cudaEventRecord(ev[0], m->stream));
    if (m->d_tc_dbg)
        dr_gru_tc_kernel<true, false><<<items * 2, kThreads, kSmemBytes, m->stream>>>(
            reinterpret_cast<const uint8_t*>(m->d_wtc), reinterpret_cast<const uint8_t*>(m->d_xtc), m->d_bias4, m->d_ct,
            S, m->d_p, B, T, Bp, m->M_loc, ntiles, m->d_tc_dbg, TcTrainOut{}, m->tc_xdrop, sig);
    else
        dr_gru_tc_kernel<false, false><<<items * 2, kThreads, kSmemBytes, m->stream>>>(
            reinterpret_cast<const uint8_t*>(m->d_wtc), reinterpret_cast<const uint8_t*>(m->d_xtc), m->d_bias4, m->d_ct,
            S, m->d_p, B, T, Bp, m->M_loc, ntiles, nullptr, TcTrainOut{}, m->tc_xdrop, sig);
    DR_CUDA(m, cudaGetLastError());
    if (ev) DR_CUDA(m, cudaEventRecord(ev[1], m->stream));
    m->launches += 2;
    return DR_OK;
}

// Training forward of one micro-batch on the tensor-core engine: the same recurrence kernel, instantiated to save the
// per-step activations (r, z, n, q, h) the backward pass needs instead of producing S and head partials (the training
// kernels apply dropout, the cross-expert sum and the heads from hs).  x points at the first window of the micro-batch.
int dr_launch_gru_tc_train(dr_model* m, const float* x, int Bm, int T, float* rzn, float* q, float* hs, long long dir_stride_rows,
                           int lane_major) {
    if (m->M_loc == 0 || Bm <= 0 || T <= 0) return DR_OK;
    const int ntiles = (Bm + 255) / 256, Bp = (Bm + 127) / 128 * 128;
    const size_t xbytes = (size_t)T * ntiles * 2 * kXStage;
    int rc = dr_reserve(m, &m->d_xtc_tr, &m->xtc_tr_cap, xbytes);      // not m->d_xtc: that one aliases a slot of the forward ring
    if (rc != DR_OK) return rc;
    {
        size_t total = (size_t)T * ntiles * 256 * 8;
        dr_tc_pack_x_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(x, reinterpret_cast<uint8_t*>(m->d_xtc_tr), Bm, T, m->cfg.F,
                                                                                    ntiles, (long long)T * m->cfg.F);
        DR_CUDA(m, cudaGetLastError());
    }
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    const int items = m->M_loc * 2 * ntiles;
    dr_gru_tc_kernel<false, true><<<items * 2, kThreads, kSmemBytes, m->stream>>>(
        reinterpret_cast<const uint8_t*>(m->d_wtc), reinterpret_cast<const uint8_t*>(m->d_xtc_tr), m->d_bias4, m->d_ct,
        nullptr, nullptr, Bm, T, Bp, m->M_loc, ntiles, nullptr, TcTrainOut{rzn, q, hs, dir_stride_rows, lane_major}, 0, TcTileSignal{nullptr, nullptr, 0u});
    DR_CUDA(m, cudaGetLastError());
    m->launches += 2;
    return DR_OK;
}
