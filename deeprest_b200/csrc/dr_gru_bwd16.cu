// K4b (bf16 training engine) — the reverse-time recurrence of the GRU backward pass on tcgen05, bf16 operands and
// bf16 activation images (dr_t16.cuh).  Same chain as dr_gru_bwd_tc.cu (SURVEY §8a "Backward":
// dh_{t-1} = dh*z + [da_r, da_z, dq] W_hh), with
//   * one tensor pass (bf16 x bf16 -> fp32) instead of three, W_hh^T resident in 96 KB of shared memory;
//   * the adjoint that enters each step formed IN the kernel — d r~ = keep/(1-p) * (Ct^T dy + G-bar) (qrnn.py:43,46-54
//     differentiated) — from 3 floats of dL/dy, the window's G-bar row (shared by all experts: L2) and the head
//     coefficients in shared memory: the [2][M][T][B][H] `dhout` tensor and its kernel do not exist here;
//   * (r, z, n, q) read from the gate image and (da_r, da_z, da_n, dq) written back IN PLACE, already in the operand
//     layout the weight-gradient GEMMs bulk-copy.
// Work item = (128-window tile, expert, direction), one CTA, thread = window (TMEM lane), warp / 4 = hidden half.
#include <cstdio>
#include <cstdlib>
#include "dr_t16.cuh"

using namespace drtc;
using namespace drt16;

namespace {

constexpr int kThreads = 384;                       // warps 0-7 epilogue, warp 8 MMA issuer + weight load, 9-11 register donors
constexpr uint32_t kWkBlk = 128 * 128;              // one K block of W_hh^T: 128 rows (n) x 64 k bf16 = 16 KB
constexpr uint32_t kWImg = 6 * kWkBlk;              // K = 384: 96 KB
constexpr uint32_t kOffCt = kWImg;                  // Q*H floats
constexpr uint32_t kOffBar = kOffCt + DR_Q * DR_H * 4;
constexpr uint32_t kSmem = kOffBar + 128;
constexpr uint32_t kColD = 0, kColA = 128;          // TMEM: D 128 fp32 columns, A = dgh as bf16 pairs (192 columns)
enum BwBar { BW_W_LAND = 0, BW_A_READY, BW_D_FULL, BW_NUM };

struct Bwd16Args {
    const uint8_t* wimg;      // [M_loc][2][kWImg]
    uint8_t* gate;            // gate images, in: (r,z,n,q)  out: (da_r,da_z,da_n,dq)
    const uint8_t* himg;      // h images
    const float* dy;          // dL/dy of the micro-batch, TIME-major [T][Bm][M_loc][Q]
    const float* gbar;        // [T*Bm][2H]  row = t*Bm + b (time-major: the 128 windows of a tile and step are contiguous)
    const float* ct;          // [M_loc][2][Q][H]
    Drop drop;
    int B, T, M_loc, ntiles, e_lo, b0, Bfull;
    unsigned long long* dbg;   // nullable: clock64 breakdown of work item 0 (DR_BWD16_DBG=1, measurement hook)
};

__global__ void __launch_bounds__(kThreads, 1) dr_gru_bwd16_kernel(Bwd16Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x;
    const int tile = item / (2 * a.M_loc);
    const int e = (item % (2 * a.M_loc)) >> 1;
    const int dir = item & 1;
    const int T = a.T, B = a.B;

    float* cs = reinterpret_cast<float*>(smem + kOffCt);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BW_NUM);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    for (int i = tid; i < DR_Q * DR_H; i += kThreads) cs[i] = a.ct[(size_t)(e * 2 + dir) * DR_Q * DR_H + i];
    if (tid == 0) {
        mbar_init(bar(BW_W_LAND), 1);
        mbar_init(bar(BW_A_READY), 8);
        mbar_init(bar(BW_D_FULL), 1);
        fence_mbar_init();
    }
    if (warp == 8) { tmem_alloc<1>(smem_u32(tmem_slot), 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp < 8) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    else          asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp < 8) {
        // ======================= epilogue warps: gate adjoints, dh carry =======================
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;
        const int b = tile * 128 + row;
        const bool live = b < B;
        const size_t bb = (size_t)(live ? b : 0);
        const size_t drop_base = (((size_t)(a.e_lo + e) * a.Bfull + (size_t)a.b0 + bb) * T) * DR_2H + (size_t)dir * DR_H + half * 64;
        float dh[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) dh[j] = 0.0f;
        uint32_t d_phase = 0;
        const bool timing = a.dbg != nullptr && item == 0 && warp == 0 && lane == 0;
        long long t_work = 0, t_wait = 0, t_dread = 0, t_ld = 0, t_gb = 0, t_math = 0, t_st = 0;
        const long long t_begin = clock64();
        for (int s = T - 1; s >= 0; --s) {                      // reverse of the forward processing order
            const long long ts0 = timing ? clock64() : 0;
            const int t = dir ? (T - 1 - s) : s;
            const int tp = dir ? t + 1 : t - 1;                 // the step whose output was this step's h_prev
            uint8_t* gimg = a.gate + blk_index(dir, e, t, tile, a.M_loc, T, a.ntiles) * kGateImg + (size_t)half * kColBlk;
            const uint8_t* hpim = a.himg + blk_index(dir, e, s > 0 ? tp : t, tile, a.M_loc, T, a.ntiles) * kHImg + (size_t)half * kColBlk;
            const float* gb = a.gbar + ((size_t)t * B + bb) * DR_2H + dir * DR_H + half * 64;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            if (live) {
                const float* dd = a.dy + (((size_t)t * B + bb) * a.M_loc + e) * DR_Q;
                d0 = dd[0]; d1 = dd[1]; d2 = dd[2];
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {                       // 16 hidden units at a time = two 16-byte chunks per array
                const long long tg0 = timing ? clock64() : 0;
                uint32_t wr[8], wz[8], wn[8], wq[8], wh[8];
                ld_cols16(gimg, row, c, wr);                              // one 256-bit load per array: both halves of the sector
                ld_cols16(gimg + 2 * kColBlk, row, c, wz);
                ld_cols16(gimg + 4 * kColBlk, row, c, wn);
                ld_cols16(gimg + 6 * kColBlk, row, c, wq);
                if (s > 0) ld_cols16(hpim, row, c, wh);
                else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) wh[i] = 0u;
                }
                uint32_t gw[16];                                          // G-bar of these 16 units: issued with the image loads (one wait)
#pragma unroll
                for (int i = 0; i < 16; ++i) gw[i] = 0u;
                if (live) {
                    uint32_t g0[8], g1[8];
                    ld256(gb + c * 16, g0);
                    ld256(gb + c * 16 + 8, g1);
#pragma unroll
                    for (int i = 0; i < 8; ++i) { gw[i] = g0[i]; gw[8 + i] = g1[i]; }
                }
                long long tg1 = 0;
                if (timing) { asm volatile("" :: "r"(wr[0]), "r"(wz[0]), "r"(wn[0]), "r"(wq[0]), "r"(wh[7])); tg1 = clock64(); }
                // adjoint arriving from the heads for these 16 units: keep/(1-p) * (Ct^T dy + G-bar)
                float dv[16];
                {
                    const uint32_t kb = live ? keep16(a.drop, drop_base + (size_t)t * DR_2H + c * 16) : 0u;
#pragma unroll
                    for (int v = 0; v < 16; v += 4) {
                        const float4 g4 = make_float4(__uint_as_float(gw[v]), __uint_as_float(gw[v + 1]), __uint_as_float(gw[v + 2]), __uint_as_float(gw[v + 3]));
                        const int u = half * 64 + c * 16 + v;
                        const float4 c0 = *reinterpret_cast<const float4*>(cs + u);
                        const float4 c1 = *reinterpret_cast<const float4*>(cs + DR_H + u);
                        const float4 c2 = *reinterpret_cast<const float4*>(cs + 2 * DR_H + u);
                        const float s0 = (c0.x * d0 + c1.x * d1 + c2.x * d2 + g4.x) * a.drop.inv_keep;
                        const float s1 = (c0.y * d0 + c1.y * d1 + c2.y * d2 + g4.y) * a.drop.inv_keep;
                        const float s2 = (c0.z * d0 + c1.z * d1 + c2.z * d2 + g4.z) * a.drop.inv_keep;
                        const float s3 = (c0.w * d0 + c1.w * d1 + c2.w * d2 + g4.w) * a.drop.inv_keep;
                        dv[v] = ((kb >> v) & 1u) ? s0 : 0.0f;
                        dv[v + 1] = ((kb >> (v + 1)) & 1u) ? s1 : 0.0f;
                        dv[v + 2] = ((kb >> (v + 2)) & 1u) ? s2 : 0.0f;
                        dv[v + 3] = ((kb >> (v + 3)) & 1u) ? s3 : 0.0f;
                    }
                }
                long long tg2 = 0;
                if (timing) { asm volatile("" :: "f"(dv[0]), "f"(dv[15])); tg2 = clock64(); }
                uint32_t par[8], paz[8], pan[8], pdq[8];
#pragma unroll
                for (int v = 0; v < 16; v += 2) {               // dr_gate_bwd_kernel's arithmetic on the bf16-stored activations
                    const float2 r2 = unpack_bf2(wr[v >> 1]), z2 = unpack_bf2(wz[v >> 1]), n2 = unpack_bf2(wn[v >> 1]);
                    const float2 q2 = unpack_bf2(wq[v >> 1]), h2 = unpack_bf2(wh[v >> 1]);
                    float out4[2][4];
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const float rr = w ? r2.y : r2.x, zz = w ? z2.y : z2.x, nn = w ? n2.y : n2.x, qq = w ? q2.y : q2.x, hp = w ? h2.y : h2.x;
                        const float dhv = dh[c * 16 + v + w] + dv[v + w];
                        const float dn = dhv * (1.0f - zz);
                        const float dz = dhv * (hp - nn);
                        const float dan = dn * (1.0f - nn * nn);
                        const float dr = dan * qq;
                        out4[w][0] = dr * rr * (1.0f - rr);     // da_r
                        out4[w][1] = dz * zz * (1.0f - zz);     // da_z
                        out4[w][2] = dan;                       // da_n
                        out4[w][3] = dan * rr;                  // dq
                        dh[c * 16 + v + w] = dhv * zz;          // + dgh W_hh below
                    }
                    par[v >> 1] = pack_bf2(out4[0][0], out4[1][0]);
                    paz[v >> 1] = pack_bf2(out4[0][1], out4[1][1]);
                    pan[v >> 1] = pack_bf2(out4[0][2], out4[1][2]);
                    pdq[v >> 1] = pack_bf2(out4[0][3], out4[1][3]);
                }
                // adjoints back into the gate image, in place (dead rows stay zero: every input of theirs is zero)
                long long tg3 = 0;
                if (timing) { asm volatile("" :: "r"(par[0]), "r"(pdq[7])); tg3 = clock64(); }
                st_cols16(gimg, row, c, par);
                st_cols16(gimg + 2 * kColBlk, row, c, paz);
                st_cols16(gimg + 4 * kColBlk, row, c, pan);
                st_cols16(gimg + 6 * kColBlk, row, c, pdq);
                if (timing) { const long long tg4 = clock64(); t_ld += tg1 - tg0; t_gb += tg2 - tg1; t_math += tg3 - tg2; t_st += tg4 - tg3; }
                if (s > 0) {                                    // A operand of this step's product: dgh = (da_r, da_z, dq), k = gate*128 + unit
                    const uint32_t col = (uint32_t)(half * 64 + c * 16) / 2;
                    tmem_st8(tbase + lane_base + kColA + col, par);
                    tmem_st8(tbase + lane_base + kColA + 64 + col, paz);
                    tmem_st8(tbase + lane_base + kColA + 128 + col, pdq);
                }
            }
            if (s > 0) {
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BW_A_READY));
                const long long ts1 = timing ? clock64() : 0;
                mbar_wait(bar(BW_D_FULL), d_phase);
                d_phase ^= 1;
                tc_fence_after();
                const long long ts2 = timing ? clock64() : 0;
                if (timing) { t_work += ts1 - ts0; t_wait += ts2 - ts1; }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t dd[16];
                    tmem_ld16(tbase + lane_base + kColD + half * 64 + c * 16, dd);
                    tc_wait_ld();
#pragma unroll
                    for (int v = 0; v < 16; ++v) dh[c * 16 + v] += __uint_as_float(dd[v]);
                }
                tc_fence_before();                              // D is free again once every warp arrives on A_READY
                if (timing) t_dread += clock64() - ts2;
            }
        }
        if (timing) {
            a.dbg[0] = (unsigned long long)(clock64() - t_begin); a.dbg[1] = (unsigned long long)t_work;
            a.dbg[2] = (unsigned long long)t_wait; a.dbg[3] = (unsigned long long)t_dread; a.dbg[4] = (unsigned long long)T;
            a.dbg[5] = (unsigned long long)t_ld; a.dbg[6] = (unsigned long long)t_gb; a.dbg[7] = (unsigned long long)t_math; a.dbg[8] = (unsigned long long)t_st;
        }
    } else {
        // ======================= weight load + MMA issuer (one elected thread of warp 8) =======================
        if (warp == 8 && elect_one()) {
            const uint8_t* wsrc = a.wimg + (size_t)(e * 2 + dir) * kWImg;
            mbar_expect_tx(bar(BW_W_LAND), kWImg);
#pragma unroll 1
            for (int i = 0; i < 6; ++i) bulk_g2s(smem_u32(smem) + i * kWkBlk, wsrc + (size_t)i * kWkBlk, kWkBlk, bar(BW_W_LAND));
            mbar_wait(bar(BW_W_LAND), 0);
            const uint32_t idesc = make_idesc_bf16(128, 128);
            const uint64_t wdesc = make_desc_sw128(smem_u32(smem));
            // L2 prefetch of the activations two steps ahead, one bulk request per contiguous 16 KB column block (a per-row
            // `prefetch.global.L2` covers one 32-byte sector: the clock64 breakdown showed every group's loads at DRAM latency)
            auto prefetch_step = [&](int sp) {
                if (sp < 0) return;
                const int tq = dir ? (T - 1 - sp) : sp;
                const uint8_t* gq = a.gate + blk_index(dir, e, tq, tile, a.M_loc, T, a.ntiles) * kGateImg;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(gq), "r"(kGateImg) : "memory");
                if (sp > 0) {
                    const int tpq = dir ? tq + 1 : tq - 1;
                    const uint8_t* hq = a.himg + blk_index(dir, e, tpq, tile, a.M_loc, T, a.ntiles) * kHImg;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(hq), "r"(kHImg) : "memory");
                }
                const int rows = min(128, a.B - tile * 128);
                const float* gbq = a.gbar + ((size_t)tq * a.B + tile * 128) * DR_2H;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(gbq), "r"((uint32_t)(rows * DR_2H * 4)) : "memory");
            };
            prefetch_step(T - 1); prefetch_step(T - 2);
            for (int it = 0; it < T - 1; ++it) {
                prefetch_step(T - 3 - it);
                mbar_wait(bar(BW_A_READY), it & 1);
                tc_fence_after();
#pragma unroll 1
                for (int kb = 0; kb < 6; ++kb) {
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        mma_ts<1>(tbase + kColD, tbase + kColA + (kb * 64 + k16 * 16) / 2,
                                  wdesc + ((kb * kWkBlk + k16 * 32) >> 4), idesc, (kb | k16) ? 1u : 0u);
                }
                mma_commit_1(bar(BW_D_FULL));
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<1>(tbase, 512);
}

// W_hh^T image: B[n][k] = W_hh[k][n]; one thread per (e, d, kb, chunk8, n)
__global__ void dr_t16_pack_whT_kernel(const float* __restrict__ blob, DrBlobOffsets off, uint8_t* __restrict__ img, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int n = (int)(i % 128); size_t r = i / 128;
    int chunk = (int)(r % 8); r /= 8;
    int kb = (int)(r % 6); r /= 6;
    int d = (int)(r % 2); r /= 2;
    int e = (int)r;
    const float* w = blob + (size_t)e * off.per_expert + off.w_hh[d];        // [3H][H]
    uint32_t p[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k0 = kb * 64 + chunk * 8 + 2 * j;
        p[j] = pack_bf2(w[(size_t)k0 * DR_H + n], w[(size_t)(k0 + 1) * DR_H + n]);
    }
    *reinterpret_cast<uint4*>(img + (size_t)(e * 2 + d) * kWImg + (size_t)kb * kWkBlk + sw128_offset(n, chunk * 8)) =
        make_uint4(p[0], p[1], p[2], p[3]);
}

}  // namespace

size_t dr_t16_whT_bytes(int M_loc) { return (size_t)M_loc * 2 * kWImg; }

int dr_t16_pack_whT(dr_model* m, uint8_t* img) {
    if (m->M_loc == 0) return DR_OK;
    size_t total = (size_t)m->M_loc * 2 * 6 * 8 * 128;
    dr_t16_pack_whT_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_blob, m->off, img, total);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_gru_bwd16(dr_model* m, const uint8_t* whT, uint8_t* gate, const uint8_t* himg, const float* dy, const float* gbar,
                        int Bm, int T, const uint8_t* mask, uint64_t seed, int b0, int Bfull) {
    const int Ml = m->M_loc;
    if (Ml == 0 || Bm <= 0 || T <= 0) return DR_OK;
    Bwd16Args a;
    a.wimg = whT; a.gate = gate; a.himg = himg; a.dy = dy; a.gbar = gbar; a.ct = m->d_ct;
    const float p = m->cfg.dropout_p;
    a.drop.mask = mask; a.drop.seed = seed; a.drop.inv_keep = 1.0f / (1.0f - p);
    a.drop.thr16 = (uint32_t)(p * 65536.0f + 0.5f);
    a.B = Bm; a.T = T; a.M_loc = Ml; a.ntiles = (Bm + 127) / 128; a.e_lo = m->e_lo; a.b0 = b0; a.Bfull = Bfull;
    a.dbg = nullptr;
    unsigned long long* dbg = nullptr;
    if (getenv("DR_BWD16_DBG")) { cudaMalloc((void**)&dbg, 128); cudaMemset(dbg, 0, 128); a.dbg = dbg; }
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_bwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    dr_gru_bwd16_kernel<<<Ml * 2 * a.ntiles, kThreads, kSmem, m->stream>>>(a);
    if (dbg) {                                                  // measurement hook: cycles per step of work item 0, warp 0
        unsigned long long h[16] = {0};
        cudaStreamSynchronize(m->stream);
        cudaMemcpy(h, dbg, 128, cudaMemcpyDeviceToHost);
        cudaFree(dbg);
        if (h[4]) fprintf(stderr, "[bwd16 timing, item 0 warp 0] cycles per step: total %.0f  loads+math+stores %.0f  wait(MMA + other warps) %.0f  D read %.0f\n",
                          (double)h[0] / h[4], (double)h[1] / h[4], (double)h[2] / h[4], (double)h[3] / h[4]);
        if (h[4]) fprintf(stderr, "[bwd16 timing] inside loads+math+stores, per step: image loads until first use %.0f  dy/G-bar adjoint %.0f  gate adjoints %.0f  stores issued %.0f\n",
                          (double)h[5] / h[4], (double)h[6] / h[4], (double)h[7] / h[4], (double)h[8] / h[4]);
    }
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
