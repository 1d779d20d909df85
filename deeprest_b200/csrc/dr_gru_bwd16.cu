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

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

struct Bwd16Args {
    const uint8_t* wimg;      // [M_loc][2][kWImg]
    uint8_t* gate;            // gate images, in: (r,z,n,q)  out: (da_r,da_z,da_n,dq)
    const uint8_t* himg;      // h images
    const float* dy;          // dL/dy of the micro-batch [Bm][T][M_loc][Q]
    const float* gbar;        // [Bm*T][2H]  row = b*T + t
    const float* ct;          // [M_loc][2][Q][H]
    Drop drop;
    int B, T, M_loc, ntiles, e_lo, b0, Bfull;
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
    for (int i = tid; i < DR_Q * DR_H; i += kThreads) cs[i] = a.ct[(size_t)(e * 2 + dir) * DR_Q * DR_H + i] * a.drop.inv_keep;
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
        // One step's inputs are consumed 8 hidden units at a time; the loads of the NEXT group (and, at the end of a step, of the
        // next step's first group) are issued before the current group's arithmetic, so the ~700-cycle L2 latency of each group
        // hides behind the previous one instead of stalling the warp eight times per step (ncu r02: 40 % long-scoreboard stalls).
        struct LoadSet { uint4 r, z, n, q, h; float4 g0, g1; };
        struct StepPtr { uint8_t* gimg; const uint8_t* hpim; const float* gb; const float* dd; int t; bool has_h; };
        auto step_ptr = [&](int s) {
            StepPtr p;
            p.t = dir ? (T - 1 - s) : s;
            const int tp = dir ? p.t + 1 : p.t - 1;            // the step whose output was this step's h_prev
            p.has_h = s > 0;
            p.gimg = a.gate + blk_index(dir, e, p.t, tile, a.M_loc, T, a.ntiles) * kGateImg + (size_t)half * kColBlk;
            p.hpim = a.himg + blk_index(dir, e, s > 0 ? tp : p.t, tile, a.M_loc, T, a.ntiles) * kHImg + (size_t)half * kColBlk;
            p.gb = a.gbar + (bb * T + p.t) * DR_2H + dir * DR_H + half * 64;
            p.dd = a.dy + ((bb * T + p.t) * a.M_loc + e) * DR_Q;
            return p;
        };
        auto load_set = [&](const StepPtr& p, int c8) {
            LoadSet L;
            const uint32_t o = img_off(row, c8);
            L.r = *reinterpret_cast<const uint4*>(p.gimg + o);
            L.z = *reinterpret_cast<const uint4*>(p.gimg + 2 * kColBlk + o);
            L.n = *reinterpret_cast<const uint4*>(p.gimg + 4 * kColBlk + o);
            L.q = *reinterpret_cast<const uint4*>(p.gimg + 6 * kColBlk + o);
            L.h = make_uint4(0, 0, 0, 0);
            if (p.has_h) L.h = __ldg(reinterpret_cast<const uint4*>(p.hpim + o));
            L.g0 = make_float4(0.f, 0.f, 0.f, 0.f); L.g1 = L.g0;
            if (live) { L.g0 = __ldg(reinterpret_cast<const float4*>(p.gb + c8 * 8)); L.g1 = __ldg(reinterpret_cast<const float4*>(p.gb + c8 * 8 + 4)); }
            return L;
        };
        StepPtr sp = step_ptr(T - 1);
        LoadSet cur = load_set(sp, 0);
        float d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (live) { d0 = __ldg(sp.dd); d1 = __ldg(sp.dd + 1); d2 = __ldg(sp.dd + 2); }
        for (int s = T - 1; s >= 0; --s) {                      // reverse of the forward processing order
            StepPtr nsp = sp;
            float nd0 = 0.f, nd1 = 0.f, nd2 = 0.f;
            if (s > 0) {
                nsp = step_ptr(s - 1);
                if (live) { nd0 = __ldg(nsp.dd); nd1 = __ldg(nsp.dd + 1); nd2 = __ldg(nsp.dd + 2); }
                if (s > 1) {                                    // two steps ahead: warm L2 (one 128-byte row per array)
                    const StepPtr fp = step_ptr(s - 2);
                    const uint32_t ro = img_off(row, 0) & ~127u;
#pragma unroll
                    for (int g = 0; g < 4; ++g) prefetch_l2(fp.gimg + (size_t)g * 2 * kColBlk + ro);
                    if (fp.has_h) prefetch_l2(fp.hpim + ro);
                    if (live) { prefetch_l2(fp.gb); prefetch_l2(fp.gb + 32); }
                }
            }
            uint32_t kb16 = 0;
            const f2 dd0 = mk2(d0, d0), dd1 = mk2(d1, d1), dd2 = mk2(d2, d2), ik2 = mk2(a.drop.inv_keep, a.drop.inv_keep), one2 = mk2(1.0f, 1.0f);
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {                    // 8 hidden units per group = one 16-byte chunk per array
                LoadSet nxt = cur;
                if (c8 < 7) nxt = load_set(sp, c8 + 1);
                else if (s > 0) nxt = load_set(nsp, 0);
                if ((c8 & 1) == 0) kb16 = live ? keep16(a.drop, drop_base + (size_t)sp.t * DR_2H + c8 * 8) : 0u;
                const uint32_t kb = (c8 & 1) ? (kb16 >> 8) : kb16;
                const uint32_t wr[4] = {cur.r.x, cur.r.y, cur.r.z, cur.r.w}, wz[4] = {cur.z.x, cur.z.y, cur.z.z, cur.z.w};
                const uint32_t wn[4] = {cur.n.x, cur.n.y, cur.n.z, cur.n.w}, wq[4] = {cur.q.x, cur.q.y, cur.q.z, cur.q.w};
                const uint32_t wh[4] = {cur.h.x, cur.h.y, cur.h.z, cur.h.w};
                const float gbv[8] = {cur.g0.x, cur.g0.y, cur.g0.z, cur.g0.w, cur.g1.x, cur.g1.y, cur.g1.z, cur.g1.w};
                uint32_t par[4], paz[4], pan[4], pdq[4];
#pragma unroll
                for (int v = 0; v < 8; v += 2) {                // dr_gate_bwd_kernel's arithmetic on the bf16-stored activations,
                    // two hidden units per instruction (FFMA2 / FMUL2 / FADD2)
                    const f2 r2 = unpack_bf2p(wr[v >> 1]), z2 = unpack_bf2p(wz[v >> 1]), n2 = unpack_bf2p(wn[v >> 1]);
                    const f2 q2 = unpack_bf2p(wq[v >> 1]), h2 = unpack_bf2p(wh[v >> 1]);
                    const int u = half * 64 + c8 * 8 + v;
                    // adjoint arriving from the heads: keep/(1-p) * (Ct^T dy + G-bar)   (cs is pre-scaled by 1/(1-p))
                    f2 sv = fma2(*reinterpret_cast<const f2*>(cs + u), dd0, mul2(mk2(gbv[v], gbv[v + 1]), ik2));
                    sv = fma2(*reinterpret_cast<const f2*>(cs + DR_H + u), dd1, sv);
                    sv = fma2(*reinterpret_cast<const f2*>(cs + 2 * DR_H + u), dd2, sv);
                    const f2 km = mk2(((kb >> v) & 1u) ? 1.0f : 0.0f, ((kb >> (v + 1)) & 1u) ? 1.0f : 0.0f);
                    const f2 dhv = fma2(sv, km, mk2(dh[c8 * 8 + v], dh[c8 * 8 + v + 1]));
                    const f2 omz = sub2(one2, z2);
                    const f2 dn = mul2(dhv, omz);
                    const f2 dz = mul2(dhv, sub2(h2, n2));
                    const f2 dan = mul2(dn, sub2(one2, mul2(n2, n2)));
                    const f2 dr = mul2(dan, q2);
                    const f2 dar = mul2(mul2(dr, r2), sub2(one2, r2));
                    const f2 daz = mul2(mul2(dz, z2), omz);
                    const f2 dqq = mul2(dan, r2);
                    un2(mul2(dhv, z2), dh[c8 * 8 + v], dh[c8 * 8 + v + 1]);     // + dgh W_hh below
                    par[v >> 1] = pack_bf2p(dar);
                    paz[v >> 1] = pack_bf2p(daz);
                    pan[v >> 1] = pack_bf2p(dan);
                    pdq[v >> 1] = pack_bf2p(dqq);
                }
                // adjoints back into the gate image, in place (dead rows stay zero: every input of theirs is zero)
                const uint32_t o = img_off(row, c8);
                *reinterpret_cast<uint4*>(sp.gimg + o) = make_uint4(par[0], par[1], par[2], par[3]);
                *reinterpret_cast<uint4*>(sp.gimg + 2 * kColBlk + o) = make_uint4(paz[0], paz[1], paz[2], paz[3]);
                *reinterpret_cast<uint4*>(sp.gimg + 4 * kColBlk + o) = make_uint4(pan[0], pan[1], pan[2], pan[3]);
                *reinterpret_cast<uint4*>(sp.gimg + 6 * kColBlk + o) = make_uint4(pdq[0], pdq[1], pdq[2], pdq[3]);
                if (s > 0) {                                    // A operand of this step's product: dgh = (da_r, da_z, dq), k = gate*128 + unit
                    const uint32_t col = (uint32_t)(half * 64 + c8 * 8) / 2;
                    tmem_st4(tbase + lane_base + kColA + col, par);
                    tmem_st4(tbase + lane_base + kColA + 64 + col, paz);
                    tmem_st4(tbase + lane_base + kColA + 128 + col, pdq);
                }
                cur = nxt;
            }
            if (s > 0) {
                tc_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(BW_A_READY));
                mbar_wait(bar(BW_D_FULL), d_phase);
                d_phase ^= 1;
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    uint32_t dd[16];
                    tmem_ld16(tbase + lane_base + kColD + half * 64 + c * 16, dd);
                    tc_wait_ld();
#pragma unroll
                    for (int v = 0; v < 16; ++v) dh[c * 16 + v] += __uint_as_float(dd[v]);
                }
                tc_fence_before();                              // D is free again once every warp arrives on A_READY
            }
            sp = nsp; d0 = nd0; d1 = nd1; d2 = nd2;
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
            for (int it = 0; it < T - 1; ++it) {
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
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_bwd16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
    dr_gru_bwd16_kernel<<<Ml * 2 * a.ntiles, kThreads, kSmem, m->stream>>>(a);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
