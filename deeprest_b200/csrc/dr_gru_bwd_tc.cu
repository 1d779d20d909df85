// K4b (tensor-core engine) — the reverse-time recurrence of the GRU backward pass on tcgen05 / TMEM.
//
// Replaces, for one micro-batch, the 2*(T-1) launches of {dr_gate_bwd_kernel, fp32 GEMM dgh·W_hh} of csrc/dr_train.cu
// (SURVEY §8a "Backward": dh_{t-1} = dh⊙z + [da_r, da_z, dq]·W_hh) by ONE persistent kernel per direction pair.
//
// Work item = (expert, direction, 128-window tile), one CTA (cta_group::1, M = 128 windows = TMEM lanes):
//   shared memory : W_hh^T of this expert-direction as split-fp16 SW128 K-major images  B[n][k] = 8·W_hh[k][n]
//                   (n = 128 hidden units, k = 384 gate rows; hi and lo: 2 x 6 K-blocks x 16 KB = 192 KB), one bulk load.
//   TMEM          : D[128 x 128] fp32 (columns 0..127), the A operand dgh·2^ka as fp16 hi (columns 128..319) and
//                   lo (columns 320..511) — written by the epilogue warps with tcgen05.st, consumed as MMA.TS.
//   per step      : epilogue warps (thread = window row, warp/4 = hidden half) read r,z,n,q,h_prev,dh_out of step t,
//                   form the gate adjoints (same formulas as dr_gate_bwd_kernel), store (da_r, da_z, da_n, dq) for the
//                   weight-gradient GEMMs, publish dgh to TMEM; the MMA thread issues 3 x 24 MMAs (hi·hi + hi·lo + lo·hi);
//                   the epilogue adds D·2^-(ka+3) to the carried dh⊙z.  Strictly serial per step (a true dependency).
// Gradients are O(1/(M·B·T)) — far below the fp16 range — so dgh is scaled by a power of two 2^ka chosen by the host
// from 1/(M·B·T) (exact; undone on D), W_hh by 2^3 so that its lo parts stay out of the fp16 subnormals.
#include "dr_common.cuh"
#include "dr_tc.cuh"

using namespace drtc;

namespace {

constexpr int kBwdThreads = 384;                    // warps 0-7 epilogue, warp 8 MMA issuer + weight load, 9-11 register donors
constexpr uint32_t kBwBlk = 128 * 128;              // one K block of B: 128 rows (n) x 64 k (fp16) = 16 KB
constexpr uint32_t kBwPart = 6 * kBwBlk;            // hi or lo image: K = 384
constexpr uint32_t kBwImg = 2 * kBwPart;            // 192 KB per expert-direction
constexpr uint32_t kBwOffBar = kBwImg;
constexpr uint32_t kBwSmem = kBwImg + 128;
constexpr uint32_t kColD = 0, kColAhi = 128, kColAlo = 320;
constexpr float kWScale = 8.0f;
enum BwBar { BW_W_LAND = 0, BW_A_READY, BW_D_FULL, BW_NUM };

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// Input arrays are addressed as  block(dir, e, t) + (c/4)*sc + b*sb  (floats; c = column, b = window):
//   row-major   [row][ncols] : sc = 4,    sb = ncols        (what the FFMA forward produces)
//   lane-major  [c/4][b][4]  : sc = 4*B,  sb = 4            (what the tcgen05 forward and dr_dhout_kernel produce: a warp's
//                                                            32 windows read one contiguous 512-byte run)
struct BwdArgs {
    const uint8_t* wimg;      // [M_loc][2][kBwImg]
    const float* rzn;         // (r,z,n)                       block = (dir*dir_rows + (e*T+t)*B) * 3H
    const float* q;           // W_hn h + b_hn                 block = (dir*dir_rows + (e*T+t)*B) * H
    const float* dhout;       // adjoint arriving from the heads   block = (dir*dho_dir_rows + (e*T+t)*B) * H
    const float* hs;          // h_t, row-major [dir][row][H]
    float* g4;                // out, row-major [dir][row][4H] = (da_r, da_z, da_n, dq): dgi = columns 0..3H-1, dgh = 0..2H-1 | 3H..4H-1
    long long rz_sc, rz_sb, q_sc, q_sb, do_sc, do_sb;
    long long dir_rows;       // rows between the directions of rzn/q/hs/g4
    long long dho_dir_rows;   // rows between the directions of dhout
    int B, T, M_loc, ntiles;
    float a_scale;            // 2^ka
    float d_unscale;          // 2^-(ka+3)
};

__global__ void __launch_bounds__(kBwdThreads, 1) dr_gru_bwd_tc_kernel(BwdArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int item = blockIdx.x;
    const int tile = item / (2 * a.M_loc);
    const int e = (item % (2 * a.M_loc)) >> 1;
    const int dir = item & 1;
    const int T = a.T, B = a.B;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBwOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + BW_NUM);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
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

    // register re-partition as in dr_gru_tc.cu: 2 x 128 x 216 + 128 x 72 = the 384 x 168 registers of the launch
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
        float dh[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) dh[j] = 0.0f;
        uint32_t d_phase = 0;
        for (int s = T - 1; s >= 0; --s) {                      // reverse of the forward processing order
            const int t = dir ? (T - 1 - s) : s;
            const int tp = dir ? t + 1 : t - 1;                 // the step whose output was this step's h_prev
            const size_t blk = (size_t)dir * (size_t)a.dir_rows + ((size_t)e * T + t) * (size_t)B;       // first row of (dir, e, t)
            const size_t R = blk + bb;
            const size_t Rp = (size_t)dir * (size_t)a.dir_rows + ((size_t)e * T + (s > 0 ? tp : t)) * (size_t)B + bb;
            const size_t dblk = (size_t)dir * (size_t)a.dho_dir_rows + ((size_t)e * T + t) * (size_t)B;
            const float* prz = a.rzn + blk * (3 * DR_H) + bb * a.rz_sb;          // + (c/4)*rz_sc
            const float* pq = a.q + blk * DR_H + bb * a.q_sb;                    // + (c/4)*q_sc
            const float* pdo = a.dhout + dblk * DR_H + bb * a.do_sb;             // + (c/4)*do_sc
            const float* php = a.hs + Rp * DR_H + half * 64;
            float* pg = a.g4 + R * (4 * DR_H) + half * 64;
            const int c40 = half * 16;                                            // first column group (of 4) of this thread's half
            if (live && s > 0) {                                // next iteration's blocks (step tp): warm L2 while this step runs
                const long long nb = (long long)(dir ? 1 : -1) * (long long)B;  // rows to the next block
#pragma unroll
                for (int c4 = 0; c4 < 16; c4 += 8) {
#pragma unroll
                    for (int g = 0; g < 3; ++g) prefetch_l2(prz + nb * (3 * DR_H) + (long long)(g * 32 + c40 + c4) * a.rz_sc);
                    prefetch_l2(pq + nb * DR_H + (long long)(c40 + c4) * a.q_sc);
                    prefetch_l2(pdo + nb * DR_H + (long long)(c40 + c4) * a.do_sc);
                }
                if (s > 1) { prefetch_l2(php + nb * DR_H); prefetch_l2(php + nb * DR_H + 32); }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {                       // 16 hidden units at a time
                float r_[16], z_[16], n_[16], q_[16], hp[16], dv[16];
                if (live) {
#pragma unroll
                    for (int v = 0; v < 16; v += 4) {
                        const long long c4 = c40 + c * 4 + (v >> 2);            // column group inside one gate / inside q, dhout
                        const float4 x0 = __ldg(reinterpret_cast<const float4*>(prz + c4 * a.rz_sc));
                        const float4 x1 = __ldg(reinterpret_cast<const float4*>(prz + (c4 + 32) * a.rz_sc));
                        const float4 x2 = __ldg(reinterpret_cast<const float4*>(prz + (c4 + 64) * a.rz_sc));
                        const float4 x3 = __ldg(reinterpret_cast<const float4*>(pq + c4 * a.q_sc));
                        const float4 x5 = __ldg(reinterpret_cast<const float4*>(pdo + c4 * a.do_sc));
                        float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (s > 0) x4 = __ldg(reinterpret_cast<const float4*>(php + c * 16 + v));
                        r_[v] = x0.x; r_[v + 1] = x0.y; r_[v + 2] = x0.z; r_[v + 3] = x0.w;
                        z_[v] = x1.x; z_[v + 1] = x1.y; z_[v + 2] = x1.z; z_[v + 3] = x1.w;
                        n_[v] = x2.x; n_[v + 1] = x2.y; n_[v + 2] = x2.z; n_[v + 3] = x2.w;
                        q_[v] = x3.x; q_[v + 1] = x3.y; q_[v + 2] = x3.z; q_[v + 3] = x3.w;
                        hp[v] = x4.x; hp[v + 1] = x4.y; hp[v + 2] = x4.z; hp[v + 3] = x4.w;
                        dv[v] = x5.x; dv[v + 1] = x5.y; dv[v + 2] = x5.z; dv[v + 3] = x5.w;
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < 16; ++v) { r_[v] = 0.f; z_[v] = 0.f; n_[v] = 0.f; q_[v] = 0.f; hp[v] = 0.f; dv[v] = 0.f; }
                }
                float dar[16], daz[16], dan[16], dq[16];
#pragma unroll
                for (int v = 0; v < 16; ++v) {                  // dr_gate_bwd_kernel's arithmetic, operation for operation
                    const float dhv = dh[c * 16 + v] + dv[v];
                    const float dn = dhv * (1.0f - z_[v]);
                    const float dz = dhv * (hp[v] - n_[v]);
                    dan[v] = dn * (1.0f - n_[v] * n_[v]);
                    const float dr = dan[v] * q_[v];
                    dq[v] = dan[v] * r_[v];
                    daz[v] = dz * z_[v] * (1.0f - z_[v]);
                    dar[v] = dr * r_[v] * (1.0f - r_[v]);
                    dh[c * 16 + v] = dhv * z_[v];               // + dgh·W_hh below
                }
                if (live) {
#pragma unroll
                    for (int v = 0; v < 16; v += 4) {
                        *reinterpret_cast<float4*>(pg + c * 16 + v) = make_float4(dar[v], dar[v + 1], dar[v + 2], dar[v + 3]);
                        *reinterpret_cast<float4*>(pg + DR_H + c * 16 + v) = make_float4(daz[v], daz[v + 1], daz[v + 2], daz[v + 3]);
                        *reinterpret_cast<float4*>(pg + 2 * DR_H + c * 16 + v) = make_float4(dan[v], dan[v + 1], dan[v + 2], dan[v + 3]);
                        *reinterpret_cast<float4*>(pg + 3 * DR_H + c * 16 + v) = make_float4(dq[v], dq[v + 1], dq[v + 2], dq[v + 3]);
                    }
                }
                if (s > 0) {                                    // A operand of this step's product: split-fp16 dgh·2^ka into TMEM
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        const float* src = (g == 0) ? dar : (g == 1) ? daz : dq;
                        uint32_t phi[8], plo[8];
#pragma unroll
                        for (int v = 0; v < 16; v += 2) {
                            const float a0 = fminf(fmaxf(src[v] * a.a_scale, -65504.0f), 65504.0f);
                            const float a1 = fminf(fmaxf(src[v + 1] * a.a_scale, -65504.0f), 65504.0f);
                            __half2 hi2 = __floats2half2_rn(a0, a1);
                            float2 back = __half22float2(hi2);
                            __half2 lo2 = __floats2half2_rn(a0 - back.x, a1 - back.y);
                            phi[v >> 1] = *reinterpret_cast<uint32_t*>(&hi2);
                            plo[v >> 1] = *reinterpret_cast<uint32_t*>(&lo2);
                        }
                        const uint32_t col = (uint32_t)(g * DR_H + half * 64 + c * 16) / 2;
                        tmem_st8(tbase + lane_base + kColAhi + col, phi);
                        tmem_st8(tbase + lane_base + kColAlo + col, plo);
                    }
                }
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
                    for (int v = 0; v < 16; ++v) dh[c * 16 + v] = fmaf(__uint_as_float(dd[v]), a.d_unscale, dh[c * 16 + v]);
                }
                tc_fence_before();                              // D is free again once every warp arrives on A_READY
            }
        }
    } else {
        // ======================= weight load + MMA issuer (one elected thread of warp 8) =======================
        if (warp == 8 && elect_one()) {
            const uint8_t* wsrc = a.wimg + (size_t)(e * 2 + dir) * kBwImg;
            mbar_expect_tx(bar(BW_W_LAND), kBwImg);
#pragma unroll 1
            for (int i = 0; i < 12; ++i) bulk_g2s(smem_u32(smem) + i * kBwBlk, wsrc + (size_t)i * kBwBlk, kBwBlk, bar(BW_W_LAND));
            mbar_wait(bar(BW_W_LAND), 0);
            const uint32_t idesc = make_idesc_f16(128, 128);
            const uint64_t wdesc = make_desc_sw128(smem_u32(smem));
            for (int it = 0; it < T - 1; ++it) {
                mbar_wait(bar(BW_A_READY), it & 1);
                tc_fence_after();
#pragma unroll
                for (int term = 0; term < 3; ++term) {          // (hi,hi) (hi,lo) (lo,hi)
                    const uint32_t acol = tbase + (term == 2 ? kColAlo : kColAhi);
                    const uint32_t boff = (term == 1) ? kBwPart : 0;
#pragma unroll 1
                    for (int kb = 0; kb < 6; ++kb) {
#pragma unroll
                        for (int k16 = 0; k16 < 4; ++k16)
                            mma_ts<1>(tbase + kColD, acol + (kb * 64 + k16 * 16) / 2,
                                      wdesc + ((boff + kb * kBwBlk + k16 * 32) >> 4), idesc, (term | kb | k16) ? 1u : 0u);
                    }
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

// W_hh^T images: one thread per (e, d, kb, chunk8, n) -> 8 consecutive k of row n, hi and lo parts
__global__ void dr_tc_pack_whT_kernel(const float* __restrict__ blob, DrBlobOffsets off, uint8_t* __restrict__ img, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int n = (int)(i % 128); size_t r = i / 128;
    int chunk = (int)(r % 8); r /= 8;
    int kb = (int)(r % 6); r /= 6;
    int d = (int)(r % 2); r /= 2;
    int e = (int)r;
    const float* w = blob + (size_t)e * off.per_expert + off.w_hh[d];        // [3H][H]
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k0 = kb * 64 + chunk * 8 + 2 * j;
        __half h0, l0, h1, l1;
        split_f16(w[(size_t)k0 * DR_H + n] * kWScale, h0, l0);
        split_f16(w[(size_t)(k0 + 1) * DR_H + n] * kWScale, h1, l1);
        hi[j] = pack_h2(h0, h1);
        lo[j] = pack_h2(l0, l1);
    }
    uint8_t* base = img + (size_t)(e * 2 + d) * kBwImg + (size_t)kb * kBwBlk;
    const uint32_t o = sw128_offset(n, chunk * 8);
    *reinterpret_cast<uint4*>(base + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + kBwPart + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

// Backward recurrence of one micro-batch, both directions, all local experts.  in_lane_major: layout of rzn and q (see BwdArgs);
// dhout is always lane-major here (dr_dhout_kernel writes it that way for this engine); g4 is [dir][row][4H] row-major.
int dr_launch_gru_bwd_tc(dr_model* m, const float* rzn, const float* q, const float* hs, const float* dhout, float* g4,
                         long long dir_rows, long long dho_dir_rows, int Bm, int T, float inv_n, int in_lane_major) {
    const int Ml = m->M_loc;
    if (Ml == 0 || Bm <= 0 || T <= 0) return DR_OK;
    const size_t bytes = (size_t)Ml * 2 * kBwImg;
    int rc = dr_reserve(m, &m->d_whT, &m->whT_cap, bytes);
    if (rc != DR_OK) return rc;
    {
        size_t total = (size_t)Ml * 2 * 6 * 8 * 128;
        dr_tc_pack_whT_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_blob, m->off, reinterpret_cast<uint8_t*>(m->d_whT), total);
        DR_CUDA(m, cudaGetLastError());
    }
    // dgh ~ dL/dy ~ inv_n = 1/(M*B*T): scale by 2^ka with 2^ka * inv_n in [4, 8) — fp16 keeps 4 decades of head room above
    // and the hi/lo split 3 decades below before the lo part goes subnormal
    const int ka = dr_grad_scale_log2(inv_n);
    BwdArgs a;
    a.wimg = reinterpret_cast<const uint8_t*>(m->d_whT);
    a.rzn = rzn; a.q = q; a.hs = hs; a.dhout = dhout; a.g4 = g4;
    a.rz_sc = in_lane_major ? 4LL * Bm : 4; a.rz_sb = in_lane_major ? 4 : 3 * DR_H;
    a.q_sc = in_lane_major ? 4LL * Bm : 4;  a.q_sb = in_lane_major ? 4 : DR_H;
    a.do_sc = 4LL * Bm; a.do_sb = 4;
    a.dir_rows = dir_rows; a.dho_dir_rows = dho_dir_rows;
    a.B = Bm; a.T = T; a.M_loc = Ml; a.ntiles = (Bm + 127) / 128;
    a.a_scale = ldexpf(1.0f, ka);
    a.d_unscale = ldexpf(1.0f, -(ka + 3));
    DR_CUDA(m, cudaFuncSetAttribute(dr_gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kBwSmem));
    dr_gru_bwd_tc_kernel<<<Ml * 2 * a.ntiles, kBwdThreads, kBwSmem, m->stream>>>(a);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 2;
    return DR_OK;
}
