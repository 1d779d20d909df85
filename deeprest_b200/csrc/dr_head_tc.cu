// K2 on the tensor cores — the cross-expert-mean head term as a split-fp16 tcgen05 GEMM.
//
//   out[b,t,i*Q+q] = own[b,t,i,q]  (4 partials the recurrence kernel stored in P)
//                  + sum_k S[b,t,k] * (A_i/(M-1))[q,k] + b_i[q]            (qrnn.py:46-54 folded)
//                  -> optional clamp + de-normalise (estimate.py:96,101-102)
//
// GEMM [B*T x 256] x [256 x 3*M_loc].  One CTA = 128 windows of one time step (cta_group::1, M = 128):
//   * the S tile is read once in its k-group-major layout (coalesced), split to fp16 hi/lo and written as SW128
//     K-major operand images by the 4 epilogue warps (generic stores + fence.proxy.async), one K-half (128) at a time;
//   * the weight tile images (96 columns = 32 experts per chunk, pre-split and pre-swizzled by K0) stream in by
//     1-D bulk copies, double buffered;
//   * accumulators for up to 384 columns live in TMEM while both K-halves and the 3 split terms accumulate;
//   * epilogue: tcgen05.ld, add the own-expert partials (coalesced over windows) and the bias, write 64-byte runs.
// The kernel is memory bound (P 1.8 GB + out 0.45 GB + S 0.3 GB at config 2); the FFMA version it replaces was
// FFMA bound at 2.2-3.3 ms.
#include "dr_common.cuh"
#include "dr_tc.cuh"

using namespace drtc;

namespace {

constexpr int kHThreads = 192;                      // warps 0-3: convert + epilogue, 4: MMA issuer, 5: weight producer
constexpr int kNC = 96;                             // columns per chunk (32 experts x Q)
constexpr int kGroupChunks = 4;                     // 384 accumulator columns per pass
constexpr uint32_t kABlk = 128 * 128;               // A block: 128 rows x 64 K fp16
constexpr uint32_t kAHalf = 4 * kABlk;              // hi kb0, hi kb1, lo kb0, lo kb1  (one K-half)
constexpr uint32_t kBBlk = kNC * 128;               // B block: 96 rows x 64 K fp16
constexpr uint32_t kBTile = 4 * kBBlk;              // hi kb0, hi kb1, lo kb0, lo kb1  (one chunk, one K-half)
// The mean-term weights A/(M-1) are O(1e-4): their fp16 lo part would fall into the subnormals (13 good bits in all).
// They are pre-scaled by 2^12 (exact) before the split and the accumulators are scaled back in the epilogue.
constexpr float kWScale = 4096.0f, kWUnscale = 1.0f / 4096.0f;
constexpr uint32_t kPSlab = 4 * 16 * 128 * 4;       // own-expert partials of 16 columns: [dir*2+half][16][128 rows] fp32 = 32 KB
constexpr uint32_t kHOffA = 0, kHOffB = kAHalf, kHOffP = kHOffB + 2 * kBTile, kHOffBar = kHOffP + 2 * kPSlab;
constexpr uint32_t kHSmem = kHOffBar + 128;

// Destinations of the forecasts. world == 1: the caller's out [B,T,N].  Expert-sharded: every rank's full forecast tensor
// [Bfull,T,world*N] (peer-mapped symmetric memory) — this rank's N columns are stored into ALL of them straight from the
// epilogue, over NVLink for the peers: the all-gather of the forecasts and the layout interleave are fused into K2.
struct HeadDst {
    float* ptr[8];
    int world;          // number of destinations
    int ld;             // row length of a destination (world*N, or N)
    int col0;           // this rank's first column (rank*N, or 0)
    long long row0;     // first window of this call inside the destination (chunked callers)
};

// The cross-expert sum as up to 8 partial sums (one per rank of an expert-sharded model; one for a single GPU).  Source w
// holds the chunk's windows in the layout [T][64][rows[w]][4] starting at its window b0[w]; they are added in index order.
struct HeadSrc {
    const float* ptr[8];
    int rows[8], b0[8];
    int n;
};

enum HBar { A_READY = 0, A_FREE, B_FULL0, B_FULL1, B_EMPTY0, B_EMPTY1, D_FULL, D_FREE, P_FULL0, P_FULL1, P_EMPTY0, P_EMPTY1, H_NUM };

__global__ void __launch_bounds__(kHThreads, 1)
dr_head_tc_kernel(HeadSrc src,                        // partial sums of S, each [T][64][rows][4]
                  const uint8_t* __restrict__ wimg,   // [n_chunks][2 khalf][kBTile]
                  const float* __restrict__ hb,       // [M_loc*Q]
                  const float* __restrict__ P,        // [T][p_tiles][ceil(N/16)][4][16][128]; this call's first tile is p_tile0
                  HeadDst dst,
                  int B, int T, int p_tiles, int p_tile0, int N, int n_chunks,
                  const float* __restrict__ dn_scale, const float* __restrict__ dn_offset, float clamp_min) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b0 = blockIdx.x * 128, t = blockIdx.y;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kHOffBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + H_NUM);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };

    if (tid == 0) {
        mbar_init(bar(A_READY), 128); mbar_init(bar(A_FREE), 1);
        mbar_init(bar(B_FULL0), 1); mbar_init(bar(B_FULL1), 1);
        mbar_init(bar(B_EMPTY0), 1); mbar_init(bar(B_EMPTY1), 1);
        mbar_init(bar(D_FULL), 1); mbar_init(bar(D_FREE), 128);
        mbar_init(bar(P_FULL0), 1); mbar_init(bar(P_FULL1), 1); mbar_init(bar(P_EMPTY0), 128); mbar_init(bar(P_EMPTY1), 128);
        fence_mbar_init();
    }
    if (warp == 4) { tmem_alloc<1>(smem_u32(tmem_slot), 512); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;
    const int n_groups = (n_chunks + kGroupChunks - 1) / kGroupChunks;

    if (warp < 4) {
        // ===================== S converter + epilogue (thread == window row) =====================
        const int row = tid, b = b0 + row;
        const bool live = b < B;
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        uint32_t a_free_uses = 0, p_uses = 0;
        for (int g = 0; g < n_groups; ++g) {
            for (int kh = 0; kh < 2; ++kh) {
                if (g > 0 || kh > 0) { mbar_wait(bar(A_FREE), a_free_uses & 1); ++a_free_uses; }   // MMAs done reading the A image
                // 32 k-groups of this K-half: float4 (4 consecutive k) per k-group, coalesced over rows
#pragma unroll 4
                for (int kg2 = 0; kg2 < 16; ++kg2) {          // two k-groups -> one 16-byte fp16 chunk (8 k)
                    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
                    if (live) {
                        for (int w = 0; w < src.n; ++w) {          // fixed order: the sum is bit-identical on every rank
                            const size_t rw = (size_t)src.rows[w];
                            const float* sp = src.ptr[w] + (((size_t)t * 64 + kh * 32 + kg2 * 2) * rw + src.b0[w] + b) * 4;
                            const float4 a0 = *reinterpret_cast<const float4*>(sp);
                            const float4 a1 = *reinterpret_cast<const float4*>(sp + rw * 4);
                            v0.x += a0.x; v0.y += a0.y; v0.z += a0.z; v0.w += a0.w;
                            v1.x += a1.x; v1.y += a1.y; v1.z += a1.z; v1.w += a1.w;
                        }
                    }
                    const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __half h0, l0, h1, l1;
                        split_f16(vv[2 * j], h0, l0); split_f16(vv[2 * j + 1], h1, l1);
                        hi[j] = pack_h2(h0, h1); lo[j] = pack_h2(l0, l1);
                    }
                    const int kb = kg2 >> 3;                   // 64-wide K block inside the half
                    const uint32_t o = sw128_offset(row, (kg2 & 7) * 8);
                    *reinterpret_cast<uint4*>(smem + kHOffA + kb * kABlk + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(smem + kHOffA + (2 + kb) * kABlk + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core (async proxy)
                mbar_arrive(bar(A_READY));
            }
            // ---- epilogue of this column group ----
            mbar_wait(bar(D_FULL), g & 1);
            tc_fence_after();
            const int c_lo = g * kGroupChunks * kNC;
            const int c_hi = min(N, c_lo + kGroupChunks * kNC);
            for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tbase + lane_base + (uint32_t)(c0 - c_lo), v);
                tc_wait_ld();
                // the 16 columns' own-expert partials arrive as one 32 KB slab by bulk copy (producer warp, 2-deep ring): with
                // 4 resident epilogue warps plain loads could not keep enough bytes in flight (ncu r01: DRAM at 7 %)
                const uint32_t pbuf = p_uses & 1;
                mbar_wait(bar(P_FULL0 + pbuf), (p_uses >> 1) & 1);
                ++p_uses;
                float pv[4][16];
                {
                    const float* ps = reinterpret_cast<const float*>(smem + kHOffP + pbuf * kPSlab) + row;
#pragma unroll
                    for (int d = 0; d < 4; ++d)
#pragma unroll
                        for (int j = 0; j < 16; ++j) pv[d][j] = ps[(d * 16 + j) * 128];
                }
                mbar_arrive(bar(P_EMPTY0 + pbuf));
                if (live) {
                    float r[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int col = c0 + j;
                        float val = 0.0f;
                        if (col < N) {
                            const int e = col / DR_Q;
                            const float own = (pv[0][j] + pv[1][j]) + (pv[2][j] + pv[3][j]);
                            val = own + __uint_as_float(v[j]) * kWUnscale + hb[col];
                            if (dn_scale) val = fmaxf(val, clamp_min) * dn_scale[e] + dn_offset[e];
                        }
                        r[j] = val;
                    }
                    const size_t off = ((size_t)(dst.row0 + b) * T + t) * dst.ld + dst.col0 + c0;
                    const bool vec = (c0 + 16 <= N) && ((dst.ld | dst.col0) & 3) == 0;
                    for (int w = 0; w < dst.world; ++w) {
                        float* o = dst.ptr[w] + off;
                        if (vec) {
#pragma unroll
                            for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
                        } else {
                            for (int j = 0; j < 16 && c0 + j < N; ++j) o[j] = r[j];
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar(D_FREE));                          // accumulators may be overwritten by the next group
        }
    } else if (warp == 4) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(128, kNC);
            const uint64_t adesc = make_desc_sw128(smem_u32(smem + kHOffA));
            uint32_t a_ready_uses = 0, bfull_uses[2] = {0, 0};
            for (int g = 0; g < n_groups; ++g) {
                if (g > 0) { mbar_wait(bar(D_FREE), (g - 1) & 1); tc_fence_after(); }
                const int cg0 = g * kGroupChunks, cg1 = min(n_chunks, cg0 + kGroupChunks);
                for (int kh = 0; kh < 2; ++kh) {
                    mbar_wait(bar(A_READY), a_ready_uses & 1); ++a_ready_uses;
                    tc_fence_after();
                    for (int c = cg0; c < cg1; ++c) {
                        const int buf = (int)((bfull_uses[0] + bfull_uses[1]) & 1);
                        mbar_wait(bar(B_FULL0 + buf), bfull_uses[buf] & 1); ++bfull_uses[buf];
                        tc_fence_after();
                        const uint64_t bdesc = make_desc_sw128(smem_u32(smem + kHOffB + buf * kBTile));
                        const uint32_t d = tbase + (uint32_t)((c - cg0) * kNC);
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t ao = (term == 2) ? 2 * kABlk : 0;        // A: hi, hi, lo
                            const uint32_t bo = (term == 1) ? 2 * kBBlk : 0;        // B: hi, lo, hi
#pragma unroll
                            for (int ks = 0; ks < 8; ++ks) {
                                const int kb = ks >> 2, k16 = ks & 3;
                                mma_ss<1>(d, adesc + ((ao + kb * kABlk + k16 * 32) >> 4), bdesc + ((bo + kb * kBBlk + k16 * 32) >> 4),
                                          idesc, (kh | term | ks) ? 1u : 0u);
                            }
                        }
                        mma_commit_1(bar(B_EMPTY0 + buf));
                    }
                    mma_commit_1(bar(A_FREE));
                }
                mma_commit_1(bar(D_FULL));
            }
        }
        __syncwarp();
    } else {
        // ===================== weight-image producer =====================
        if (elect_one()) {
            uint32_t n = 0, np = 0;
            const int ngrp16 = (N + 15) >> 4;
            const uint8_t* pcta = reinterpret_cast<const uint8_t*>(P) + ((size_t)t * p_tiles + p_tile0 + blockIdx.x) * ngrp16 * kPSlab;
            for (int g = 0; g < n_groups; ++g) {
                const int cg0 = g * kGroupChunks, cg1 = min(n_chunks, cg0 + kGroupChunks);
                for (int kh = 0; kh < 2; ++kh)
                    for (int c = cg0; c < cg1; ++c, ++n) {
                        const int buf = (int)(n & 1);
                        if (n >= 2) mbar_wait(bar(B_EMPTY0 + buf), ((n >> 1) - 1) & 1);
                        mbar_expect_tx(bar(B_FULL0 + buf), kBTile);
                        bulk_g2s(smem_u32(smem + kHOffB + buf * kBTile), wimg + ((size_t)c * 2 + kh) * kBTile, kBTile, bar(B_FULL0 + buf));
                    }
                // own-expert partial slabs of this column group, in the order the epilogue consumes them
                const int s_lo = g * kGroupChunks * kNC / 16, s_hi = (min(N, (g + 1) * kGroupChunks * kNC) + 15) / 16;
                for (int sl = s_lo; sl < s_hi; ++sl, ++np) {
                    const int buf = (int)(np & 1);
                    if (np >= 2) mbar_wait(bar(P_EMPTY0 + buf), ((np >> 1) - 1) & 1);
                    mbar_expect_tx(bar(P_FULL0 + buf), kPSlab);
                    bulk_g2s(smem_u32(smem + kHOffP + buf * kPSlab), pcta + (size_t)sl * kPSlab, kPSlab, bar(P_FULL0 + buf));
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<1>(tbase, 512);
}

// weight images: one thread per (chunk, khalf, row 0..95, 16-byte K chunk 0..15)
__global__ void dr_head_tc_pack_kernel(const float* __restrict__ abar, int N, uint8_t* __restrict__ wimg, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int c8 = (int)(i % 16); size_t r = i / 16;
    int row = (int)(r % kNC); r /= kNC;
    int kh = (int)(r % 2); int c = (int)(r / 2);
    int col = c * kNC + row;
    uint32_t hi[4], lo[4];
    for (int j = 0; j < 4; ++j) {
        float v0 = 0.f, v1 = 0.f;
        if (col < N) {
            const float* a = abar + (size_t)col * DR_2H + kh * 128 + c8 * 8 + 2 * j;
            v0 = a[0] * kWScale; v1 = a[1] * kWScale;
        }
        __half h0, l0, h1, l1;
        split_f16(v0, h0, l0); split_f16(v1, h1, l1);
        hi[j] = pack_h2(h0, h1); lo[j] = pack_h2(l0, l1);
    }
    int kb = c8 >> 3;
    uint32_t o = sw128_offset(row, (c8 & 7) * 8);
    uint8_t* base = wimg + ((size_t)c * 2 + kh) * kBTile;
    *reinterpret_cast<uint4*>(base + kb * kBBlk + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(base + (2 + kb) * kBBlk + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

}  // namespace

int dr_head_tc_prep(dr_model* m) {
    int N = m->M_loc * DR_Q;
    if (N == 0) return DR_OK;
    int n_chunks = (N + kNC - 1) / kNC;
    size_t bytes = (size_t)n_chunks * 2 * kBTile;
    if (!m->d_himg) DR_CUDA(m, cudaMalloc((void**)&m->d_himg, bytes));
    size_t total = (size_t)n_chunks * 2 * kNC * 16;
    dr_head_tc_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, m->stream>>>(m->d_abar, N, m->d_himg, total);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_heads_tc_dst(dr_model* m, const float* S, int B, int T, void* const* dst_ptrs, int n_dst, long long row0) {
    int N = m->M_loc * DR_Q;
    if (N == 0) return DR_OK;
    if (n_dst < 1 || n_dst > 8) return dr_fail(m, DR_EINVAL, "peer-write head: 1..8 destinations");
    HeadDst dst;
    for (int w = 0; w < 8; ++w) dst.ptr[w] = (w < n_dst) ? reinterpret_cast<float*>(dst_ptrs[w]) : nullptr;
    dst.world = n_dst; dst.ld = m->cfg.world * N; dst.col0 = m->cfg.rank * N; dst.row0 = row0;
    int n_chunks = (N + kNC - 1) / kNC;
    DR_CUDA(m, cudaFuncSetAttribute(dr_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmem));
    dim3 grid((B + 127) / 128, T);
    HeadSrc src{};
    src.ptr[0] = S; src.rows[0] = dr_s_rows(B); src.b0[0] = 0; src.n = 1;
    dr_head_tc_kernel<<<grid, kHThreads, kHSmem, m->stream>>>(
        src, m->d_himg, m->d_hb, m->d_p, dst, B, T, dr_s_rows(B) >> 7, 0, N, n_chunks,
        m->dn_on ? m->d_dn : nullptr, m->dn_on ? m->d_dn + m->M_loc : nullptr, m->dn_clamp);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_heads_tc(dr_model* m, const float* S, int B, int T, float* out_local) {
    int N = m->M_loc * DR_Q;
    if (N == 0) return DR_OK;
    HeadDst dst;
    for (int w = 0; w < 8; ++w) dst.ptr[w] = nullptr;
    dst.ptr[0] = out_local; dst.world = 1; dst.ld = N; dst.col0 = 0; dst.row0 = 0;
    int n_chunks = (N + kNC - 1) / kNC;
    DR_CUDA(m, cudaFuncSetAttribute(dr_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmem));
    dim3 grid((B + 127) / 128, T);
    HeadSrc src{};
    src.ptr[0] = S; src.rows[0] = dr_s_rows(B); src.b0[0] = 0; src.n = 1;
    dr_head_tc_kernel<<<grid, kHThreads, kHSmem, m->stream>>>(
        src, m->d_himg, m->d_hb, m->d_p, dst, B, T, dr_s_rows(B) >> 7, 0, N, n_chunks,
        m->dn_on ? m->d_dn : nullptr, m->dn_on ? m->d_dn + m->M_loc : nullptr, m->dn_clamp);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}

int dr_launch_heads_tc_multi(dr_model* m, const float* const* srcp, const int* src_rows, const int* src_b0, int nsrc,
                             const float* P, int p_tiles, int p_tile0, int B, int T, float* out_local) {
    int N = m->M_loc * DR_Q;
    if (N == 0) return DR_OK;
    if (nsrc < 1 || nsrc > 8) return dr_fail(m, DR_EINVAL, "head kernel: 1..8 partial sums");
    HeadDst dst;
    for (int w = 0; w < 8; ++w) dst.ptr[w] = nullptr;
    dst.ptr[0] = out_local; dst.world = 1; dst.ld = N; dst.col0 = 0; dst.row0 = 0;
    HeadSrc src{};
    for (int w = 0; w < nsrc; ++w) { src.ptr[w] = srcp[w]; src.rows[w] = src_rows[w]; src.b0[w] = src_b0[w]; }
    src.n = nsrc;
    int n_chunks = (N + kNC - 1) / kNC;
    DR_CUDA(m, cudaFuncSetAttribute(dr_head_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kHSmem));
    dim3 grid((B + 127) / 128, T);
    dr_head_tc_kernel<<<grid, kHThreads, kHSmem, m->stream>>>(
        src, m->d_himg, m->d_hb, P, dst, B, T, p_tiles, p_tile0, N, n_chunks,
        m->dn_on ? m->d_dn : nullptr, m->dn_on ? m->d_dn + m->M_loc : nullptr, m->dn_clamp);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
