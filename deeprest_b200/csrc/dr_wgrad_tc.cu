// K4c (tensor-core engine) — weight-gradient reductions over the (t, b) rows of a micro-batch on tcgen05:
//     C[z][m][n] (+)= sum_k A[z][k][m] * B[z][k][n]        (SURVEY §8a "Backward": dW_hh += dgh (x) h_{t-1},  P = dgi^T x)
// A = gate adjoints [rows][3H] and B = h_{t-1} [rows][H] or x [rows][F] are fp32, row-major over k (the REDUCED index), i.e.
// "MN-major" for the tensor core.  8 converter warps read them coalesced along m / n, split every value into fp16 hi + lo
// (fp32 parity: hi·hi + hi·lo + lo·hi, fp32 accumulate in TMEM) and write the K-major SW128 operand images the MMA wants
// straight into shared memory — a thread gathers 8 consecutive k of one row m and issues ONE 16-byte store per part, so
// the transposition costs no extra pass over memory.  2-stage ring, one MMA-issuing thread, cta_group::1, M tile = 128.
// Operands are pre-scaled by powers of two (gradients are O(1/(M·B·T)), far below the fp16 range); undone in the epilogue.
#include "dr_common.cuh"
#include "dr_tc.cuh"

using namespace drtc;

namespace {

constexpr int kWgThreads = 288;                       // warps 0-7 converters + epilogue, warp 8 MMA issuer
constexpr int kWgConv = 256;
constexpr uint32_t kWgAPart = 128 * 128;              // A tile part: 128 rows (m) x 64 k fp16 = 16 KB
enum WgBar { WG_FULL0 = 0, WG_FULL1, WG_EMPTY0, WG_EMPTY1, WG_DFULL, WG_NUM };

struct WgArgs {                                       // blockIdx.z selects one of up to two problems (the two GRU directions)
    const float* A[2]; long long lda, bsA;
    const float* B[2]; long long ldb, bsB;
    float* C[2]; long long ldc, bsC;
    int M, N, Npad, K;
    int acol[3];              // A column of the first row of m-tile 0, 1, 2 (-1: m-tile i starts at column 128*i)
    float a_scale, b_scale, c_unscale;
    int accumulate;
};

__device__ __forceinline__ void split8_store(const float (&v)[8], float scale, uint8_t* hi_dst, uint8_t* lo_dst) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {                       // packed conversions (F2FP): no XU-pipe traffic
        const float a0 = fminf(fmaxf(v[2 * j] * scale, -65504.0f), 65504.0f);
        const float a1 = fminf(fmaxf(v[2 * j + 1] * scale, -65504.0f), 65504.0f);
        __half2 hi2 = __floats2half2_rn(a0, a1);
        const float2 back = __half22float2(hi2);
        __half2 lo2 = __floats2half2_rn(a0 - back.x, a1 - back.y);
        hi[j] = *reinterpret_cast<uint32_t*>(&hi2);
        lo[j] = *reinterpret_cast<uint32_t*>(&lo2);
    }
    *reinterpret_cast<uint4*>(hi_dst) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(lo_dst) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(kWgThreads, 1) dr_wgrad_tc_kernel(WgArgs g) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.x * 128;
    const float* __restrict__ A = (blockIdx.z ? g.A[1] : g.A[0]) + (size_t)blockIdx.y * g.bsA;
    const float* __restrict__ B = (blockIdx.z ? g.B[1] : g.B[0]) + (size_t)blockIdx.y * g.bsB;
    float* C = (blockIdx.z ? g.C[1] : g.C[0]) + (size_t)blockIdx.y * g.bsC;
    const uint32_t bpart = (uint32_t)g.Npad * 128u;                   // B tile part: Npad rows (n) x 64 k fp16
    const uint32_t stage_bytes = 2 * kWgAPart + 2 * bpart;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * stage_bytes);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + WG_NUM);
    auto bar = [&](int i) { return smem_u32(&bars[i]); };
    const int nchunks = (g.K + 63) / 64;

    if (tid == 0) {
        mbar_init(bar(WG_FULL0), kWgConv); mbar_init(bar(WG_FULL1), kWgConv);
        mbar_init(bar(WG_EMPTY0), 1); mbar_init(bar(WG_EMPTY1), 1);
        mbar_init(bar(WG_DFULL), 1);
        fence_mbar_init();
    }
    if (warp == 8) { tmem_alloc<1>(smem_u32(tmem_slot), 256); tmem_relinquish<1>(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (warp < 8) {
        // ===================== converters: fp32 [k][m] / [k][n] -> split-fp16 K-major images =====================
        const int am = tid & 127, akg0 = tid >> 7;                      // A: row m, k-groups akg0, akg0+2, akg0+4, akg0+6
        const bool a_live = m0 + am < g.M;
        // the m-tile's 128 source columns need not sit at column m0 of A (dW_hh takes (da_r, da_z, dq) out of 4H-wide rows)
        const int amap = (blockIdx.x == 0) ? g.acol[0] : (blockIdx.x == 1) ? g.acol[1] : (blockIdx.x == 2) ? g.acol[2] : -1;
        const int acol = (amap >= 0 ? amap : m0) + am;
        for (int c = 0; c < nchunks; ++c) {
            const int st = c & 1;
            if (c >= 2) mbar_wait(bar(WG_EMPTY0 + st), (uint32_t)(((c >> 1) - 1) & 1));   // MMAs of chunk c-2 have read this stage
            uint8_t* sA = smem + (size_t)st * stage_bytes;
            uint8_t* sB = sA + 2 * kWgAPart;
            const int k0 = c * 64;
            // every load of the chunk's A tile and of the first B batch is issued before the first conversion (64 independent
            // loads in flight per thread: the kernel streams ~1 KB per reduced row and is otherwise bound by memory latency)
            for (int base = 0; base < g.Npad * 8; base += 4 * kWgConv) {
                float va[4][8], vb[4][8];
                if (base == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int kg = akg0 + 2 * i;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int k = k0 + kg * 8 + j;
                            va[i][j] = (a_live && k < g.K) ? A[(size_t)k * g.lda + acol] : 0.0f;
                        }
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int task = base + i * kWgConv + tid;
                    const int n = task % g.Npad, kg = task / g.Npad;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = k0 + kg * 8 + j;
                        vb[i][j] = (task < g.Npad * 8 && n < g.N && k < g.K) ? B[(size_t)k * g.ldb + n] : 0.0f;
                    }
                }
                if (base == 0) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t o = sw128_offset(am, (akg0 + 2 * i) * 8);
                        split8_store(va[i], g.a_scale, sA + o, sA + kWgAPart + o);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int task = base + i * kWgConv + tid;
                    if (task < g.Npad * 8) {
                        const uint32_t o = sw128_offset(task % g.Npad, (task / g.Npad) * 8);
                        split8_store(vb[i], g.b_scale, sB + o, sB + bpart + o);
                    }
                }
            }
            fence_proxy_async();                                        // generic-proxy stores -> visible to the tensor core
            mbar_arrive(bar(WG_FULL0 + st));
        }
        // ===================== epilogue: D -> C =====================
        mbar_wait(bar(WG_DFULL), 0);
        tc_fence_after();
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const int m = m0 + (warp & 3) * 32 + lane;
        for (int grp = (warp >> 2); grp * 16 < g.Npad; grp += 2) {     // 16-column groups alternate between the two warp sets
            uint32_t v[16];
            tmem_ld16(tbase + lane_base + (uint32_t)(grp * 16), v);
            tc_wait_ld();
            if (m < g.M) {
                float* crow = C + (size_t)m * g.ldc + grp * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    if (grp * 16 + j < g.N) {
                        const float val = __uint_as_float(v[j]) * g.c_unscale;
                        crow[j] = g.accumulate ? crow[j] + val : val;
                    }
                }
            }
        }
    } else {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            const uint32_t idesc = make_idesc_f16(128, g.Npad);
            for (int c = 0; c < nchunks; ++c) {
                const int st = c & 1;
                mbar_wait(bar(WG_FULL0 + st), (uint32_t)((c >> 1) & 1));
                tc_fence_after();
                const uint32_t sA = smem_u32(smem) + (uint32_t)st * stage_bytes;
                const uint64_t adesc = make_desc_sw128(sA);
                const uint64_t bdesc = make_desc_sw128(sA + 2 * kWgAPart);
#pragma unroll
                for (int term = 0; term < 3; ++term) {                  // (hi,hi) (hi,lo) (lo,hi)
                    const uint32_t ao = (term == 2) ? kWgAPart : 0;
                    const uint32_t bo = (term == 1) ? bpart : 0;
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        mma_ss<1>(tbase, adesc + ((ao + k16 * 32) >> 4), bdesc + ((bo + k16 * 32) >> 4), idesc, (c | term | k16) ? 1u : 0u);
                }
                mma_commit_1(bar(WG_EMPTY0 + st));
            }
            mma_commit_1(bar(WG_DFULL));
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<1>(tbase, 256);
}

}  // namespace

bool dr_wgrad_tc_ok(int M, int N, int K) { return M >= 1 && N >= 1 && N <= 256 && K >= 1; }

// log2 of the power-of-two scale applied to gradient operands: 2^ka * inv_n in [4, 8)
int dr_grad_scale_log2(float inv_n) {
    int ex = 0;
    frexpf(inv_n, &ex);                                   // inv_n = f * 2^ex, f in [0.5, 1)
    return 3 - ex;
}

// ndir problems (1 or 2: the GRU directions) with identical shapes and strides run in one grid (blockIdx.z)
int dr_launch_wgrad_tc(dr_model* m, int ndir, const float* const* A, long long lda, long long bsA, const int* acol3 /* nullable */,
                       const float* const* B, long long ldb, long long bsB, float* const* C, long long ldc, long long bsC,
                       int M, int N, int K, int batch, int a_scale_log2, int b_scale_log2, int accumulate) {
    if (batch <= 0 || ndir < 1 || ndir > 2 || !dr_wgrad_tc_ok(M, N, K)) return DR_OK;
    WgArgs g;
    for (int d = 0; d < 2; ++d) { g.A[d] = A[d < ndir ? d : 0]; g.B[d] = B[d < ndir ? d : 0]; g.C[d] = C[d < ndir ? d : 0]; }
    g.lda = lda; g.bsA = bsA; g.ldb = ldb; g.bsB = bsB; g.ldc = ldc; g.bsC = bsC;
    g.M = M; g.N = N; g.Npad = (N + 15) / 16 * 16; g.K = K;
    for (int i = 0; i < 3; ++i) g.acol[i] = acol3 ? acol3[i] : -1;
    g.a_scale = ldexpf(1.0f, a_scale_log2); g.b_scale = ldexpf(1.0f, b_scale_log2);
    g.c_unscale = ldexpf(1.0f, -(a_scale_log2 + b_scale_log2));
    g.accumulate = accumulate;
    const size_t smem = 2 * (size_t)(2 * kWgAPart + 2 * g.Npad * 128) + 128;
    DR_CUDA(m, cudaFuncSetAttribute(dr_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * (2 * kWgAPart + 2 * 256 * 128) + 128)));
    dim3 grid((M + 127) / 128, batch, ndir);
    dr_wgrad_tc_kernel<<<grid, kWgThreads, smem, m->stream>>>(g);
    DR_CUDA(m, cudaGetLastError());
    m->launches += 1;
    return DR_OK;
}
