// Hardware probe for the tensor-core engine's building blocks (diagnostic entry point
// dr_tc_probe): one tcgen05 GEMM tile D[128*CG x N] = A[128*CG x K] * B[N x K]^T in bf16 with
//   * operands staged by 1-D bulk (TMA-engine) copies of pre-swizzled SW128 K-major images,
//   * A either from shared memory (SS) or written to TMEM with tcgen05.st (TS),
//   * cta_group::1, or cta_group::2 on a 2-CTA cluster (B split N/2 per CTA, multicast commit).
// tests/test_gpu_tc_probe.py checks every variant against numpy, so each layout / descriptor
// assumption the GRU kernel relies on is pinned separately from the kernel's own logic.
#include "dr_common.cuh"
#include "dr_tc.cuh"

using namespace drtc;

namespace {

constexpr int kProbeThreads = 160;   // warps 0-3: loaders + epilogue (TMEM lanes), warp 4: MMA issuer

template <int CG, bool ATMEM>
__global__ void __launch_bounds__(kProbeThreads, 1)
dr_tc_probe_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int N, int K, int flags,
                   float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int KB = K / 64;
    const int Nloc = N / CG;
    const uint32_t a_bytes = ATMEM ? 0u : (uint32_t)KB * 128u * 128u;
    const uint32_t b_bytes = (uint32_t)KB * (uint32_t)Nloc * 128u;
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + ((a_bytes + 1023u) & ~1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + ((b_bytes + 1023u) & ~1023u));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t load_bar = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[1]);

    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t cta = (CG == 2) ? cluster_ctarank() : 0u;

    if (warp == 0) { tmem_alloc<CG>(smem_u32(tmem_slot), 512); tmem_relinquish<CG>(); }
    if (tid == 0) { mbar_init(load_bar, 1); mbar_init(mma_bar, 1); fence_mbar_init(); }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;

    if (tid == 0) {
        mbar_expect_tx(load_bar, a_bytes + b_bytes);
        if (!ATMEM) bulk_g2s(smem_u32(a_s), a + (size_t)cta * a_bytes, a_bytes, load_bar);
        bulk_g2s(smem_u32(b_s), b + (size_t)cta * b_bytes, b_bytes, load_bar);
    }
    mbar_wait(load_bar, 0);

    constexpr uint32_t kACol = 256;     // A operand columns in TMEM (TS variants)
    if constexpr (ATMEM) {
        if (warp < 4) {
            const __nv_bfloat16* arow = reinterpret_cast<const __nv_bfloat16*>(a) + ((size_t)cta * 128 + tid) * K;
            for (int c0 = 0; c0 < K / 2; c0 += 8) {
                uint32_t v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint32_t e0 = __bfloat16_as_ushort(arow[2 * (c0 + j)]);
                    uint32_t e1 = __bfloat16_as_ushort(arow[2 * (c0 + j) + 1]);
                    v[j] = (flags & 1) ? (e1 | (e0 << 16)) : (e0 | (e1 << 16));
                }
                tmem_st8(tbase + ((uint32_t)(warp * 32) << 16) + kACol + c0, v);
            }
            tc_wait_st();
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();

    if (warp == 4 && cta == 0) {
        if (elect_one()) {
            const uint32_t idesc = (flags & 2) ? make_idesc_f16(128 * CG, N) : make_idesc_bf16(128 * CG, N);
            uint32_t acc = 0;
            for (int kb = 0; kb < KB; ++kb)
                for (int k16 = 0; k16 < 4; ++k16) {
                    uint64_t bd = make_desc_sw128(smem_u32(b_s) + kb * Nloc * 128 + k16 * 32);
                    if constexpr (ATMEM) {
                        mma_ts<CG>(tbase, tbase + kACol + (kb * 64 + k16 * 16) / 2, bd, idesc, acc);
                    } else {
                        uint64_t ad = make_desc_sw128(smem_u32(a_s) + kb * 128 * 128 + k16 * 32);
                        mma_ss<CG>(tbase, ad, bd, idesc, acc);
                    }
                    acc = 1;
                }
            if constexpr (CG == 2) mma_commit_2(mma_bar, 0x3); else mma_commit_1(mma_bar);
        }
        __syncwarp();
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();

    if (warp < 4) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
            tc_wait_ld();
            float* o = out + ((size_t)cta * 128 + tid) * N + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 0) tmem_dealloc<CG>(tbase, 512);
}

template <int CG, bool ATMEM>
int run_probe(const uint8_t* a, const uint8_t* b, int N, int K, int flags, float* out, std::string& err) {
    int KB = K / 64, Nloc = N / CG;
    size_t a_bytes = ATMEM ? 0 : (size_t)KB * 128 * 128, b_bytes = (size_t)KB * Nloc * 128;
    size_t smem = ((a_bytes + 1023) & ~size_t(1023)) + ((b_bytes + 1023) & ~size_t(1023)) + 64 + 1024;
    cudaError_t e = cudaFuncSetAttribute(dr_tc_probe_kernel<CG, ATMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return DR_ECUDA; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(kProbeThreads); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, dr_tc_probe_kernel<CG, ATMEM>, a, b, N, K, flags, out);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { err = cudaGetErrorString(e); return DR_ECUDA; }
    return DR_OK;
}


// ---- second probe: operands in the MN-major SW128 canonical layout (the reduced index K is the slow one) ----
// D[128 x N] = sum_k A[k][m] * B[k][n], both operand images [MN/64 blocks][K/8 groups][8 k-rows x 128 B] with the 16-byte
// chunks of a row XOR-swizzled by (k & 7).  Descriptor fields come from the caller so that tests/test_gpu_tc_probe.py can pin
// (and, on a mismatch, diagnose) the LBO / SBO / K-advance convention the bf16 weight-gradient kernel relies on.
struct ProbeMn { uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_mn, b_mn, a_bytes, b_bytes; };

__host__ __device__ inline uint64_t make_desc_generic(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(kProbeThreads, 1)
dr_tc_probe_mn_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int N, int K, ProbeMn p, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* a_s = smem;
    uint8_t* b_s = smem + ((p.a_bytes + 1023u) & ~1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(b_s + ((p.b_bytes + 1023u) & ~1023u));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    const uint32_t load_bar = smem_u32(&bars[0]), mma_bar = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) { tmem_alloc<1>(smem_u32(tmem_slot), 512); tmem_relinquish<1>(); }
    if (tid == 0) { mbar_init(load_bar, 1); mbar_init(mma_bar, 1); fence_mbar_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = *tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(load_bar, p.a_bytes + p.b_bytes);
        bulk_g2s(smem_u32(a_s), a, p.a_bytes, load_bar);
        bulk_g2s(smem_u32(b_s), b, p.b_bytes, load_bar);
    }
    mbar_wait(load_bar, 0);
    if (warp == 4) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, N) | (p.a_mn << 15) | (p.b_mn << 16);
            for (int k16 = 0; k16 < K / 16; ++k16) {
                const uint64_t ad = make_desc_generic(smem_u32(a_s) + k16 * p.a_kstep, p.a_lbo, p.a_sbo);
                const uint64_t bd = make_desc_generic(smem_u32(b_s) + k16 * p.b_kstep, p.b_lbo, p.b_sbo);
                mma_ss<1>(tbase, ad, bd, idesc, k16 ? 1u : 0u);
            }
            mma_commit_1(mma_bar);
        }
        __syncwarp();
    }
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    if (warp < 4) {
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
            tc_wait_ld();
            float* o = out + (size_t)tid * N + c0;
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<1>(tbase, 512);
}

}  // namespace

// variant: bit0 = A from TMEM (TS), bit1 = cta_group::2.  flags bit0: swap the 16-bit halves when packing A;
// flags bit1: operands are fp16 instead of bf16.
// a: SS -> swizzled image [CG][K/64][128 rows x 128 B]; TS -> raw bf16 [CG*128][K].
// b: swizzled image [CG][K/64][N/CG rows x 128 B].   d_out: fp32 [CG*128][N] (host).
extern "C" int dr_tc_probe(int32_t variant, const void* a_host, size_t a_bytes, const void* b_host, size_t b_bytes,
                           int32_t N, int32_t K, int32_t flags, float* d_out_host) {
    if (!a_host || !b_host || !d_out_host || K % 64 || K < 64 || N % 32 || N < 32 || N > 256) return DR_EINVAL;
    int CG = (variant & 2) ? 2 : 1;
    uint8_t *da = nullptr, *db = nullptr; float* dout = nullptr;
    std::string err;
    int rc = DR_ECUDA;
    size_t out_bytes = (size_t)CG * 128 * N * sizeof(float);
    if (cudaMalloc(&da, a_bytes) == cudaSuccess && cudaMalloc(&db, b_bytes) == cudaSuccess &&
        cudaMalloc(&dout, out_bytes) == cudaSuccess &&
        cudaMemcpy(da, a_host, a_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(db, b_host, b_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemset(dout, 0xFF, out_bytes) == cudaSuccess) {
        switch (variant & 3) {
            case 0: rc = run_probe<1, false>(da, db, N, K, flags, dout, err); break;
            case 1: rc = run_probe<1, true>(da, db, N, K, flags, dout, err); break;
            case 2: rc = run_probe<2, false>(da, db, N, K, flags, dout, err); break;
            default: rc = run_probe<2, true>(da, db, N, K, flags, dout, err); break;
        }
        if (rc == DR_OK && cudaMemcpy(d_out_host, dout, out_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) rc = DR_ECUDA;
    }
    if (rc != DR_OK) dr_fail(nullptr, rc, "dr_tc_probe: " + (err.empty() ? std::string(cudaGetErrorString(cudaGetLastError())) : err));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return rc;
}

// MN-major probe.  params: 8 x uint32 {a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_mn, b_mn} (bytes; *_mn = operand major bit).
// a / b: ready shared-memory images (copied verbatim).  d_out: fp32 [128][N] (host).  N % 16 == 0, K % 16 == 0.
extern "C" int dr_tc_probe_mn(const void* a_host, size_t a_bytes, const void* b_host, size_t b_bytes, int32_t N, int32_t K,
                              const uint32_t* params, float* d_out_host) {
    if (!a_host || !b_host || !params || !d_out_host || K % 16 || K < 16 || N % 16 || N < 16 || N > 256 || a_bytes % 16 || b_bytes % 16) return DR_EINVAL;
    ProbeMn p{params[0], params[1], params[2], params[3], params[4], params[5], params[6] & 1u, params[7] & 1u, (uint32_t)a_bytes, (uint32_t)b_bytes};
    uint8_t *da = nullptr, *db = nullptr; float* dout = nullptr;
    int rc = DR_ECUDA;
    const size_t out_bytes = (size_t)128 * N * sizeof(float);
    const size_t smem = ((a_bytes + 1023) & ~size_t(1023)) + ((b_bytes + 1023) & ~size_t(1023)) + 64 + 1024;
    if (smem <= 227 * 1024 && cudaMalloc(&da, a_bytes) == cudaSuccess && cudaMalloc(&db, b_bytes) == cudaSuccess &&
        cudaMalloc(&dout, out_bytes) == cudaSuccess &&
        cudaMemcpy(da, a_host, a_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(db, b_host, b_bytes, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemset(dout, 0xFF, out_bytes) == cudaSuccess &&
        cudaFuncSetAttribute(dr_tc_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess) {
        dr_tc_probe_mn_kernel<<<1, kProbeThreads, smem>>>(da, db, N, K, p, dout);
        if (cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(d_out_host, dout, out_bytes, cudaMemcpyDeviceToHost) == cudaSuccess) rc = DR_OK;
    }
    if (rc != DR_OK) dr_fail(nullptr, rc, std::string("dr_tc_probe_mn: ") + cudaGetErrorString(cudaGetLastError()));
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return rc;
}
