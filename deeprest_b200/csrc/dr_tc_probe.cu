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
