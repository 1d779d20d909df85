// bf16 training engine (dr_config.dtype == DR_DTYPE_BF16): shared layout of the saved activations and small helpers.
//
// Everything the backward pass re-reads is stored ONCE, in bf16, as ready tensor-core operand images: the unit of storage
// is a 64-column x 128-window "column block" of 16 KB laid out as [16 window groups][8 windows][128 B] with the eight
// 16-byte chunks of a 128-byte row XOR-swizzled by (window & 7).  Read along a row it is what one thread (= one window)
// of the recurrence kernels loads or stores (128 contiguous bytes); read as a whole it is at the same time
//   * the K-major SW128 operand image [rows = windows][K = 64 columns]        (A operand of the forward x-part), and
//   * the MN-major SW128 operand image [K = windows][MN = 64 columns]         (both operands of the weight-gradient
//     GEMMs, whose reduced index is the window) — pinned by tests/test_gpu_tc_probe.py::test_tcgen05_tile_mn_major,
// so the weight-gradient kernel moves its operands with plain bulk copies and never converts or transposes anything.
//
//   gate image  [dir][e][t][tile] -> 8 column blocks: (r, z, n, q) x 2 halves of the 128 hidden units      128 KB
//               the backward kernel overwrites it IN PLACE with (da_r, da_z, da_n, dq) — same blocks, same chunks
//   h image     [dir][e][t][tile] -> 2 column blocks: h_t                                                 32 KB
//   x image     [t][tile]         -> 1 column block: x_t (F <= 64, zero padded)                           16 KB
// = 2.5 KB per expert-window-step-direction, against 15.4 KB of fp32 traffic per step in the split-fp16 engine.
#pragma once
#include "dr_common.cuh"
#include "dr_tc.cuh"

namespace drt16 {

constexpr uint32_t kColBlk = 128 * 128;            // one column block: 128 windows x 64 columns bf16 = 16 KB
constexpr uint32_t kGateImg = 8 * kColBlk;         // (r,z,n,q) x 2 halves
constexpr uint32_t kHImg = 2 * kColBlk;

// byte offset of 16-byte chunk `chunk` (8 columns) of window `row` inside a column block
__host__ __device__ inline uint32_t img_off(int row, int chunk) {
    return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((chunk ^ (row & 7)) << 4);
}
// index of the (dir, e, t, tile) block
__host__ __device__ inline size_t blk_index(int dir, int e, int t, int tile, int M_loc, int T, int ntiles) {
    return (((size_t)dir * M_loc + e) * T + t) * ntiles + tile;
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}

// ---- 32-byte sector access to a thread's image row ----
// A thread of the recurrence kernels touches its window's 128-byte row 16 columns (two 16-byte chunks 2c, 2c+1) at a time.
// The two chunks are the halves of ONE aligned 32-byte sector (the swizzle flips only their order, by the row's parity), so
// they move with a single 256-bit load / store (LDG.256 / STG.256, sm_100).  This matters: with one row per lane every warp
// instruction touches 32 different cache lines, and the clock64 breakdown of the reverse chain (profiles/r02_bwd16_lsu.md)
// showed the kernel bound by exactly that — 22.7k of 25.9k cycles per step in loads + stores at one line per cycle.
__device__ __forceinline__ void ld256(const void* p, uint32_t (&v)[8]) {
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st256(void* p, const uint32_t (&v)[8]) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
// byte offset of the sector holding chunks 2c and 2c+1 of window `row` inside a column block
__host__ __device__ inline uint32_t sector_off(int row, int c) {
    return (uint32_t)(row >> 3) * 1024u + (uint32_t)(row & 7) * 128u + (uint32_t)((((2 * c) ^ (row & 7)) & ~1) << 4);
}
// v[0..7] = the 16 consecutive bf16 columns 16c .. 16c+15 of the row (as 8 pairs), in column order
__device__ __forceinline__ void ld_cols16(const uint8_t* colblk, int row, int c, uint32_t (&v)[8]) {
    uint32_t w[8];
    ld256(colblk + sector_off(row, c), w);
    const bool flip = row & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[i] = flip ? w[4 + i] : w[i]; v[4 + i] = flip ? w[i] : w[4 + i]; }
}
__device__ __forceinline__ void st_cols16(uint8_t* colblk, int row, int c, const uint32_t (&v)[8]) {
    uint32_t w[8];
    const bool flip = row & 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[i] = flip ? v[4 + i] : v[i]; w[4 + i] = flip ? v[i] : v[4 + i]; }
    st256(colblk + sector_off(row, c), w);
}

// ---- dropout keep decisions (qrnn.py:43): replayed uint8 mask (parity tests) or a counter-based draw ----
// One 64-bit hash serves 4 consecutive elements (16 bits each): element idx keeps iff lane(idx & 3) of hash(idx >> 2) >= thr16,
// thr16 = round(p * 65536).  Identical in the forward (r~ = keep * h), the backward (adjoint) and the head-gradient kernels.
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Drop {
    const uint8_t* mask;     // nullable: replayed keep mask, uint8 [M][B][T][2H] (reference rnn_out order)
    uint64_t seed;
    uint32_t thr16;
    float inv_keep;          // 1 / (1 - p)
};
// keep bits of the 16 consecutive elements starting at idx (idx % 16 == 0): bit j set -> element idx + j is kept
__device__ __forceinline__ uint32_t keep16(const Drop& d, size_t idx) {
    uint32_t bits = 0;
    if (d.mask) {
        const uint4 mv = *reinterpret_cast<const uint4*>(d.mask + idx);
        const uint32_t w[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) bits |= (((w[i] >> (8 * j)) & 0xFFu) ? 1u : 0u) << (4 * i + j);
        return bits;
    }
    if (d.thr16 == 0) return 0xFFFFu;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint64_t h = mix64(d.seed * 0x9E3779B97F4A7C15ull + (uint64_t)(idx >> 2) + i);
#pragma unroll
        for (int j = 0; j < 4; ++j) bits |= (((uint32_t)(h >> (16 * j)) & 0xFFFFu) >= d.thr16 ? 1u : 0u) << (4 * i + j);
    }
    return bits;
}
// keep bits of the 4 consecutive elements starting at idx (idx % 4 == 0)
__device__ __forceinline__ uint32_t keep4(const Drop& d, size_t idx) {
    if (d.mask) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(d.mask + idx);
        return ((w & 0xFFu) ? 1u : 0u) | ((w & 0xFF00u) ? 2u : 0u) | ((w & 0xFF0000u) ? 4u : 0u) | ((w & 0xFF000000u) ? 8u : 0u);
    }
    if (d.thr16 == 0) return 0xFu;
    const uint64_t h = mix64(d.seed * 0x9E3779B97F4A7C15ull + (uint64_t)(idx >> 2));
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) bits |= (((uint32_t)(h >> (16 * j)) & 0xFFFFu) >= d.thr16 ? 1u : 0u) << j;
    return bits;
}
}  // namespace drt16
