// Expert-sharded forward behind the C ABI: one call per rank, no NCCL, no SM-resident collective.
//
// The cross-expert mean (qrnn.py:46-52) forces ONE exchange: head i needs S = sum over ALL experts of their GRU outputs.
// Each rank owns M/world experts and produces a partial S.  Instead of an all-reduce (whose reduction CTAs can only run
// where a recurrence CTA has retired — the recurrence kernel owns every SM), the partials are EXCHANGED BY THE COPY ENGINES
// and SUMMED BY THE HEAD KERNEL:
//   * every rank owns an "arena" in device memory (S slots [set][source rank][chunk], two full forecast tensors, a page
//     of flags) that every peer maps (CUDA IPC between processes, plain peer access inside one process);
//   * the recurrence kernel runs as ONE launch over the whole batch and publishes, per 256-window tile, a completion flag
//     (csrc/dr_gru_tc.cu::TcTileSignal); a stream memory-wait on that flag (cuStreamWaitValue32 — no SM) releases strided 2-D
//     DMA copies of that tile's partial S into slot [rank] of every peer, followed by a 4-byte DMA "landed" signal;
//   * the head kernel of a chunk waits (stream memory-waits again) for the peers' signals and adds the `world` partials in
//     rank order while it converts S to operand images — bit-identical forecasts on every rank;
//   * the forecast columns of the chunk are placed into EVERY rank's stacked tensor [B,T,M,Q] by strided 2-D peer copies
//     (all-gather + layout interleave on the DMA engines).
// Buffers are double-buffered by forward parity and guarded by an "entered forward n" epoch signal, so consecutive forwards
// pipeline: the exchange tail of forward n runs under the recurrence of forward n+1.
#include <cuda.h>
#include <cstring>
#include <new>
#include "dr_common.cuh"

namespace {

constexpr int kMaxWorld = 8, kMaxChunks = 64, kChunk = 256;
constexpr int kRing = 4096;      // epoch staging slots: far more forwards than the launch queues can hold ahead of the GPU
typedef CUresult (*PFN_wait32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

// One arena per device and process, kept until the process exits: a later model on the same device (e.g. a second estimator
// with more experts) reuses the arena and the peer mappings instead of exchanging IPC handles again, and an exported
// allocation is never freed while an importer may still map it (the CUDA IPC teardown rule).  The arena must fit the largest
// model: create that one first, or set DR_COMM_ARENA_GB.  Epochs are per arena, so flags left by an earlier model are always
// older than any new forward's.
struct GlobalArena {
    uint8_t* base; size_t bytes;
    uint8_t* peer[kMaxWorld]; bool ipc_open[kMaxWorld]; bool have_peers;
    cudaIpcMemHandle_t handle; bool have_handle;
    unsigned int epoch;
    int world, rank;
};
GlobalArena g_arena[64];

struct DrComm {
    int world, rank, Bmax, T, nch_max;
    bool attached;
    GlobalArena* ga;
    uint8_t* arena; size_t bytes;
    uint8_t* peer[kMaxWorld];
    size_t off_S, s_slot, off_out, out_set;
    cudaStream_t cs, ss, xs, os, ds;               // recurrence | S sends | head kernels | forecast scatter | D2H (host entry)
    cudaEvent_t ev_call, ev_done[2], ev_k2[kMaxChunks], ev_x, ev_xs_end[2], ev_ss_end[2], ev_os_end[2], ev_d2h;
    unsigned int *tile_count, *tile_flag;          // local, written by the recurrence kernel
    unsigned int* d_ring; unsigned int* h_ring;    // epoch values staged for the 4-byte signal copies
    float* S_full[2]; size_t S_cap[2];
    float* P_full[2]; size_t P_cap[2];
    float* out_local[2]; size_t ol_cap[2];
    float* x_stage; size_t x_cap;
    PFN_wait32 wait32;
};

// flag page: [enter: world][sdone: world x kMaxChunks][odone: world x kMaxChunks] uint32
inline size_t flag_enter(int p) { return (size_t)p * 4; }
inline size_t flag_sdone(int p, int c) { return (size_t)(kMaxWorld + p * kMaxChunks + c) * 4; }
inline size_t flag_odone(int p, int c) { return (size_t)(kMaxWorld + kMaxWorld * kMaxChunks + p * kMaxChunks + c) * 4; }
constexpr size_t kFlagBytes = 8192;
static_assert((kMaxWorld + 2 * kMaxWorld * kMaxChunks) * 4 <= kFlagBytes, "flag page too small");

inline DrComm* comm_of(dr_model* m) { return reinterpret_cast<DrComm*>(m->comm); }

int wait_flag(dr_model* m, DrComm* c, cudaStream_t st, const void* addr, unsigned int value) {
    CUresult r = c->wait32(reinterpret_cast<CUstream>(st), (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) return dr_fail(m, DR_ECUDA, "cuStreamWaitValue32 failed (code " + std::to_string((int)r) + ")");
    return DR_OK;
}

void layout_arena(DrComm* c, int M_total) {
    c->off_S = kFlagBytes;
    c->s_slot = (size_t)c->T * 64 * kChunk * 4 * sizeof(float);
    c->off_out = c->off_S + (size_t)2 * c->world * c->nch_max * c->s_slot;
    c->out_set = (size_t)c->Bmax * c->T * M_total * DR_Q * sizeof(float);
    c->bytes = c->off_out + 2 * c->out_set;
}

}  // namespace

// drain this handle's queued sharded work
static void comm_detach(DrComm* c) {
    cudaStream_t sts[] = {c->cs, c->ss, c->xs, c->os, c->ds};
    for (cudaStream_t s : sts) if (s) cudaStreamSynchronize(s);
    // the peer mappings belong to the process-wide arena and stay open (see GlobalArena)
}

void dr_comm_free(dr_model* m) {
    DrComm* c = comm_of(m);
    if (!c) return;
    cudaStream_t sts[] = {c->cs, c->ss, c->xs, c->os, c->ds};
    comm_detach(c);
    void* ptrs[] = {c->tile_count, c->tile_flag, c->d_ring, c->S_full[0], c->S_full[1], c->P_full[0], c->P_full[1],
                    c->out_local[0], c->out_local[1], c->x_stage};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->h_ring) cudaFreeHost(c->h_ring);
    for (cudaStream_t s : sts) if (s) cudaStreamDestroy(s);
    cudaEvent_t evs[] = {c->ev_call, c->ev_done[0], c->ev_done[1], c->ev_x, c->ev_xs_end[0], c->ev_xs_end[1], c->ev_ss_end[0], c->ev_ss_end[1],
                         c->ev_os_end[0], c->ev_os_end[1], c->ev_d2h};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    for (int i = 0; i < kMaxChunks; ++i) if (c->ev_k2[i]) cudaEventDestroy(c->ev_k2[i]);
    delete c;
    m->comm = nullptr;
}

extern "C" {

int64_t dr_comm_arena_bytes(const dr_model* m, int32_t Bmax, int32_t T) {
    if (!m || Bmax < 1 || T < 1) return -1;
    DrComm tmp{};
    tmp.world = m->cfg.world; tmp.Bmax = Bmax; tmp.T = T; tmp.nch_max = (Bmax + kChunk - 1) / kChunk;
    layout_arena(&tmp, m->cfg.M);
    return (int64_t)tmp.bytes;
}

int dr_comm_init(dr_model* m, int32_t Bmax, int32_t T, void* ipc_handle_out, void** arena_ptr_out) {
    if (!m) return DR_EINVAL;
    if (Bmax < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_comm_init: bad shape");
    if (m->cfg.world < 2 || m->cfg.world > kMaxWorld) return dr_fail(m, DR_EINVAL, "dr_comm_init: world must be 2..8");
    if ((Bmax + kChunk - 1) / kChunk > kMaxChunks) return dr_fail(m, DR_EINVAL, "dr_comm_init: at most 16384 windows per call");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    dr_comm_free(m);
    DrComm* c = new (std::nothrow) DrComm();
    if (!c) return dr_fail(m, DR_ENOMEM, "host allocation failed");
    memset(c, 0, sizeof(*c));
    m->comm = c;
    c->world = m->cfg.world; c->rank = m->cfg.rank; c->Bmax = Bmax; c->T = T; c->nch_max = (Bmax + kChunk - 1) / kChunk;
    layout_arena(c, m->cfg.M);
    {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        cudaError_t e = cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qr);
        if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn)
            return dr_fail(m, DR_ECUDA, "cuStreamWaitValue32 is not available from this driver (stream memory operations are required)");
        c->wait32 = reinterpret_cast<PFN_wait32>(fn);
    }
    {
        GlobalArena* ga = &g_arena[m->cfg.device & 63];
        if (ga->base && (ga->world != c->world || ga->rank != c->rank))
            return dr_fail(m, DR_ESTATE, "this device already holds an arena of a different world/rank in this process");
        if (ga->base && ga->bytes < c->bytes)
            return dr_fail(m, DR_ENOMEM, "the process-wide exchange arena of this device (" + std::to_string(ga->bytes >> 20) + " MiB) is smaller than this model needs (" +
                                         std::to_string(c->bytes >> 20) + " MiB) and cannot be re-allocated once peers have mapped it: create the largest "
                                         "model first or set DR_COMM_ARENA_GB");
        if (!ga->base) {
            size_t want = c->bytes;
            if (const char* gb = getenv("DR_COMM_ARENA_GB")) { const size_t v = (size_t)atof(gb) << 30; if (v > want) want = v; }
            DR_CUDA(m, cudaMalloc((void**)&ga->base, want));
            DR_CUDA(m, cudaMemset(ga->base, 0, kFlagBytes));
            ga->bytes = want; ga->world = c->world; ga->rank = c->rank; ga->epoch = 0;
        }
        c->ga = ga; c->arena = ga->base;
    }
    DR_CUDA(m, cudaMalloc((void**)&c->tile_count, kMaxChunks * sizeof(unsigned int)));
    DR_CUDA(m, cudaMalloc((void**)&c->tile_flag, kMaxChunks * sizeof(unsigned int)));
    DR_CUDA(m, cudaMemset(c->tile_count, 0, kMaxChunks * sizeof(unsigned int)));
    DR_CUDA(m, cudaMemset(c->tile_flag, 0, kMaxChunks * sizeof(unsigned int)));
    DR_CUDA(m, cudaMalloc((void**)&c->d_ring, kRing * sizeof(unsigned int)));
    DR_CUDA(m, cudaMallocHost((void**)&c->h_ring, kRing * sizeof(unsigned int)));
    int lo = 0, hi = 0;
    DR_CUDA(m, cudaDeviceGetStreamPriorityRange(&lo, &hi));        // hi = numerically lowest = highest priority
    DR_CUDA(m, cudaStreamCreateWithPriority(&c->cs, cudaStreamNonBlocking, lo));
    DR_CUDA(m, cudaStreamCreateWithPriority(&c->ss, cudaStreamNonBlocking, hi));
    DR_CUDA(m, cudaStreamCreateWithPriority(&c->xs, cudaStreamNonBlocking, hi));   // head kernels get freed SMs before queued recurrence CTAs
    DR_CUDA(m, cudaStreamCreateWithPriority(&c->os, cudaStreamNonBlocking, hi));
    DR_CUDA(m, cudaStreamCreateWithPriority(&c->ds, cudaStreamNonBlocking, hi));
    cudaEvent_t* evs[] = {&c->ev_call, &c->ev_done[0], &c->ev_done[1], &c->ev_x, &c->ev_xs_end[0], &c->ev_xs_end[1], &c->ev_ss_end[0], &c->ev_ss_end[1],
                          &c->ev_os_end[0], &c->ev_os_end[1], &c->ev_d2h};
    for (cudaEvent_t* e : evs) DR_CUDA(m, cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i) DR_CUDA(m, cudaEventCreateWithFlags(&c->ev_k2[i], cudaEventDisableTiming));
    DR_CUDA(m, cudaDeviceSynchronize());
    if (ipc_handle_out) {
        if (!c->ga->have_handle) { DR_CUDA(m, cudaIpcGetMemHandle(&c->ga->handle, c->arena)); c->ga->have_handle = true; }
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
        memcpy(ipc_handle_out, &c->ga->handle, sizeof(cudaIpcMemHandle_t));
    }
    if (arena_ptr_out) *arena_ptr_out = c->arena;
    return DR_OK;
}

int dr_comm_detach(dr_model* m) {
    if (!m) return DR_EINVAL;
    DrComm* c = comm_of(m);
    if (!c) return DR_OK;
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    comm_detach(c);
    c->attached = false;
    return DR_OK;
}

int dr_comm_attach(dr_model* m, const void* ipc_handles, void* const* arena_ptrs) {
    if (!m) return DR_EINVAL;
    DrComm* c = comm_of(m);
    if (!c) return dr_fail(m, DR_ESTATE, "dr_comm_attach before dr_comm_init");
    if (!ipc_handles && !arena_ptrs) return dr_fail(m, DR_EINVAL, "dr_comm_attach: pass IPC handles (one process per GPU) or arena pointers (one process)");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    for (int p = 0; p < c->world; ++p) {
        if (p == c->rank) { c->peer[p] = c->arena; continue; }
        if (c->ga->peer[p]) { c->peer[p] = c->ga->peer[p]; continue; }      // mapped by an earlier model of this process
        if (arena_ptrs) {
            cudaPointerAttributes at;
            DR_CUDA(m, cudaPointerGetAttributes(&at, arena_ptrs[p]));
            int can = 0;
            DR_CUDA(m, cudaDeviceCanAccessPeer(&can, m->cfg.device, at.device));
            if (!can) return dr_fail(m, DR_ECUDA, "no peer access between the GPUs of this model");
            cudaError_t e = cudaDeviceEnablePeerAccess(at.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return dr_cuda_fail(m, e, "cudaDeviceEnablePeerAccess");
            cudaGetLastError();
            c->peer[p] = c->ga->peer[p] = reinterpret_cast<uint8_t*>(arena_ptrs[p]);
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, reinterpret_cast<const uint8_t*>(ipc_handles) + (size_t)p * sizeof(h), sizeof(h));
            void* ptr = nullptr;
            DR_CUDA(m, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            c->peer[p] = c->ga->peer[p] = reinterpret_cast<uint8_t*>(ptr);
            c->ga->ipc_open[p] = true;
        }
    }
    c->attached = true;
    return DR_OK;
}

}  // extern "C"

namespace {

// grows a per-set scratch buffer
int reserve_f(dr_model* m, float** p, size_t* cap, size_t bytes) { return dr_reserve(m, reinterpret_cast<void**>(p), cap, bytes); }

// The whole sharded forward, enqueued asynchronously.  x is on the device.  If out_host is given, this rank's own forecast
// columns are additionally copied to out_host[b,t,rank*M_loc..,:] (row pitch = M*Q floats) chunk by chunk.
int forward_sharded(dr_model* m, const float* x, int B, int T, float** out_dev, float* out_host, bool wait_caller, int32_t* ticket) {
    DrComm* c = comm_of(m);
    if (!c || !c->attached) return dr_fail(m, DR_ESTATE, "sharded forward before dr_comm_init / dr_comm_attach");
    if (B > c->Bmax || T != c->T) return dr_fail(m, DR_EINVAL, "sharded forward: shape exceeds what dr_comm_init sized the arena for");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "forward before dr_load_weights");
    const bool use_tc = (m->cfg.engine != DR_ENGINE_FFMA) && dr_tc_supported(m, B, T);
    if (!use_tc) return dr_fail(m, DR_EUNSUPPORTED, "dr_forward_sharded needs the tcgen05 engine (input_size <= 64); use the phase calls for the FFMA engine");
    const int world = c->world, rank = c->rank, Ml = m->M_loc;
    const int nloc = Ml * DR_Q, ntot = nloc * world;
    const int nch = (B + kChunk - 1) / kChunk;
    const int Bp = dr_s_rows(B);
    const unsigned int n = ++c->ga->epoch;
    const int set = (int)(n & 1u);
    cudaStream_t caller = m->stream;

    int rc;
    if ((rc = reserve_f(m, &c->S_full[set], &c->S_cap[set], dr_s_floats(B, T) * sizeof(float)))) return rc;
    if ((rc = reserve_f(m, &c->P_full[set], &c->P_cap[set], (size_t)T * (Bp / 128) * ((nloc + 15) / 16) * 4 * 16 * 128 * sizeof(float)))) return rc;
    if ((rc = reserve_f(m, &c->out_local[set], &c->ol_cap[set], (size_t)nch * kChunk * T * nloc * sizeof(float)))) return rc;

    // epoch value for the 4-byte signal copies of this forward
    const int ri = (int)(n % kRing);
    c->h_ring[ri] = n;
    DR_CUDA(m, cudaEventRecord(c->ev_call, caller));
    cudaStream_t all[] = {c->cs, c->ss, c->xs, c->os};
    for (cudaStream_t s : all) DR_CUDA(m, cudaStreamWaitEvent(s, c->ev_call, 0));
    // buffers of this parity were last used by forward n-2: its sends, head kernels and scatters must be done
    if (n > 2) {
        DR_CUDA(m, cudaStreamWaitEvent(c->cs, c->ev_ss_end[set], 0));
        DR_CUDA(m, cudaStreamWaitEvent(c->cs, c->ev_xs_end[set], 0));
        DR_CUDA(m, cudaStreamWaitEvent(c->xs, c->ev_os_end[set], 0));
    }
    // ---- xs: announce "entered forward n" to every peer (implies: my head kernels of forwards < n are done) ----
    DR_CUDA(m, cudaMemcpyAsync(c->d_ring + ri, c->h_ring + ri, sizeof(unsigned int), cudaMemcpyHostToDevice, c->xs));
    for (int w = 1; w < world; ++w) {
        const int p = (rank + w) % world;
        DR_CUDA(m, cudaMemcpyAsync(c->peer[p] + flag_enter(rank), c->d_ring + ri, sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->xs));
    }
    DR_CUDA(m, cudaEventRecord(c->ev_x, c->xs));
    DR_CUDA(m, cudaStreamWaitEvent(c->ss, c->ev_x, 0));           // d_ring[ri] is written
    DR_CUDA(m, cudaStreamWaitEvent(c->os, c->ev_x, 0));
    // nobody writes into a peer before that peer has entered forward n (its S slots of this parity are consumed, its
    // forecast tensor of this parity may be overwritten)
    for (int w = 1; w < world; ++w) {
        const int p = (rank + w) % world;
        if ((rc = wait_flag(m, c, c->ss, c->arena + flag_enter(p), n))) return rc;
        if ((rc = wait_flag(m, c, c->os, c->arena + flag_enter(p), n))) return rc;
    }

    // ---- cs: ONE recurrence launch over the whole batch, publishing a flag per 256-window tile ----
    {
        m->stream = c->cs;
        DR_CUDA(m, cudaMemsetAsync(c->S_full[set], 0, dr_s_floats(B, T) * sizeof(float), c->cs));
        float* keep_p = m->d_p; size_t keep_cap = m->p_cap;
        m->d_p = c->P_full[set]; m->p_cap = c->P_cap[set];
        m->tile_count = c->tile_count; m->tile_flag = c->tile_flag; m->tile_value = n;
        const int slot = m->ws_slot;                                // x image workspace rotates as in dr_forward_local_dev
        m->ws_slot = (slot + 1) % 4;
        m->d_xtc = m->ws_xtc[slot]; m->xtc_cap = m->ws_xtc_cap[slot];
        rc = dr_launch_gru_tc(m, x, B, T, c->S_full[set], nullptr);
        m->ws_xtc[slot] = m->d_xtc; m->ws_xtc_cap[slot] = m->xtc_cap;
        m->tile_count = nullptr; m->tile_flag = nullptr;
        c->P_full[set] = m->d_p; c->P_cap[set] = m->p_cap;
        m->d_p = keep_p; m->p_cap = keep_cap;
        m->stream = caller;
        m->last_engine = "tcgen05";
        if (rc) return rc;
    }

    uint8_t* my_out = c->arena + c->off_out + (size_t)set * c->out_set;
    for (int ch = 0; ch < nch; ++ch) {
        const int b0 = ch * kChunk, bn = std::min(kChunk, B - b0);
        const int wrows = std::min(kChunk, Bp - b0);               // rows of this chunk present in the batch-wide S
        // ---- ss: as soon as tile ch is complete, its partial S goes to slot [rank][ch] of every peer ----
        if ((rc = wait_flag(m, c, c->ss, c->tile_flag + ch, n))) return rc;
        for (int w = 1; w < world; ++w) {
            const int p = (rank + w) % world;                       // start with the next rank: the ranks do not all hit one peer
            uint8_t* dst = c->peer[p] + c->off_S + (((size_t)set * world + rank) * c->nch_max + ch) * c->s_slot;
            DR_CUDA(m, cudaMemcpy2DAsync(dst, (size_t)kChunk * 16, c->S_full[set] + (size_t)b0 * 4, (size_t)Bp * 16, (size_t)wrows * 16,
                                         (size_t)T * 64, cudaMemcpyDeviceToDevice, c->ss));
            DR_CUDA(m, cudaMemcpyAsync(c->peer[p] + flag_sdone(rank, ch), c->d_ring + ri, sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->ss));
        }
        // ---- xs: head kernel of the chunk once the own tile and every peer's partial are there ----
        if ((rc = wait_flag(m, c, c->xs, c->tile_flag + ch, n))) return rc;
        for (int w = 1; w < world; ++w) {
            const int p = (rank + w) % world;
            if ((rc = wait_flag(m, c, c->xs, c->arena + flag_sdone(p, ch), n))) return rc;
        }
        const float* src[kMaxWorld]; int rows[kMaxWorld], sb0[kMaxWorld];
        for (int p = 0; p < world; ++p) {
            if (p == rank) { src[p] = c->S_full[set]; rows[p] = Bp; sb0[p] = b0; }
            else { src[p] = reinterpret_cast<const float*>(c->arena + c->off_S + (((size_t)set * world + p) * c->nch_max + ch) * c->s_slot); rows[p] = kChunk; sb0[p] = 0; }
        }
        float* ol = c->out_local[set] + (size_t)ch * kChunk * T * nloc;
        cudaEvent_t* pev = (ch == 0 || ch == nch - 1) ? dr_prof_slot(m) : nullptr;   // dr_profile: head events span first..last chunk
        if (pev && ch == 0) DR_CUDA(m, cudaEventRecord(pev[2], c->xs));
        m->stream = c->xs;
        rc = dr_launch_heads_tc_multi(m, src, rows, sb0, world, c->P_full[set], Bp / 128, b0 / 128, bn, T, ol);
        m->stream = caller;
        if (rc) return rc;
        if (pev && ch == nch - 1) { DR_CUDA(m, cudaEventRecord(pev[3], c->xs)); m->prof_n += 1; }
        DR_CUDA(m, cudaEventRecord(c->ev_k2[ch], c->xs));
        // ---- os: the chunk's forecast columns into EVERY rank's stacked tensor (strided 2-D copies), then "landed" ----
        DR_CUDA(m, cudaStreamWaitEvent(c->os, c->ev_k2[ch], 0));
        for (int w = 0; w < world; ++w) {
            const int p = (rank + 1 + w) % world;                   // peers first, own copy last
            float* dst = reinterpret_cast<float*>(c->peer[p] + c->off_out + (size_t)set * c->out_set) + ((size_t)b0 * T) * ntot + (size_t)rank * nloc;
            DR_CUDA(m, cudaMemcpy2DAsync(dst, (size_t)ntot * sizeof(float), ol, (size_t)nloc * sizeof(float), (size_t)nloc * sizeof(float),
                                         (size_t)bn * T, cudaMemcpyDeviceToDevice, c->os));
            if (p != rank)
                DR_CUDA(m, cudaMemcpyAsync(c->peer[p] + flag_odone(rank, ch), c->d_ring + ri, sizeof(unsigned int), cudaMemcpyDeviceToDevice, c->os));
        }
        if (out_host) {                                             // host entry point: own columns straight to the caller's tensor
            DR_CUDA(m, cudaStreamWaitEvent(c->ds, c->ev_k2[ch], 0));
            float* hdst = out_host + ((size_t)b0 * T) * ntot + (size_t)rank * nloc;
            DR_CUDA(m, cudaMemcpy2DAsync(hdst, (size_t)ntot * sizeof(float), ol, (size_t)nloc * sizeof(float), (size_t)nloc * sizeof(float),
                                         (size_t)bn * T, cudaMemcpyDeviceToHost, c->ds));
        }
    }
    DR_CUDA(m, cudaEventRecord(c->ev_ss_end[set], c->ss));
    DR_CUDA(m, cudaEventRecord(c->ev_xs_end[set], c->xs));
    // the stacked tensor on THIS rank is complete when every peer's columns of every chunk have landed
    for (int ch = 0; ch < nch; ++ch)
        for (int w = 1; w < world; ++w) {
            const int p = (rank + w) % world;
            if ((rc = wait_flag(m, c, c->os, c->arena + flag_odone(p, ch), n))) return rc;
        }
    DR_CUDA(m, cudaEventRecord(c->ev_os_end[set], c->os));
    DR_CUDA(m, cudaEventRecord(c->ev_done[set], c->os));
    if (wait_caller) DR_CUDA(m, cudaStreamWaitEvent(caller, c->ev_done[set], 0));
    if (ticket) *ticket = (int32_t)n;
    if (out_host) {
        DR_CUDA(m, cudaEventRecord(c->ev_d2h, c->ds));
        DR_CUDA(m, cudaStreamWaitEvent(caller, c->ev_d2h, 0));
    }
    if (out_dev) *out_dev = reinterpret_cast<float*>(my_out);
    return DR_OK;
}

}  // namespace

extern "C" {

int dr_forward_sharded_dev(dr_model* m, const float* x_dev, int32_t B, int32_t T, float** out_dev) {
    if (!m) return DR_EINVAL;
    if (!x_dev || !out_dev || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_forward_sharded_dev: bad argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return forward_sharded(m, x_dev, B, T, out_dev, nullptr, true, nullptr);
}

int dr_forward_sharded_issue_dev(dr_model* m, const float* x_dev, int32_t B, int32_t T, float** out_dev, int32_t* ticket) {
    if (!m) return DR_EINVAL;
    if (!x_dev || !out_dev || !ticket || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_forward_sharded_issue_dev: bad argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return forward_sharded(m, x_dev, B, T, out_dev, nullptr, false, ticket);
}

int dr_forward_sharded_wait(dr_model* m, int32_t ticket) {
    if (!m) return DR_EINVAL;
    DrComm* c = comm_of(m);
    if (!c || !c->attached) return dr_fail(m, DR_ESTATE, "dr_forward_sharded_wait before dr_comm_init / dr_comm_attach");
    const unsigned int n = (unsigned int)ticket;
    if (n == 0 || n > c->ga->epoch || c->ga->epoch - n > 1) return dr_fail(m, DR_EINVAL, "dr_forward_sharded_wait: the ticket must be one of the last two forwards");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    DR_CUDA(m, cudaStreamWaitEvent(m->stream, c->ev_done[n & 1u], 0));
    return DR_OK;
}

int dr_forward_sharded(dr_model* m, const float* x_host, int32_t B, int32_t T, float* out_host, float** out_dev) {
    if (!m) return DR_EINVAL;
    if (!x_host || !out_host || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "dr_forward_sharded: bad argument");
    DrComm* c = comm_of(m);
    if (!c || !c->attached) return dr_fail(m, DR_ESTATE, "sharded forward before dr_comm_init / dr_comm_attach");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    const size_t nx = (size_t)B * T * m->cfg.F * sizeof(float);
    int rc = dr_reserve(m, reinterpret_cast<void**>(&c->x_stage), &c->x_cap, nx);
    if (rc) return rc;
    DR_CUDA(m, cudaMemcpyAsync(c->x_stage, x_host, nx, cudaMemcpyHostToDevice, m->stream));
    float* dev = nullptr;
    rc = forward_sharded(m, c->x_stage, B, T, &dev, out_host, true, nullptr);
    if (rc) return rc;
    DR_CUDA(m, cudaStreamSynchronize(m->stream));                   // own columns are in out_host, the stacked tensor is complete on the device
    if (out_dev) *out_dev = dev;
    return DR_OK;
}

}  // extern "C"
