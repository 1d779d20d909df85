// C-ABI layer of libdeeprest_b200.so (see include/deeprest_b200.h for the contract and the
// reference interface each entry point replaces). No torch types, no exceptions across the ABI.
#include <cstdio>
#include <cstring>
#include <new>
#include "dr_common.cuh"

static thread_local std::string g_create_err;

int dr_fail(dr_model* m, int code, const std::string& msg) {
    if (m) m->err = msg; else g_create_err = msg;
    return code;
}
int dr_cuda_fail(dr_model* m, cudaError_t e, const char* what) {
    std::string msg = std::string(what) + ": " + cudaGetErrorString(e);
    return dr_fail(m, e == cudaErrorMemoryAllocation ? DR_ENOMEM : DR_ECUDA, msg);
}

int dr_reserve(dr_model* m, void** ptr, size_t* cap, size_t bytes) {
    if (*cap >= bytes && *ptr) return DR_OK;
    if (*ptr) { DR_CUDA(m, cudaFree(*ptr)); *ptr = nullptr; *cap = 0; }
    if (bytes == 0) bytes = 16;
    DR_CUDA(m, cudaMalloc(ptr, bytes));
    *cap = bytes;
    return DR_OK;
}

static int check_handle(dr_model* m) { return m ? DR_OK : DR_EINVAL; }

extern "C" {

int dr_version(void) { return 100; }

int dr_has_engine(int32_t engine) {
    if (engine == DR_ENGINE_AUTO || engine == DR_ENGINE_FFMA) return 1;
    if (engine == DR_ENGINE_TC) return dr_tc_built() ? 1 : 0;
    return 0;
}

const char* dr_last_error(const dr_model* m) { return m ? m->err.c_str() : g_create_err.c_str(); }

int dr_create(const dr_config* cfg, dr_model** out) {
    if (!cfg || !out) return dr_fail(nullptr, DR_EINVAL, "dr_create: null argument");
    *out = nullptr;
    if (cfg->H != DR_H) return dr_fail(nullptr, DR_EUNSUPPORTED, "hidden_layer_size must be 128 (reference default, qrnn.py:7)");
    if (cfg->Q != DR_Q) return dr_fail(nullptr, DR_EUNSUPPORTED, "exactly 3 quantiles supported (reference default, qrnn.py:8)");
    if (cfg->F < 1) return dr_fail(nullptr, DR_EINVAL, "input_size must be >= 1");
    if (cfg->M < 2) return dr_fail(nullptr, DR_EINVAL, "num_metrics must be >= 2 (the reference's forward stacks the M-1 other experts, qrnn.py:52)");
    if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return dr_fail(nullptr, DR_EINVAL, "bad rank/world");
    if (cfg->M % cfg->world) return dr_fail(nullptr, DR_EINVAL, "num_metrics must be divisible by world (equal expert shards)");
    if (cfg->engine < DR_ENGINE_AUTO || cfg->engine > DR_ENGINE_TC) return dr_fail(nullptr, DR_EINVAL, "bad engine");
    if (cfg->dtype != DR_DTYPE_F32 && cfg->dtype != DR_DTYPE_BF16) return dr_fail(nullptr, DR_EINVAL, "bad dtype");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return dr_fail(nullptr, DR_ECUDA, std::string("no CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return dr_fail(nullptr, DR_EINVAL, "bad device ordinal");
    ce = cudaSetDevice(cfg->device);
    if (ce != cudaSuccess) return dr_cuda_fail(nullptr, ce, "cudaSetDevice");
    cudaDeviceProp prop;
    ce = cudaGetDeviceProperties(&prop, cfg->device);
    if (ce != cudaSuccess) return dr_cuda_fail(nullptr, ce, "cudaGetDeviceProperties");
    if (prop.major != 10) return dr_fail(nullptr, DR_ECUDA, "libdeeprest_b200 is built for sm_100a (B200) only");

    dr_model* m = new (std::nothrow) dr_model();
    if (!m) return dr_fail(nullptr, DR_ENOMEM, "host allocation failed");
    m->cfg = *cfg;
    m->M_loc = cfg->M / cfg->world;
    m->e_lo = cfg->rank * m->M_loc;
    m->e_hi = m->e_lo + m->M_loc;
    m->Fp = (cfg->F + DR_KC - 1) / DR_KC * DR_KC;
    m->off = dr_blob_offsets(cfg->F);
    m->loaded = false;
    m->sm_count = prop.multiProcessorCount;
    m->last_engine = "none";
    m->launches = 0;
    m->profile = false; m->prof_n = 0; m->ev = nullptr;
    m->stream = nullptr; m->own_stream = nullptr;
    m->d_blob = m->d_mask = m->d_wf = m->d_bias4 = m->d_ct = m->d_abar = m->d_hb = nullptr;
    m->d_wtc = nullptr; m->wtc_bytes = 0;
    m->d_wihm = nullptr; m->d_grad = nullptr; m->d_adam_m = nullptr; m->d_adam_v = nullptr; m->adam_step = 0;
    m->train_ws = nullptr; m->train_mb = 0; m->d_dropmask = nullptr; m->dropmask_cap = 0;
    m->copy_stream = nullptr; m->stream2 = nullptr; m->d_tc_dbg = nullptr; m->tc_xdrop = 0;
    m->tile_count = nullptr; m->tile_flag = nullptr; m->tile_value = 0; m->comm = nullptr;
    for (int i = 0; i < 4; ++i) { m->ws_S[i] = nullptr; m->ws_S_cap[i] = 0; }
    m->x_bstride = 0; m->d_dn = nullptr; m->dn_on = false; m->dn_clamp = 0.0f;
    for (int i = 0; i < 10; ++i) m->ev_pipe[i] = nullptr;
    m->d_xT = nullptr; m->xT_cap = 0; m->d_xtc = nullptr; m->xtc_cap = 0; m->ws_slot = 0;
    for (int i = 0; i < 4; ++i) { m->ws_xT[i] = nullptr; m->ws_xT_cap[i] = 0; m->ws_xtc[i] = nullptr; m->ws_xtc_cap[i] = 0; m->ws_p[i] = nullptr; m->ws_p_cap[i] = 0; m->ws_key[i] = nullptr; m->ws_tc[i] = false; }
    m->d_p = nullptr; m->p_cap = 0; m->p_live = false; m->d_himg = nullptr;
    m->d_xtc_tr = nullptr; m->xtc_tr_cap = 0; m->d_whT = nullptr; m->whT_cap = 0;
    m->d_S = nullptr; m->S_cap = 0; m->d_out = nullptr; m->out_cap = 0;
    m->d_xin = nullptr; m->xin_cap = 0; m->d_loss = nullptr; m->d_y = nullptr; m->y_cap = 0;

    int rc = DR_OK;
    auto alloc = [&](float** p, size_t n) {
        if (rc != DR_OK) return;
        cudaError_t e = cudaMalloc((void**)p, (n ? n : 4) * sizeof(float));
        if (e != cudaSuccess) rc = dr_cuda_fail(nullptr, e, "cudaMalloc(weights)");
    };
    size_t Ml = (size_t)m->M_loc, KT = (size_t)m->Fp + DR_H;
    ce = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (ce != cudaSuccess) rc = dr_cuda_fail(nullptr, ce, "cudaStreamCreate");
    m->stream = m->own_stream;
    alloc(&m->d_blob, Ml * m->off.per_expert);
    alloc(&m->d_mask, Ml * cfg->F);
    alloc(&m->d_wihm, 2 * Ml * 3 * DR_H * cfg->F);
    alloc(&m->d_wf, Ml * 2 * 2 * KT * 3 * 64);
    alloc(&m->d_bias4, Ml * 2 * 4 * DR_H);
    alloc(&m->d_ct, Ml * 2 * DR_Q * DR_H);
    alloc(&m->d_abar, Ml * DR_Q * DR_2H);
    alloc(&m->d_hb, Ml * DR_Q);
    alloc(&m->d_loss, 8);
    if (rc != DR_OK) { std::string keep = g_create_err; dr_destroy(m); g_create_err = keep; return rc; }
    *out = m;
    return DR_OK;
}

void dr_destroy(dr_model* m) {
    if (!m) return;
    cudaSetDevice(m->cfg.device);
    if (m->own_stream) cudaStreamSynchronize(m->own_stream);
    dr_train_free(m);
    dr_comm_free(m);
    void* ptrs[] = {m->d_xtc_tr, m->d_whT, m->d_himg, m->d_dn, m->d_tc_dbg, m->d_wihm, m->d_grad, m->d_adam_m, m->d_adam_v, m->d_dropmask, m->d_blob, m->d_mask, m->d_wf, m->d_bias4, m->d_ct, m->d_abar, m->d_hb, m->d_wtc,
                    m->ws_xT[0], m->ws_xT[1], m->ws_xT[2], m->ws_xT[3], m->ws_xtc[0], m->ws_xtc[1], m->ws_xtc[2], m->ws_xtc[3], m->ws_p[0], m->ws_p[1], m->ws_p[2], m->ws_p[3], m->ws_S[0], m->ws_S[1], m->ws_S[2], m->ws_S[3], m->d_out, m->d_xin, m->d_loss, m->d_y};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
    if (m->stream2) cudaStreamDestroy(m->stream2);
    for (int i = 0; i < 10; ++i) if (m->ev_pipe[i]) cudaEventDestroy(m->ev_pipe[i]);
    if (m->ev) { for (int i = 0; i < 4 * DR_PROF_MAX; ++i) if (m->ev[i]) cudaEventDestroy(m->ev[i]); delete[] m->ev; }
    delete m;
}

int dr_set_stream(dr_model* m, void* cuda_stream, int32_t use_caller_stream) {
    if (check_handle(m)) return DR_EINVAL;
    m->stream = use_caller_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : m->own_stream;
    return DR_OK;
}

int dr_profile(dr_model* m, int32_t enable) {
    if (check_handle(m)) return DR_EINVAL;
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    if (enable && !m->ev) {
        m->ev = new (std::nothrow) cudaEvent_t[4 * DR_PROF_MAX]();
        if (!m->ev) return dr_fail(m, DR_ENOMEM, "host allocation failed");
        for (int i = 0; i < 4 * DR_PROF_MAX; ++i) DR_CUDA(m, cudaEventCreate(&m->ev[i]));
    }
    m->profile = enable != 0;
    m->prof_n = 0;
    return DR_OK;
}

int dr_profile_read(dr_model* m, int32_t* n_forwards, float* gru_ms_sum, float* head_ms_sum) {
    if (check_handle(m)) return DR_EINVAL;
    if (!m->ev) return dr_fail(m, DR_ESTATE, "dr_profile_read: profiling was never enabled");
    float g = 0.f, h = 0.f;
    for (int i = 0; i < m->prof_n; ++i) {
        float a = 0.f, b = 0.f;
        DR_CUDA(m, cudaEventSynchronize(m->ev[4 * i + 3]));
        DR_CUDA(m, cudaEventElapsedTime(&a, m->ev[4 * i + 0], m->ev[4 * i + 1]));
        DR_CUDA(m, cudaEventElapsedTime(&b, m->ev[4 * i + 2], m->ev[4 * i + 3]));
        g += a; h += b;
    }
    if (n_forwards) *n_forwards = m->prof_n;
    if (gru_ms_sum) *gru_ms_sum = g;
    if (head_ms_sum) *head_ms_sum = h;
    return DR_OK;
}

int dr_local_experts(const dr_model* m, int32_t* lo, int32_t* hi) {
    if (!m || !lo || !hi) return DR_EINVAL;
    *lo = m->e_lo; *hi = m->e_hi;
    return DR_OK;
}

int64_t dr_s_elems(int32_t B, int32_t T) { return (B < 1 || T < 1) ? -1 : (int64_t)dr_s_floats(B, T); }

int64_t dr_launch_count(const dr_model* m) { return m ? m->launches : -1; }
const char* dr_last_engine(const dr_model* m) { return m ? m->last_engine : "none"; }

int dr_load_weights(dr_model* m, const float* blob, size_t n) {
    if (check_handle(m)) return DR_EINVAL;
    size_t pe = (size_t)m->off.per_expert;
    if (!blob || n != pe * m->cfg.M)
        return dr_fail(m, DR_EINVAL, "dr_load_weights: blob must hold M*per_expert floats (" + std::to_string(pe * m->cfg.M) + ")");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    DR_CUDA(m, cudaMemcpyAsync(m->d_blob, blob + (size_t)m->e_lo * pe, (size_t)m->M_loc * pe * sizeof(float),
                               cudaMemcpyHostToDevice, m->stream));
    int rc = dr_launch_prep(m);
    if (rc != DR_OK) return rc;
    rc = dr_tc_prep_weights(m);
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaStreamSynchronize(m->stream));   // the caller's blob is not retained past return
    m->loaded = true;
    return DR_OK;
}

int dr_get_weights(dr_model* m, float* blob, size_t n) {
    if (check_handle(m)) return DR_EINVAL;
    size_t pe = (size_t)m->off.per_expert;
    if (!blob || n != pe * m->cfg.M) return dr_fail(m, DR_EINVAL, "dr_get_weights: wrong blob size");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "dr_get_weights before dr_load_weights");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    DR_CUDA(m, cudaMemcpyAsync(blob + (size_t)m->e_lo * pe, m->d_blob, (size_t)m->M_loc * pe * sizeof(float),
                               cudaMemcpyDeviceToHost, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    return DR_OK;
}

static int check_shape(dr_model* m, int B, int T) {
    if (B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "B and T must be >= 1");
    if (T > 65535) return dr_fail(m, DR_EINVAL, "T too large (max 65535)");
    if ((long long)B * T > (1LL << 31) / DR_2H) return dr_fail(m, DR_EINVAL, "B*T too large for one call; split the batch");
    if (!m->loaded) return dr_fail(m, DR_ESTATE, "forward before dr_load_weights");
    return DR_OK;
}

static int forward_local_slot(dr_model* m, const float* x, int32_t B, int32_t T, float* S, float* out_local) {
    int rc = DR_OK;
    bool use_tc = (m->cfg.engine == DR_ENGINE_TC) || (m->cfg.engine == DR_ENGINE_AUTO && dr_tc_supported(m, B, T));
    if (use_tc) {
        if (!dr_tc_supported(m, B, T)) return dr_fail(m, DR_EUNSUPPORTED, "tcgen05 engine does not support this shape");
        m->last_engine = "tcgen05";
        m->p_live = true;          // dr_forward_heads_dev (called next for this chunk) reads the partials from d_p
        return dr_launch_gru_tc(m, x, B, T, S, out_local);     // records its own profile events
    }
    int BT = 16 * dr_ffma_rows_per_thread(B);
    int Bp = (B + BT - 1) / BT * BT;
    rc = dr_reserve(m, (void**)&m->d_xT, &m->xT_cap, (size_t)T * m->Fp * Bp * sizeof(float));
    if (rc != DR_OK) return rc;
    rc = dr_launch_xT(m, x, B, T, Bp);
    if (rc != DR_OK) return rc;
    m->last_engine = "ffma";
    m->p_live = false;           // this engine REDs its own-expert head term straight into out_local
    DR_CUDA(m, cudaMemsetAsync(out_local, 0, (size_t)B * T * m->M_loc * DR_Q * sizeof(float), m->stream));
    cudaEvent_t* ev = dr_prof_slot(m);
    if (ev) DR_CUDA(m, cudaEventRecord(ev[0], m->stream));
    rc = dr_launch_gru_ffma(m, B, T, Bp, S, out_local);
    if (ev) DR_CUDA(m, cudaEventRecord(ev[1], m->stream));
    return rc;
}

int dr_forward_local_dev(dr_model* m, const float* x, int32_t B, int32_t T, float* S, float* out_local) {
    if (check_handle(m)) return DR_EINVAL;
    int rc = check_shape(m, B, T);
    if (rc != DR_OK) return rc;
    if (!x || !S || !out_local) return dr_fail(m, DR_EINVAL, "null device pointer");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    DR_CUDA(m, cudaMemsetAsync(S, 0, dr_s_floats(B, T) * sizeof(float), m->stream));
    if (m->M_loc == 0) return DR_OK;

    const int slot = m->ws_slot;
    m->ws_slot = (slot + 1) % 4;
    m->d_xT = m->ws_xT[slot]; m->xT_cap = m->ws_xT_cap[slot];
    m->d_xtc = m->ws_xtc[slot]; m->xtc_cap = m->ws_xtc_cap[slot];
    m->d_p = m->ws_p[slot]; m->p_cap = m->ws_p_cap[slot];
    for (int i = 0; i < 4; ++i) if (m->ws_key[i] == S) m->ws_key[i] = nullptr;     // S_dev is being reused: older tickets for it are void
    rc = forward_local_slot(m, x, B, T, S, out_local);
    m->ws_key[slot] = (rc == DR_OK) ? S : nullptr; m->ws_tc[slot] = m->p_live;
    m->ws_xT[slot] = m->d_xT; m->ws_xT_cap[slot] = m->xT_cap;
    m->ws_xtc[slot] = m->d_xtc; m->ws_xtc_cap[slot] = m->xtc_cap;
    m->ws_p[slot] = m->d_p; m->ws_p_cap[slot] = m->p_cap;
    return rc;
}

// The head phase consumes the own-expert partials its local phase left in one of the 4 workspace slots.  The pairing is
// by the S_dev pointer (the caller passes the same buffer to both phases), so chunk pipelines may interleave
// local(c0), local(c1), heads(c0), heads(c1); a heads call whose S_dev no live local call produced is an error, not a
// silent read of another chunk's partials.
static int bind_partials(dr_model* m, const float* S) {
    if (m->M_loc == 0) return DR_OK;
    for (int i = 0; i < 4; ++i)
        if (m->ws_key[i] == S && S != nullptr) {
            m->d_p = m->ws_p[i]; m->p_cap = m->ws_p_cap[i]; m->p_live = m->ws_tc[i];
            return DR_OK;
        }
    return dr_fail(m, DR_ESTATE, "head phase without a matching dr_forward_local_dev: no live local phase wrote this S_dev "
                                 "(at most 4 local phases may be in flight; each S_dev pairs one local with one heads call)");
}

int dr_forward_heads_dev(dr_model* m, const float* S, int32_t B, int32_t T, float* out_local) {
    if (check_handle(m)) return DR_EINVAL;
    int rc = check_shape(m, B, T);
    if (rc != DR_OK) return rc;
    if (!S || !out_local) return dr_fail(m, DR_EINVAL, "null device pointer");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    rc = bind_partials(m, S);
    if (rc != DR_OK) return rc;
    cudaEvent_t* ev = dr_prof_slot(m);
    if (ev) DR_CUDA(m, cudaEventRecord(ev[2], m->stream));
    rc = m->p_live ? dr_launch_heads_tc(m, S, B, T, out_local) : dr_launch_heads(m, S, B, T, out_local);
    if (ev) { DR_CUDA(m, cudaEventRecord(ev[3], m->stream)); m->prof_n += 1; }
    return rc;
}

int dr_forward_heads_p2p_dev(dr_model* m, const float* S, int32_t B, int32_t T, void* const* out_ptrs, int32_t n_ptrs,
                             int64_t row0) {
    if (check_handle(m)) return DR_EINVAL;
    int rc = check_shape(m, B, T);
    if (rc != DR_OK) return rc;
    if (!S || !out_ptrs || n_ptrs != m->cfg.world) return dr_fail(m, DR_EINVAL, "dr_forward_heads_p2p_dev: one destination per rank");
    rc = bind_partials(m, S);
    if (rc != DR_OK) return rc;
    if (!m->p_live) return dr_fail(m, DR_EUNSUPPORTED, "peer-write heads need the tcgen05 engine (input_size <= 64)");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    cudaEvent_t* ev = dr_prof_slot(m);
    if (ev) DR_CUDA(m, cudaEventRecord(ev[2], m->stream));
    rc = dr_launch_heads_tc_dst(m, S, B, T, out_ptrs, n_ptrs, row0);
    if (ev) { DR_CUDA(m, cudaEventRecord(ev[3], m->stream)); m->prof_n += 1; }
    return rc;
}

int dr_scatter_forecasts_dev(dr_model* m, const float* out_local, int32_t B, int32_t T, void* const* out_ptrs,
                             int32_t n_ptrs, int64_t row0) {
    if (check_handle(m)) return DR_EINVAL;
    if (!out_local || !out_ptrs || n_ptrs != m->cfg.world || B < 1 || T < 1)
        return dr_fail(m, DR_EINVAL, "dr_scatter_forecasts_dev: one destination per rank");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    const size_t nloc = (size_t)m->M_loc * DR_Q, ntot = nloc * m->cfg.world;
    const size_t rows = (size_t)B * T;
    for (int w = 0; w < n_ptrs; ++w) {
        // peers first, starting with the next rank, so that the 8 ranks do not all target the same peer at once
        int dstw = (m->cfg.rank + 1 + w) % n_ptrs;
        float* dst = reinterpret_cast<float*>(out_ptrs[dstw]) + ((size_t)row0 * T) * ntot + (size_t)m->cfg.rank * nloc;
        DR_CUDA(m, cudaMemcpy2DAsync(dst, ntot * sizeof(float), out_local, nloc * sizeof(float), nloc * sizeof(float), rows,
                                     cudaMemcpyDeviceToDevice, m->stream));
    }
    return DR_OK;
}

int dr_interleave_dev(dr_model* m, const float* gathered, int32_t B, int32_t T, float* out) {
    if (check_handle(m)) return DR_EINVAL;
    if (!gathered || !out || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "bad argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return dr_launch_interleave(m, gathered, B, T, out);
}

int dr_forward_dev(dr_model* m, const float* x, int32_t B, int32_t T, float* out) {
    if (check_handle(m)) return DR_EINVAL;
    if (m->cfg.world != 1)
        return dr_fail(m, DR_ESTATE, "dr_forward needs world == 1; sharded handles use dr_forward_local_dev / dr_forward_heads_dev");
    int rc = check_shape(m, B, T);
    if (rc != DR_OK) return rc;
    // S rotates through the same slots as the operand-image workspaces, so chunked callers may keep several
    // forwards in flight on different streams; d_S is the slot of the latest call (dr_debug_read "S")
    const int slot = m->ws_slot;
    m->d_S = m->ws_S[slot]; m->S_cap = m->ws_S_cap[slot];
    rc = dr_reserve(m, (void**)&m->d_S, &m->S_cap, dr_s_floats(B, T) * sizeof(float));
    m->ws_S[slot] = m->d_S; m->ws_S_cap[slot] = m->S_cap;
    if (rc != DR_OK) return rc;
    rc = dr_forward_local_dev(m, x, B, T, m->d_S, out);     // world == 1: out_local IS out [B,T,M,Q]
    if (rc != DR_OK) return rc;
    return dr_forward_heads_dev(m, m->d_S, B, T, out);
}

int dr_forward(dr_model* m, const float* x, int32_t B, int32_t T, float* out) {
    if (check_handle(m)) return DR_EINVAL;
    if (!x || !out) return dr_fail(m, DR_EINVAL, "null host pointer");
    int rc = check_shape(m, B, T);
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    size_t nx = (size_t)B * T * m->cfg.F, no = (size_t)B * T * m->cfg.M * DR_Q;
    rc = dr_reserve(m, (void**)&m->d_xin, &m->xin_cap, nx * sizeof(float));
    if (rc != DR_OK) return rc;
    rc = dr_reserve(m, (void**)&m->d_out, &m->out_cap, no * sizeof(float));
    if (rc != DR_OK) return rc;
    // Windows are independent, so a large batch runs as 2-4 chunks of whole 256-window pair tiles: the H2D copy of
    // the next chunks and the D2H copy of finished forecasts ride the copy engine (own stream, ordered with events)
    // while other chunks compute; the chunks alternate between two compute streams so that one chunk's recurrence
    // CTAs fill the SMs the previous chunk's last wave leaves idle.
    int nc = (B >= 1024) ? 4 : (B >= 512) ? 2 : 1;
    int Bc = (nc > 1) ? ((B / nc + 255) / 256) * 256 : B;
    if (nc > 1) nc = (B + Bc - 1) / Bc;
    if (nc == 1) {
        DR_CUDA(m, cudaMemcpyAsync(m->d_xin, x, nx * sizeof(float), cudaMemcpyHostToDevice, m->stream));
        rc = dr_forward_dev(m, m->d_xin, B, T, m->d_out);
        if (rc != DR_OK) return rc;
        DR_CUDA(m, cudaMemcpyAsync(out, m->d_out, no * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
        DR_CUDA(m, cudaStreamSynchronize(m->stream));
        return DR_OK;
    }
    if (!m->copy_stream) {
        DR_CUDA(m, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        DR_CUDA(m, cudaStreamCreateWithFlags(&m->stream2, cudaStreamNonBlocking));
        for (int i = 0; i < 10; ++i) DR_CUDA(m, cudaEventCreateWithFlags(&m->ev_pipe[i], cudaEventDisableTiming));
    }
    const size_t xrow = (size_t)T * m->cfg.F, orow = (size_t)T * m->cfg.M * DR_Q;
    cudaStream_t main_stream = m->stream;
    // nothing may start before earlier work on the caller's stream that still uses the staging buffers
    DR_CUDA(m, cudaEventRecord(m->ev_pipe[8], main_stream));
    DR_CUDA(m, cudaStreamWaitEvent(m->copy_stream, m->ev_pipe[8], 0));
    DR_CUDA(m, cudaStreamWaitEvent(m->stream2, m->ev_pipe[8], 0));
    for (int c = 0; c < nc; ++c) {
        int b0 = c * Bc, bn = (c == nc - 1) ? B - b0 : Bc;
        DR_CUDA(m, cudaMemcpyAsync(m->d_xin + b0 * xrow, x + b0 * xrow, bn * xrow * sizeof(float), cudaMemcpyHostToDevice, m->copy_stream));
        DR_CUDA(m, cudaEventRecord(m->ev_pipe[c], m->copy_stream));
    }
    for (int c = 0; c < nc; ++c) {
        int b0 = c * Bc, bn = (c == nc - 1) ? B - b0 : Bc;
        cudaStream_t cs = (c & 1) ? m->stream2 : main_stream;
        DR_CUDA(m, cudaStreamWaitEvent(cs, m->ev_pipe[c], 0));
        m->stream = cs;
        rc = dr_forward_dev(m, m->d_xin + b0 * xrow, bn, T, m->d_out + b0 * orow);
        m->stream = main_stream;
        if (rc != DR_OK) return rc;
        DR_CUDA(m, cudaEventRecord(m->ev_pipe[4 + c], cs));
        DR_CUDA(m, cudaStreamWaitEvent(m->copy_stream, m->ev_pipe[4 + c], 0));
        DR_CUDA(m, cudaMemcpyAsync(out + b0 * orow, m->d_out + b0 * orow, bn * orow * sizeof(float), cudaMemcpyDeviceToHost, m->copy_stream));
    }
    DR_CUDA(m, cudaStreamSynchronize(m->copy_stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream2));
    DR_CUDA(m, cudaStreamSynchronize(main_stream));
    return DR_OK;
}

// ---- N1: forward straight from the raw series (on-device windowing) ----
static int series_windows(int32_t N, int32_t W, int32_t stride) {
    // utils.py:4-5 builds windows for starts i in range(N - W) (it drops the last full window); estimate.py:85-86
    // then evaluates every `step_size`-th of them.  Starts kept here: 0, stride, 2*stride, ... < N - W.
    if (N - W <= 0 || stride < 1) return 0;
    return (N - W - 1) / stride + 1;
}

int dr_series_windows(int32_t N, int32_t W, int32_t stride) { return series_windows(N, W, stride); }

int dr_forward_series_dev(dr_model* m, const float* series, int32_t N, int32_t W, int32_t stride, float* out) {
    if (check_handle(m)) return DR_EINVAL;
    int B = series_windows(N, W, stride);
    if (!series || !out || B < 1) return dr_fail(m, DR_EINVAL, "dr_forward_series: the series is shorter than one window (the reference's sliding_window returns an empty array)");
    m->x_bstride = (long long)stride * m->cfg.F;
    int rc = dr_forward_dev(m, series, B, W, out);
    m->x_bstride = 0;
    return rc;
}

int dr_forward_series(dr_model* m, const float* series, int32_t N, int32_t W, int32_t stride, float* out) {
    if (check_handle(m)) return DR_EINVAL;
    int B = series_windows(N, W, stride);
    if (!series || !out || B < 1) return dr_fail(m, DR_EINVAL, "dr_forward_series: the series is shorter than one window (the reference's sliding_window returns an empty array)");
    int rc = check_shape(m, B, W);
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    size_t nx = (size_t)N * m->cfg.F, no = (size_t)B * W * m->cfg.M * DR_Q;
    if ((rc = dr_reserve(m, (void**)&m->d_xin, &m->xin_cap, nx * sizeof(float)))) return rc;
    if ((rc = dr_reserve(m, (void**)&m->d_out, &m->out_cap, no * sizeof(float)))) return rc;
    DR_CUDA(m, cudaMemcpyAsync(m->d_xin, series, nx * sizeof(float), cudaMemcpyHostToDevice, m->stream));   // N*F floats, not B*W*F
    rc = dr_forward_series_dev(m, m->d_xin, N, W, stride, m->d_out);
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaMemcpyAsync(out, m->d_out, no * sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    return DR_OK;
}

// ---- N2: clamp + de-normalise fused into the head kernel's epilogue ----
int dr_set_output_transform(dr_model* m, const float* scale, const float* offset, float clamp_min) {
    if (check_handle(m)) return DR_EINVAL;
    if (!scale || !offset) { m->dn_on = false; return DR_OK; }
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    if (!m->d_dn) DR_CUDA(m, cudaMalloc((void**)&m->d_dn, (size_t)2 * (m->M_loc ? m->M_loc : 1) * sizeof(float)));
    DR_CUDA(m, cudaMemcpyAsync(m->d_dn, scale + m->e_lo, m->M_loc * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    DR_CUDA(m, cudaMemcpyAsync(m->d_dn + m->M_loc, offset + m->e_lo, m->M_loc * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    m->dn_on = true; m->dn_clamp = clamp_min;
    return DR_OK;
}

int dr_quantile_loss_dev(dr_model* m, const float* out, const float* y, int32_t B, int32_t T, float* loss) {
    if (check_handle(m)) return DR_EINVAL;
    if (!out || !y || !loss || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "bad argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    return dr_launch_loss(m, out, y, B, T, loss);
}

int dr_quantile_loss(dr_model* m, const float* out, const float* y, int32_t B, int32_t T, float* loss) {
    if (check_handle(m)) return DR_EINVAL;
    if (!out || !y || !loss || B < 1 || T < 1) return dr_fail(m, DR_EINVAL, "bad argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    size_t no = (size_t)B * T * m->cfg.M * DR_Q, ny = (size_t)B * T * m->cfg.M;
    int rc = dr_reserve(m, (void**)&m->d_out, &m->out_cap, no * sizeof(float));
    if (rc != DR_OK) return rc;
    rc = dr_reserve(m, (void**)&m->d_y, &m->y_cap, ny * sizeof(float));
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaMemcpyAsync(m->d_out, out, no * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    DR_CUDA(m, cudaMemcpyAsync(m->d_y, y, ny * sizeof(float), cudaMemcpyHostToDevice, m->stream));
    rc = dr_launch_loss(m, m->d_out, m->d_y, B, T, m->d_loss + 4);
    if (rc != DR_OK) return rc;
    DR_CUDA(m, cudaMemcpyAsync(loss, m->d_loss + 4, sizeof(float), cudaMemcpyDeviceToHost, m->stream));
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    return DR_OK;
}

int dr_debug_read(dr_model* m, const char* what, float* host, size_t n) {
    if (check_handle(m)) return DR_EINVAL;
    if (!what || !host) return dr_fail(m, DR_EINVAL, "null argument");
    DR_CUDA(m, cudaSetDevice(m->cfg.device));
    if (!strcmp(what, "tc_timing_on")) {          // enable the in-kernel cycle breakdown (host_buf unused)
        if (!m->d_tc_dbg) { DR_CUDA(m, cudaMalloc((void**)&m->d_tc_dbg, 32 * sizeof(unsigned long long))); }
        DR_CUDA(m, cudaMemset(m->d_tc_dbg, 0, 32 * sizeof(unsigned long long)));
        return DR_OK;
    }
    if (!strncmp(what, "tc_xdrop", 8) && what[8] >= '0' && what[8] <= '2' && !what[9]) { m->tc_xdrop = what[8] - '0'; return DR_OK; }
    if (!strcmp(what, "tc_timing")) {             // 18 counters as floats (cycles), see dr_gru_tc.cu
        if (!m->d_tc_dbg || n < 18) return dr_fail(m, DR_EINVAL, "tc_timing: enable with tc_timing_on first; needs 18 floats");
        unsigned long long h[32];
        DR_CUDA(m, cudaStreamSynchronize(m->stream));
        DR_CUDA(m, cudaMemcpy(h, m->d_tc_dbg, sizeof(h), cudaMemcpyDeviceToHost));
        for (int i = 0; i < 18; ++i) host[i] = (float)h[i];
        return DR_OK;
    }
    const float* src = nullptr; size_t avail = 0;
    if (!strcmp(what, "mask")) { src = m->d_mask; avail = (size_t)m->M_loc * m->cfg.F; }
    else if (!strcmp(what, "S")) { src = m->d_S; avail = m->S_cap / sizeof(float); }
    else if (!strcmp(what, "ct")) { src = m->d_ct; avail = (size_t)m->M_loc * 2 * DR_Q * DR_H; }
    else if (!strcmp(what, "bias4")) { src = m->d_bias4; avail = (size_t)m->M_loc * 2 * 4 * DR_H; }
    else return dr_fail(m, DR_EINVAL, std::string("dr_debug_read: unknown tensor ") + what);
    if (!src || n > avail) return dr_fail(m, DR_EINVAL, "dr_debug_read: tensor not available at that size");
    DR_CUDA(m, cudaStreamSynchronize(m->stream));
    DR_CUDA(m, cudaMemcpy(host, src, n * sizeof(float), cudaMemcpyDeviceToHost));
    return DR_OK;
}

}  // extern "C"
