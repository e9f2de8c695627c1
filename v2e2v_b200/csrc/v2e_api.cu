// v2e_api.cu — the extern "C" surface declared in include/v2e2v_b200.h.
//
// Validation, engine selection and error capture live here; the kernels are in
// v2e_stepwise.cu, v2e_persistent.cu and v2e_voxel.cu.  No torch types, no allocation.
#include "v2e_engine.cuh"

namespace v2e {

static std::atomic<long long> g_launches{0};
static thread_local cudaError_t t_last_cuda = cudaSuccess;

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int cuda_ok(cudaError_t e) {
    if (e == cudaSuccess) return V2E_OK;
    t_last_cuda = e;
    return V2E_ERR_CUDA;
}

static bool aligned16(const void* p) { return ((uintptr_t)p & 15u) == 0; }

static int make_cfg(const v2e_config* c, CfgDev& d) {
    if (!c || c->abi_version != V2E_ABI_VERSION) return V2E_ERR_BAD_ARG;
    if (c->height <= 0 || c->width <= 0 || c->num_bins < 1 || c->num_bins > V2E_MAX_BINS)
        return V2E_ERR_BAD_ARG;
    if ((long long)c->height * c->width > (1ll << 30)) return V2E_ERR_BAD_ARG;
    if (c->frame_dtype < V2E_DT_F32 || c->frame_dtype > V2E_DT_BF16) return V2E_ERR_BAD_ARG;
    if (c->voxel_dtype != V2E_DT_F32 && c->voxel_dtype != V2E_DT_BF16) return V2E_ERR_BAD_ARG;
    if (c->output_mode < V2E_OUT_VOXEL || c->output_mode > V2E_OUT_VOXEL_UNNORM)
        return V2E_ERR_BAD_ARG;
    if (c->noise_mode < V2E_NOISE_OFF || c->noise_mode > V2E_NOISE_PHILOX) return V2E_ERR_BAD_ARG;
    if (!c->lut) return V2E_ERR_BAD_ARG;
    d.H = c->height;
    d.W = c->width;
    d.P = c->height * c->width;
    d.nb = c->num_bins;
    d.cutoff_on = c->cutoff_on != 0;
    d.leak_on = c->leak_on != 0;
    d.shot_on = c->shot_on != 0;
    d.noise_mode = c->noise_mode;
    d.output_mode = c->output_mode;
    d.voxel_bf16 = c->voxel_dtype == V2E_DT_BF16;
    d.frames_integer = c->frames_integer != 0 || c->frame_dtype == V2E_DT_U8;
    d.leak_rate = c->leak_rate;
    d.leak_jitter = c->leak_jitter;
    d.rn_pos = c->rn_pos;
    d.rn_neg = c->rn_neg;
    d.shot_inten_m1 = c->shot_inten_m1;
    d.lut = c->lut;
    return V2E_OK;
}

static int check_pack(const CfgDev& cfg, const v2e_pack& p, int T) {
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    if (p.sc.num_frames != T) return V2E_ERR_UNSUPPORTED;  // one T per launch
    if (!p.frames || !p.base_log_frame || !p.timestamp_mem || !p.pos_thres || !p.neg_thres ||
        !p.num_events)
        return V2E_ERR_BAD_ARG;
    if (cfg.cutoff_on && !p.lp_log_frame0) return V2E_ERR_BAD_ARG;
    if (leak && !p.noise_rate) return V2E_ERR_BAD_ARG;
    if (cfg.output_mode == V2E_OUT_RAW) {
        if (!p.raw_events || p.raw_capacity < 0) return V2E_ERR_BAD_ARG;
    } else if (!p.voxel) {
        return V2E_ERR_BAD_ARG;
    }
    if (cfg.noise_mode == V2E_NOISE_INJECTED) {
        if (leak && !p.leak_randn) return V2E_ERR_BAD_ARG;
        if (shot && (!p.shot_rand || p.shot_imax <= 0)) return V2E_ERR_BAD_ARG;
    }
    const void* planes[] = {p.frames, p.base_log_frame, p.lp_log_frame0, p.timestamp_mem,
                            p.pos_thres, p.neg_thres, p.noise_rate, p.voxel};
    for (const void* q : planes)
        if (q && !aligned16(q)) return V2E_ERR_BAD_ARG;
    return V2E_OK;
}

}  // namespace v2e

using namespace v2e;

extern "C" {

int v2e_abi_version(void) { return V2E_ABI_VERSION; }

const char* v2e_error_string(int code) {
    switch (code) {
        case V2E_OK: return "ok";
        case V2E_ERR_BAD_ARG: return "bad argument (null / misaligned pointer or bad dims)";
        case V2E_ERR_UNSUPPORTED: return "unsupported request";
        case V2E_ERR_TIME_ORDER: return "frame time must be later than previous frame time";
        case V2E_ERR_CUDA: return "CUDA error";
        case V2E_ERR_WORKSPACE: return "workspace too small";
        case V2E_ERR_NO_DEVICE: return "no sm_100 CUDA device";
        default: return "unknown error";
    }
}

int v2e_last_cuda_error(void) { return (int)t_last_cuda; }
const char* v2e_last_cuda_error_string(void) { return cudaGetErrorString(t_last_cuda); }
int64_t v2e_launch_count(void) { return (int64_t)g_launches.load(); }

size_t v2e_workspace_bytes(const v2e_config* c, int n_seq) {
    CfgDev cfg;
    if (make_cfg(c, cfg) != V2E_OK || n_seq < 1 || n_seq > V2E_MAX_SEQ_PER_LAUNCH) return 0;
    return step_layout(cfg, n_seq).total;
}

int v2e_resolve_engine(const v2e_config* c, int n_seq) {
    CfgDev cfg;
    const int rc = make_cfg(c, cfg);
    if (rc != V2E_OK) return rc;
    if (n_seq < 1 || n_seq > V2E_MAX_SEQ_PER_LAUNCH) return V2E_ERR_BAD_ARG;
    if (c->engine == V2E_ENGINE_STEPWISE) return V2E_ENGINE_STEPWISE;
    if (c->engine < V2E_ENGINE_AUTO || c->engine > V2E_ENGINE_FIXPOINT) return V2E_ERR_BAD_ARG;
    if (c->engine == V2E_ENGINE_FIXPOINT) {
        const int fx = fixpoint_fits(cfg, c->frame_dtype, n_seq);
        if (fx < 0) return fx;
        return fx ? V2E_ENGINE_FIXPOINT : V2E_ERR_UNSUPPORTED;
    }
    if (c->engine == V2E_ENGINE_RUNAHEAD) {
        const int ra = runahead_fits(cfg, c->frame_dtype, n_seq);
        if (ra < 0) return ra;
        return ra ? V2E_ENGINE_RUNAHEAD : V2E_ERR_UNSUPPORTED;
    }
    // AUTO: the fixed-point engine serves the voxel modes (fastest on B200, DESIGN.md "Engines");
    // raw event rows go to the lock-step persistent engine; the run-ahead engine is opt-in.
    if (c->engine == V2E_ENGINE_AUTO && cfg.output_mode != V2E_OUT_RAW) {
        const int fx = fixpoint_fits(cfg, c->frame_dtype, n_seq);
        if (fx > 0) return V2E_ENGINE_FIXPOINT;
    }
    const int fits = persistent_fits(cfg, c->frame_dtype, n_seq);
    if (fits < 0) return fits;
    if (c->engine == V2E_ENGINE_PERSISTENT) return fits ? V2E_ENGINE_PERSISTENT : V2E_ERR_UNSUPPORTED;
    if (!fits && cfg.output_mode != V2E_OUT_RAW) {
        const int ra = runahead_fits(cfg, c->frame_dtype, n_seq);
        if (ra > 0) return V2E_ENGINE_RUNAHEAD;
    }
    return fits ? V2E_ENGINE_PERSISTENT : V2E_ENGINE_STEPWISE;
}

int v2e_pack_forward(const v2e_config* c, const v2e_pack* packs, int n_seq, void* workspace,
                     size_t workspace_bytes, void* stream) {
    CfgDev cfg;
    int rc = make_cfg(c, cfg);
    if (rc != V2E_OK) return rc;
    if (!packs || n_seq < 1 || n_seq > V2E_MAX_SEQ_PER_LAUNCH || !workspace) return V2E_ERR_BAD_ARG;
    const int T = packs[0].sc.num_frames;
    if (T < 2 || T > V2E_MAX_FRAMES) return V2E_ERR_BAD_ARG;
    for (int s = 0; s < n_seq; ++s)
        if ((rc = check_pack(cfg, packs[s], T)) != V2E_OK) return rc;
    if (workspace_bytes < step_layout(cfg, n_seq).total) return V2E_ERR_WORKSPACE;
    if (!aligned16(workspace)) return V2E_ERR_BAD_ARG;
    const int rounds = c->philox_rounds == 7 ? 7 : 10;
    const int engine = v2e_resolve_engine(c, n_seq);
    if (engine < 0) return engine;
    cudaStream_t st = (cudaStream_t)stream;
    if (engine == V2E_ENGINE_FIXPOINT)
        return run_fixpoint(cfg, c->frame_dtype, rounds, packs, n_seq, workspace, st);
    if (engine == V2E_ENGINE_RUNAHEAD)
        return run_runahead(cfg, c->frame_dtype, rounds, packs, n_seq, workspace, st);
    if (engine == V2E_ENGINE_PERSISTENT)
        return run_persistent(cfg, c->frame_dtype, rounds, packs, n_seq, workspace, st);
    return run_stepwise(cfg, c->frame_dtype, rounds, packs, n_seq, workspace, st);
}

int v2e_clip_forward(const v2e_config* c, const v2e_pack* packs, int n_packs, void* workspace,
                     size_t workspace_bytes, void* stream, void* const* events) {
    if (!packs || n_packs < 1) return V2E_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // Chains: when the fixed-point engine's specialised kernel serves the request, up to
    // V2E_MAX_SEQ_PER_LAUNCH consecutive packs go out as ONE launch (the kernel loops over the
    // packs; no launch gap, no set-up in between).  The first pack's event pair then brackets the
    // whole launch and the other pairs of the chain are empty intervals after it, so that the
    // intervals still add up to the device time of the clip.
    static const bool no_chain = getenv("V2E_NO_CHAIN") != nullptr;   // development aid
    CfgDev cfg;
    int i0 = 0;
    if (!no_chain && n_packs > 1 && make_cfg(c, cfg) == V2E_OK && workspace && aligned16(workspace) &&
        workspace_bytes >= step_layout(cfg, 1).total && v2e_resolve_engine(c, 1) == V2E_ENGINE_FIXPOINT) {
        const int rounds = c->philox_rounds == 7 ? 7 : 10;
        while (n_packs - i0 > 1) {
            const int m = n_packs - i0 < V2E_MAX_SEQ_PER_LAUNCH ? n_packs - i0 : V2E_MAX_SEQ_PER_LAUNCH;
            const int T = packs[i0].sc.num_frames;
            bool ok = T >= 2 && T <= V2E_MAX_FRAMES;
            for (int j = 0; ok && j < m; ++j) ok = check_pack(cfg, packs[i0 + j], T) == V2E_OK;
            if (!ok) break;   // (the per-pack path reports what is wrong)
            const int rc = run_fixpoint(cfg, c->frame_dtype, rounds, packs + i0, 1, workspace, st, m,
                                        events ? events[2 * i0] : nullptr);
            if (rc == V2E_ERR_UNSUPPORTED) break;   // nothing was enqueued: one launch per pack below
            if (rc != V2E_OK) return rc;
            if (events)
                for (int j = 0; j < m; ++j) {
                    void* const e1 = events[2 * (i0 + j) + 1];
                    void* const e0 = j + 1 < m ? events[2 * (i0 + j + 1)] : nullptr;
                    if (e1 && cuda_ok(cudaEventRecord((cudaEvent_t)e1, st)) != V2E_OK) return V2E_ERR_CUDA;
                    if (e0 && cuda_ok(cudaEventRecord((cudaEvent_t)e0, st)) != V2E_OK) return V2E_ERR_CUDA;
                }
            i0 += m;
        }
    }
    for (int i = i0; i < n_packs; ++i) {
        if (events && events[2 * i]) {
            const int rc = cuda_ok(cudaEventRecord((cudaEvent_t)events[2 * i], st));
            if (rc != V2E_OK) return rc;
        }
        const int rc = v2e_pack_forward(c, packs + i, 1, workspace, workspace_bytes, stream);
        if (rc != V2E_OK) return rc;
        if (events && events[2 * i + 1]) {
            const int rc2 = cuda_ok(cudaEventRecord((cudaEvent_t)events[2 * i + 1], st));
            if (rc2 != V2E_OK) return rc2;
        }
    }
    return V2E_OK;
}

int v2e_batch_clip_forward(const v2e_config* c, const v2e_pack* packs, int n_packs, int n_seq,
                           void* workspace, size_t workspace_bytes, void* stream,
                           void* const* events) {
    if (!packs || n_packs < 1 || n_seq < 1 || n_seq > V2E_MAX_SEQ_PER_LAUNCH) return V2E_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    // Chains, as in v2e_clip_forward: consecutive steps go out as ONE launch when the specialised
    // fixed-point kernel serves the request (as many steps as the launch-record ring holds).
    static const bool no_chain = getenv("V2E_NO_CHAIN") != nullptr;   // development aid
    CfgDev cfg;
    int i0 = 0;
    if (!no_chain && n_packs > 1 && make_cfg(c, cfg) == V2E_OK && workspace && aligned16(workspace) &&
        workspace_bytes >= step_layout(cfg, n_seq).total &&
        v2e_resolve_engine(c, n_seq) == V2E_ENGINE_FIXPOINT) {
        const int rounds = c->philox_rounds == 7 ? 7 : 10;
        int cap = kFxBlkSlots / n_seq;     // (launch records per launch)
        if (cap > V2E_MAX_SEQ_PER_LAUNCH) cap = V2E_MAX_SEQ_PER_LAUNCH;
        while (cap > 1 && n_packs - i0 > 1) {
            const int m = n_packs - i0 < cap ? n_packs - i0 : cap;
            const int T = packs[(size_t)i0 * n_seq].sc.num_frames;
            bool ok = T >= 2 && T <= V2E_MAX_FRAMES;
            for (int j = 0; ok && j < m * n_seq; ++j)
                ok = check_pack(cfg, packs[(size_t)i0 * n_seq + j], T) == V2E_OK;
            if (!ok) break;   // (the per-step path reports what is wrong)
            const int rc = run_fixpoint(cfg, c->frame_dtype, rounds, packs + (size_t)i0 * n_seq, n_seq,
                                        workspace, st, m, events ? events[2 * i0] : nullptr);
            if (rc == V2E_ERR_UNSUPPORTED) break;   // nothing was enqueued: one launch per step below
            if (rc != V2E_OK) return rc;
            if (events)      // the first step's pair brackets the launch, the others are empty intervals
                for (int j = 0; j < m; ++j) {
                    void* const e1 = events[2 * (i0 + j) + 1];
                    void* const e0 = j + 1 < m ? events[2 * (i0 + j + 1)] : nullptr;
                    if (e1 && cuda_ok(cudaEventRecord((cudaEvent_t)e1, st)) != V2E_OK) return V2E_ERR_CUDA;
                    if (e0 && cuda_ok(cudaEventRecord((cudaEvent_t)e0, st)) != V2E_OK) return V2E_ERR_CUDA;
                }
            i0 += m;
        }
    }
    for (int i = i0; i < n_packs; ++i) {
        if (events && events[2 * i]) {
            const int rc = cuda_ok(cudaEventRecord((cudaEvent_t)events[2 * i], st));
            if (rc != V2E_OK) return rc;
        }
        const int rc = v2e_pack_forward(c, packs + (size_t)i * n_seq, n_seq, workspace, workspace_bytes,
                                        stream);
        if (rc != V2E_OK) return rc;
        if (events && events[2 * i + 1]) {
            const int rc2 = cuda_ok(cudaEventRecord((cudaEvent_t)events[2 * i + 1], st));
            if (rc2 != V2E_OK) return rc2;
        }
    }
    return V2E_OK;
}

}  // extern "C"
