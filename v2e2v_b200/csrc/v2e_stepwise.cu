// v2e_stepwise.cu — the "stepwise" engine: two launches per frame, all state in HBM/L2.
//
// General-purpose fallback of v2e_pack_forward (any image size, any number of sequences per
// launch): phase A computes low-pass / leak / diff / count for every pixel and reduces the
// per-image max event count (v2e/v2e_model.py:375-406); phase B runs the per-pixel iteration
// loop that depends on that max (v2e/v2e_model.py:416-522).  The persistent engine
// (v2e_persistent.cu) fuses all of this into one launch with the state on chip; this engine is
// the reference point it is tested against and the path for images that do not fit on chip.
#include "v2e_engine.cuh"

namespace v2e {

constexpr int kThreads = 256;

struct StepParams {
    CfgDev cfg;
    int n_seq;
    int n;  // frame index of this launch
    StepWork* work;   // [n_seq]
    float* vox_acc;   // [n_seq][nb][P] f32 accumulators
    int* cntpol;      // [n_seq][P] packed (cnt << 2) | (pol + 1)
    SeqDev seq[V2E_MAX_SEQ_PER_LAUNCH];
    FrameSc fr[V2E_MAX_SEQ_PER_LAUNCH];  // scalars of frame n
};

// ---- init / rollover ---------------------------------------------------------------------------
template <int FDT>
__global__ void __launch_bounds__(kThreads) k_begin_pack(const __grid_constant__ StepParams prm) {
    __shared__ float2 s_lut[256];
    const CfgDev& cfg = prm.cfg;
    const SeqDev& sq = prm.seq[blockIdx.y];
    load_lut(s_lut, cfg.lut);
    __syncthreads();
    const int P = cfg.P;
    float* vox = prm.vox_acc + (size_t)blockIdx.y * cfg.nb * P;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        if (sq.first_pack) {
            // _init: v2e_model.py:170-174,239-240
            float L, I;
            log_inten(s_lut, load_frame1<FDT>(sq.frames, p), L, I);
            sq.base[p] = L;
            if (cfg.cutoff_on) sq.lp[p] = L;
            sq.tmem[p] = fadd(0.f, -sq.Tr);
        } else {
            // rollover: v2e_model.py:329-330
            float tm = sq.tmem[p];
            if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
            if (tm < 0.f) tm = -sq.Tr;
            sq.tmem[p] = tm;
        }
        if (cfg.output_mode != V2E_OUT_RAW)
            for (int b = 0; b < cfg.nb; ++b) vox[(size_t)b * P + p] = 0.f;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *sq.num_events = 0ull;
}

// ---- phase A -----------------------------------------------------------------------------------
template <int FDT, int R>
__global__ void __launch_bounds__(kThreads) k_phase_a(const __grid_constant__ StepParams prm) {
    __shared__ float2 s_lut[256];
    __shared__ int s_max[kThreads / 32];
    const CfgDev& cfg = prm.cfg;
    const SeqDev& sq = prm.seq[blockIdx.y];
    const FrameSc& fs = prm.fr[blockIdx.y];
    const int n = prm.n;
    load_lut(s_lut, cfg.lut);
    __syncthreads();
    const int P = cfg.P;
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool philox = cfg.noise_mode == V2E_NOISE_PHILOX;
    int* cntpol = prm.cntpol + (size_t)blockIdx.y * P;
    int lmax = 0;
    // one thread per 4-pixel group so the Philox stream layout matches the persistent engine
    const int G4 = (P + 3) / 4;
    for (int G = blockIdx.x * blockDim.x + threadIdx.x; G < G4; G += gridDim.x * blockDim.x) {
        uint4 w = make_uint4(0, 0, 0, 0);
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (philox) {
            w = noise_words<R>(sq.philox_ctr, n, G, 0);
            if (leak) {
                box_muller16(w.x, z[0], z[1]);
                box_muller16(w.y, z[2], z[3]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = 4 * G + e;
            if (p >= P) break;
            const float x = load_frame1<FDT>(sq.frames, (long long)n * P + p);
            float L, I;
            log_inten(s_lut, x, L, I);
            float nw = L;
            if (cfg.cutoff_on) {
                const float s = is_small_px(p, cfg.W) ? fs.s_small : fs.s_large;
                nw = lowpass(sq.lp[p], L, I, s);
                sq.lp[p] = nw;
            }
            float base = sq.base[p];
            const float pos = sq.pos[p], neg = sq.neg[p];
            if (leak) {
                float zz = z[e];
                if (!philox) zz = sq.leak_randn[(size_t)(n - 1) * P + p];
                else if (sq.dbg_leak_randn) sq.dbg_leak_randn[(size_t)(n - 1) * P + p] = zz;
                base = leak_sub(base, fmul(cfg.leak_rate, sq.nra[p]), cfg.leak_jitter, zz, fs.sdt,
                                pos);
                sq.base[p] = base;
            }
            const PixelA a = diff_count(nw, base, pos, neg);
            const int pol = (int)a.pol;
            int cand = 0;
            if (philox && shot && pol != 0) {
                const float thr256 = shot_thr256(pol > 0 ? fs.kp : fs.kn, cfg.shot_inten_m1, I, a.C);
                const unsigned byte = (w.z >> (8 * e)) & 0xffu;
                if ((float)byte < thr256)
                    cand = shot_uniform<R>(sq.philox_ctr, n, p, byte) * 256.0f < thr256;
            }
            cntpol[p] = (a.cnt << 3) | (cand << 2) | (pol + 1);
            lmax = max(lmax, a.cnt);
        }
    }
    // block max -> one atomicMax per block
    lmax = __reduce_max_sync(0xffffffffu, lmax);
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = lmax;
    __syncthreads();
    if (threadIdx.x == 0) {
        int m = 0;
        for (int i = 0; i < kThreads / 32; ++i) m = max(m, s_max[i]);
        if (m > 0) atomicMax((int*)&prm.work[blockIdx.y].bar[n] + 1, m);
    }
}

// ---- phase B -----------------------------------------------------------------------------------
template <int FDT, int R>
__global__ void __launch_bounds__(kThreads) k_phase_b(const __grid_constant__ StepParams prm) {
    __shared__ float2 s_lut[256];
    __shared__ int s_cnt[kThreads / 32];
    const CfgDev& cfg = prm.cfg;
    const SeqDev& sq = prm.seq[blockIdx.y];
    const FrameSc& fs = prm.fr[blockIdx.y];
    const int n = prm.n;
    load_lut(s_lut, cfg.lut);
    __syncthreads();
    const int P = cfg.P;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool philox = cfg.noise_mode == V2E_NOISE_PHILOX;
    StepWork& wk = prm.work[blockIdx.y];
    const int mx = ((const int*)&wk.bar[n])[1];
    const FrameGlobals g = frame_globals(mx, sq.duration, sq.Tr);
    if (blockIdx.x == 0 && threadIdx.x == 0 && sq.max_iters) sq.max_iters[n] = mx;
    const int* cntpol = prm.cntpol + (size_t)blockIdx.y * P;
    float* vox = prm.vox_acc + (size_t)blockIdx.y * cfg.nb * P;
    int nev = 0;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < P; p += gridDim.x * blockDim.x) {
        const int cp = cntpol[p];
        const int cnt = cp >> 3, cand = (cp >> 2) & 1, pol = (cp & 3) - 1;
        int fin = 0;
        const bool inj_shot = shot && !philox && mx > 0 && pol != 0;
        if (cnt > 0 || cand || inj_shot) {
            const float C = pol > 0 ? sq.pos[p] : sq.neg[p];
            float I = 0.f, u = 0.f;
            if (shot) {
                float L;
                log_inten(s_lut, load_frame1<FDT>(sq.frames, (long long)n * P + p), L, I);
                if (philox && cand) {
                    const uint4 w = noise_words<R>(sq.philox_ctr, n, p >> 2, 0);
                    u = shot_uniform<R>(sq.philox_ctr, n, p,
                                        (w.z >> (8 * (p & 3))) & 0xffu);
                }
            }
            float tm = sq.tmem[p];
            const int mode = inj_shot ? 1 : (cand ? 2 : 0);
            fin = iterate_pixel<R>(
                cfg, sq, fs, g, n, p, (float)pol, pol == 0 ? 0.f : C, cnt, I, tm, mode, u,
                [&](int b, float v) { vox[(size_t)b * P + p] = fadd(vox[(size_t)b * P + p], v); },
                [&](float ts, float polf) {
                    const unsigned long long slot = atomicAdd(&wk.raw_count, 1ull);
                    if ((long long)slot < sq.raw_capacity) {
                        float* row = sq.raw_events + 5 * slot;
                        const int y = p / cfg.W;
                        row[0] = ts; row[1] = (float)(p - y * cfg.W); row[2] = (float)y;
                        row[3] = polf; row[4] = 0.f;
                    }
                });
            sq.tmem[p] = tm;
            // memory update, v2e_model.py:522
            if (fin > 0) sq.base[p] = fadd(sq.base[p], fmul(fmul((float)pol, (float)fin), C));
            nev += fin;
        }
        if (sq.dbg_pol) sq.dbg_pol[(size_t)(n - 1) * P + p] = (signed char)pol;
        if (sq.dbg_cnt) sq.dbg_cnt[(size_t)(n - 1) * P + p] = cnt;
        if (sq.dbg_final) sq.dbg_final[(size_t)(n - 1) * P + p] = fin;
    }
    nev = __reduce_add_sync(0xffffffffu, nev);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = nev;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < kThreads / 32; ++i) t += s_cnt[i];
        if (t) atomicAdd(sq.num_events, (unsigned long long)t);
    }
}

// ---- normalisation (event_preprocess_pytorch, utils/event_process.py:165-175) --------------------
__global__ void __launch_bounds__(kThreads) k_vox_partials(const __grid_constant__ StepParams prm) {
    __shared__ double s_red[3][kThreads / 32];
    const CfgDev& cfg = prm.cfg;
    const size_t N = (size_t)cfg.nb * cfg.P;
    const float* vox = prm.vox_acc + (size_t)blockIdx.y * N;
    double s1 = 0., s2 = 0., nz = 0.;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < N;
         k += (size_t)gridDim.x * blockDim.x) {
        const float v = vox[k];
        s1 += (double)v;
        s2 += (double)v * (double)v;
        nz += (v != 0.f) ? 1. : 0.;
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nz += __shfl_down_sync(0xffffffffu, nz, o);
    }
    if ((threadIdx.x & 31) == 0) {
        s_red[0][threadIdx.x >> 5] = s1;
        s_red[1][threadIdx.x >> 5] = s2;
        s_red[2][threadIdx.x >> 5] = nz;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0., b = 0., c = 0.;
        for (int i = 0; i < kThreads / 32; ++i) { a += s_red[0][i]; b += s_red[1][i]; c += s_red[2][i]; }
        double* part = prm.work[blockIdx.y].partials + 3 * blockIdx.x;
        part[0] = a; part[1] = b; part[2] = c;
    }
}

__global__ void __launch_bounds__(kThreads) k_vox_finish(const __grid_constant__ StepParams prm,
                                                         int n_partials) {
    __shared__ float s_stat[3];
    const CfgDev& cfg = prm.cfg;
    const SeqDev& sq = prm.seq[blockIdx.y];
    const size_t N = (size_t)cfg.nb * cfg.P;
    const float* vox = prm.vox_acc + (size_t)blockIdx.y * N;
    StepWork& wk = prm.work[blockIdx.y];
    if (threadIdx.x == 0) {
        // fixed-order sum of the block partials: deterministic, identical in every block
        double a = 0., b = 0., c = 0.;
        for (int i = 0; i < n_partials; ++i) {
            a += wk.partials[3 * i]; b += wk.partials[3 * i + 1]; c += wk.partials[3 * i + 2];
        }
        float mean = 0.f, den = 1.f;
        if (c > 0. && cfg.output_mode == V2E_OUT_VOXEL) {
            mean = fdiv((float)a, (float)c);
            const float var = fsub(fdiv((float)b, (float)c), fmul(mean, mean));
            den = fadd(sqrtf(var), 1e-8f);
        }
        s_stat[0] = mean; s_stat[1] = den;
        s_stat[2] = (c > 0. && cfg.output_mode == V2E_OUT_VOXEL) ? 1.f : 0.f;
    }
    __syncthreads();
    const float mean = s_stat[0], den = s_stat[1];
    const bool norm = s_stat[2] != 0.f;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < N;
         k += (size_t)gridDim.x * blockDim.x) {
        float v = vox[k];
        if (norm) v = fdiv(fmul(v != 0.f ? 1.f : 0.f, fsub(v, mean)), den);
        if (cfg.voxel_bf16) ((__nv_bfloat16*)sq.voxel)[k] = __float2bfloat16_rn(v);
        else ((float*)sq.voxel)[k] = v;
    }
}

// reset the per-sequence work area for the next pack (the workspace is "left zeroed")
__global__ void k_end_pack(const __grid_constant__ StepParams prm) {
    const int s = blockIdx.x;
    StepWork& wk = prm.work[s];
    const SeqDev& sq = prm.seq[s];
    if (threadIdx.x == 0 && prm.cfg.output_mode == V2E_OUT_RAW) wk.raw_count = 0ull;
    for (int i = threadIdx.x; i < V2E_MAX_FRAMES + 4; i += blockDim.x) wk.bar[i] = 0ull;
    (void)sq;
}

template <int FDT, int R>
static int run_stepwise_t(const CfgDev& cfg, const v2e_pack* packs, int n_seq, void* workspace,
                          cudaStream_t st) {
    StepParams prm;
    memset(&prm, 0, sizeof(prm));
    prm.cfg = cfg;
    prm.n_seq = n_seq;
    const StepLayout lay = step_layout(cfg, n_seq);
    char* ws = (char*)workspace;
    prm.work = (StepWork*)(ws + lay.work_off);
    prm.vox_acc = (float*)(ws + lay.vox_off);
    prm.cntpol = (int*)(ws + lay.cntpol_off);
    int T = 0;
    for (int s = 0; s < n_seq; ++s) {
        fill_seq(prm.seq[s], packs[s]);
        T = packs[s].sc.num_frames > T ? packs[s].sc.num_frames : T;
    }
    const int P = cfg.P;
    const int G4 = (P + 3) / 4;
    dim3 grid_px((unsigned)((P + kThreads - 1) / kThreads), (unsigned)n_seq);
    dim3 grid_g4((unsigned)((G4 + kThreads - 1) / kThreads), (unsigned)n_seq);
    k_begin_pack<FDT><<<grid_px, kThreads, 0, st>>>(prm);
    count_launch();
    for (int n = 1; n < T; ++n) {
        prm.n = n;
        for (int s = 0; s < n_seq; ++s) {
            const v2e_frame_scalars& f = packs[s].sc.fr[n];
            prm.fr[s] = make_frame_sc(cfg, f);
        }
        // sequences with fewer frames than T are not supported in one launch (checked by caller)
        k_phase_a<FDT, R><<<grid_g4, kThreads, 0, st>>>(prm);
        k_phase_b<FDT, R><<<grid_px, kThreads, 0, st>>>(prm);
        count_launch();
        count_launch();
    }
    if (cfg.output_mode != V2E_OUT_RAW) {
        const int nparts = lay.n_partials;
        dim3 grid_r((unsigned)nparts, (unsigned)n_seq);
        k_vox_partials<<<grid_r, kThreads, 0, st>>>(prm);
        k_vox_finish<<<grid_r, kThreads, 0, st>>>(prm, nparts);
        count_launch();
        count_launch();
    }
    k_end_pack<<<n_seq, 32, 0, st>>>(prm);
    count_launch();
    return cuda_ok(cudaGetLastError());
}

int run_stepwise(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st) {
#define V2E_DISPATCH(FDT)                                                              \
    return rounds == 7 ? run_stepwise_t<FDT, 7>(cfg, packs, n_seq, workspace, st)      \
                       : run_stepwise_t<FDT, 10>(cfg, packs, n_seq, workspace, st)
    switch (frame_dtype) {
        case V2E_DT_U8: V2E_DISPATCH(V2E_DT_U8);
        case V2E_DT_BF16: V2E_DISPATCH(V2E_DT_BF16);
        default: V2E_DISPATCH(V2E_DT_F32);
    }
#undef V2E_DISPATCH
}

}  // namespace v2e
