// v2e_fixpoint.cu — the "fixed-point" engine: ONE cooperative launch per pack, and the per-frame
// dependency chain of the reference replaced by speculate-and-verify.
//
// The only cross-pixel coupling of EventEmulator.forward is the image-wide max event count of
// each frame (v2e/v2e_model.py:405-406): it sets the event time step of that frame (:416-422),
// hence through the refractory gate (:469-472) the events a pixel really emits and its memorised
// brightness (:522).  The lock-step engines pay one grid-wide reduction per frame (T-1 dependent
// barriers per pack).  Here the vector of maxima g[1..T-1] is GUESSED (last pack's maxima), and
// given a guess every pixel is independent of every other for the whole pack:
//
//   pass 1 (stream)   every pixel is time-stepped through all frames at memory speed.  ~75 % of
//                     the pixels stay quiet for the whole pack: nothing they do depends on g,
//                     they are finished here (state written, zero voxel column).  A pixel that
//                     crosses a threshold (or is a shot-noise candidate) at frame f is snapshot
//                     (state before f, the frame bytes, its thresholds) into a compact entry in
//                     shared memory; its low-pass state, which no event ever touches, is still
//                     carried to the end of the pack and written here.
//   pass 2.. (resolve) the active entries are re-run from their snapshot under the current guess,
//                     one entry per lane, no cross-lane traffic; the pass produces the maxima m
//                     the guess implies, the voxel columns, the event count and the statistics
//                     for the normalisation.  One grid barrier publishes (m, statistics).
//                     m == g  =>  g was the true chain (induction over frames: m[1] never
//                     depends on g; if g[1..k] is true, everything up to frame k is the
//                     reference's computation, so m[k+1] is true) and the pass's results are the
//                     reference's.  Otherwise g := m and the entries are resolved again; at most
//                     T-1 passes, two on typical video (scripts/fixpoint_probe.py).
//   commit            normalise (utils/event_process.py:157-177) and scatter the entries' voxel
//                     columns and state.
//
// Statistics are accumulated as exact integers (voxel values are multiples of 2^-33 in all
// practical cases: fixed point, 64-bit sum and 128-bit sum of squares), so they do not depend
// on the order in which lanes, warps and tiles add them: results are run-to-run deterministic.
//
// Entries that do not fit the shared-memory table (more than `cap` active pixels in a tile,
// non-integer frame values, packs longer than 11 frames) are "spilled": only their index is
// kept, every pass re-reads their inputs from global memory and keeps their results in the
// workspace until commit.  Slower, same results.
#include <cstdlib>

#include "v2e_engine.cuh"

namespace v2e {

#ifndef V2E_FX_THREADS
#define V2E_FX_THREADS 512
#endif
constexpr int kFxThreads = V2E_FX_THREADS;
constexpr int kFxWarps = kFxThreads / 32;
constexpr int kFxChunk = 10;        // frames per unrolled chunk = frames per Philox block
constexpr int kFxHeader = 4096;     // LUT, frame tables, reduction scratch
constexpr int kFxSpin = 1 << 25;    // barrier poll limit (a lost CTA must not hang the GPU)

// entry words (structure of arrays: word w of entry e lives at s_ent[w * cap + e])
enum {
    FX_META = 0,   // [4:0] first active frame f, [5] timestamp_mem == 0 (else -Tr), [31:8] tile-local pixel
    FX_BASE, FX_LP, FX_POS, FX_NEG, FX_NLK,
    FX_F0, FX_F1, FX_F2,   // frame bytes 1..10 (frame n in byte n-1)
    FX_OB, FX_OT,          // results: base_log_frame, timestamp_mem
    FX_V0                  // results: num_bins voxel accumulators
};

struct FxParams {
    CfgDev cfg;
    int n_seq;
    int tiles_per_seq;
    int tile_px;       // pixels per tile, multiple of 32
    int cap;           // compact entries per tile, multiple of 32
    FxWork* fx;        // [n_seq]
    unsigned* spill_list;  // [n_seq][P]   tile-local indices of spilled pixels, per tile region
    float* spill_vox;      // [n_seq][nb][P] voxel accumulators of spilled pixels
    float* spill_state;    // [n_seq][3][P]  base, lp, tmem results of spilled pixels
    long long* timeline;   // development aid (V2E_TIMELINE=1): [cta][32] globaltimer stamps
    SeqDev seq[V2E_MAX_SEQ_PER_LAUNCH];
    FrameSc fr[V2E_MAX_SEQ_PER_LAUNCH][V2E_MAX_FRAMES];
};

// ---- memory-ordering helpers ---------------------------------------------------------------------
__device__ __forceinline__ void fx_st_release(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned fx_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned fx_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ long long fx_now() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FX_STAMP(i) do { if (tl && tid == 0) tl[(i)] = fx_now(); } while (0)

// ---- per-pixel noise stream of this engine --------------------------------------------------------
// Block c covers frames 10c+1 .. 10c+10 of pixel p with two Philox calls:
//   A: 4 words -> 8 normals (Box-Muller pairs), frames +1..+8
//   B: .x .y .z -> 12 bytes, first stage of the shot test of frames +1..+10 (+2 spare);
//      .w -> normals of frames +9, +10
// (The lock-step engines index their stage-0 stream by (4-pixel group, frame); a pixel that is
// resolved on its own would waste 3/4 of every call.  The second-stage streams are shared.)
template <int R>
__device__ __forceinline__ uint4 noise_px(unsigned long long ctr64, unsigned p, unsigned c,
                                          unsigned which) {
    return philox4x32<R>(
        make_uint4(p, c | ((0x80u + which) << 8), (unsigned)ctr64, (unsigned)(ctr64 >> 32)));
}

// exact integer statistics of one voxel value: v = k * 2^-33 (truncated if finer, never in practice)
__device__ __forceinline__ void fx_stat(float v, unsigned& nnz, long long& s1,
                                        unsigned long long& lo, unsigned long long& hi) {
    if (v != 0.f) {
        nnz += 1;
        const long long k = __float2ll_rz(v * 8589934592.0f);
        s1 += k;
        const unsigned long long a = (unsigned long long)(k < 0 ? -k : k);
        const unsigned long long plo = a * a;
        lo += plo;
        hi += __umul64hi(a, a) + (lo < plo ? 1ull : 0ull);
    }
}

// ---- the iteration loop of one pixel-frame, out of line (only ~12 % of the entry-frames) ----------
// returns (final << 32) | bits(timestamp_mem)
template <int R>
static __device__ __noinline__ unsigned long long fx_events(
    const CfgDev* cfg, const SeqDev* sq, const FrameSc* fs, int mx, float ts_step, int gate,
    int n, int p, float pol, float C, int cnt, float I, float tm, int mode, float u, float* vp,
    int vstride) {
    FrameGlobals g;
    g.mx = mx;
    g.ts_step = ts_step;
    g.gate = gate;
    const int fin = iterate_pixel<R>(
        *cfg, *sq, *fs, g, n, p, pol, C, cnt, I, tm, mode, u,
        [&](int b, float v) { vp[(size_t)b * vstride] = fadd(vp[(size_t)b * vstride], v); },
        [&](float, float) {});
    return ((unsigned long long)(unsigned)fin << 32) | (unsigned long long)__float_as_uint(tm);
}

struct FxAcc {   // per-lane accumulators of a resolve pass
    unsigned nnz, nev;
    long long s1;
    unsigned long long lo, hi;
};

// ---- resolve one pixel from frame f under the current guess ---------------------------------------
// getLI(k) -> (lin_log, inten01) of frame k+1; vp/vstride: the pixel's voxel accumulators.
// All lanes of the warp call this together (the per-frame max is a warp reduction); `on` masks
// lanes without an entry.
template <int FDT, int R, bool GENERIC, class GetLI>
__device__ __forceinline__ void fx_resolve(const FxParams& prm, const SeqDev& sq, bool on, int p,
                                           int f, float& base, float& lp, float& tm, float pos,
                                           float neg, float nlk, bool small, bool leak, bool shot,
                                           bool philox, bool cutoff, const FrameSc* s_fr,
                                           const FrameGlobals* s_g, unsigned* s_max, float* vp,
                                           int vstride, FxAcc& acc, GetLI getLI) {
    const CfgDev& cfg = prm.cfg;
    const int T = sq.T, P = cfg.P;
    const int lane = threadIdx.x & 31;
    const int nchunks = (T - 1 + kFxChunk - 1) / kFxChunk;
    for (int c = 0; c < nchunks; ++c) {
        const int n0 = 1 + c * kFxChunk;
        uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
        if (philox && on) {
            if (leak) wa = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 0);
            if (shot || (leak && T - n0 > 8)) wb = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 1);
        }
        float z0 = 0.f, z1 = 0.f;
#pragma unroll
        for (int k = 0; k < kFxChunk; ++k) {
            const int n = n0 + k;
            if (n >= T) break;
            const FrameSc& fs = s_fr[n];
            const bool act = on && n >= f;
            int cnt = 0;
            float pol = 0.f, C = 0.f, I = 0.f, z = 0.f;
            if (leak && philox && (k & 1) == 0) box_muller16(k < 8 ? (&wa.x)[k >> 1] : wb.w, z0, z1);
            if (act) {
                const float2 li = getLI(n);
                I = li.y;
                float nw = li.x;
                if (cutoff) {
                    lp = lowpass(lp, li.x, I, small ? fs.s_small : fs.s_large);
                    nw = lp;
                }
                if (leak) {
                    z = philox ? ((k & 1) ? z1 : z0) : __ldg(sq.leak_randn + (size_t)(n - 1) * P + p);
                    base = leak_sub(base, nlk, cfg.leak_jitter, z, fs.sdt, pos);
                }
                const PixelA a = diff_count(nw, base, pos, neg);
                cnt = a.cnt;
                pol = a.pol;
                C = a.C;
            }
            const int wm = __reduce_max_sync(0xffffffffu, cnt);
            if (lane == 0 && wm > 0) atomicMax(&s_max[n], (unsigned)wm);
            if (act) {
                const FrameGlobals g = s_g[n];
                const bool live = pol != 0.f;
                bool cand = false;
                float u = 0.f;
                if (shot && live) {
                    if (philox) {
                        const float thr256 = shot_thr256(pol > 0.f ? fs.kp : fs.kn, cfg.shot_inten_m1, I, C);
                        const unsigned byte = ((&wb.x)[k >> 2] >> (8 * (k & 3))) & 0xffu;
                        if ((float)byte < thr256) {
                            u = shot_uniform<R>(sq.philox_ctr, n, p, byte);
                            cand = u * 256.0f < thr256;
                        }
                    } else {
                        cand = true;   // injected draws: every pol != 0 pixel walks the iterations
                    }
                }
                int fin = 0;
                if (GENERIC && sq.dbg_shot)
                    for (int i = 0; i < sq.dbg_shot_imax; ++i)
                        sq.dbg_shot[((size_t)(n - 1) * sq.dbg_shot_imax + i) * P + p] = 0;
                if (cnt > 0 || cand) {
                    const int mode = (cand && g.mx > 0) ? (philox ? 2 : 1) : 0;
                    const unsigned long long r = fx_events<R>(
                        &prm.cfg, &sq, &fs, g.mx, g.ts_step, g.gate, n, p, pol, C, min(cnt, g.mx),
                        I, tm, mode, u, vp, vstride);
                    fin = (int)(r >> 32);
                    tm = __uint_as_float((unsigned)r);
                    if (fin) {
                        acc.nev += (unsigned)fin;
                        base = fadd(base, fmul(fmul(pol, (float)fin), C));   // v2e_model.py:522
                    }
                }
                if (GENERIC) {
                    const size_t o = (size_t)(n - 1) * P + p;
                    if (sq.dbg_pol) sq.dbg_pol[o] = (signed char)pol;
                    if (sq.dbg_cnt) sq.dbg_cnt[o] = cnt;
                    if (sq.dbg_final) sq.dbg_final[o] = fin;
                    if (sq.dbg_leak_randn && leak && philox) sq.dbg_leak_randn[o] = z;
                }
            }
        }
    }
    if (on)
        for (int b = 0; b < cfg.nb; ++b) fx_stat(vp[(size_t)b * vstride], acc.nnz, acc.s1, acc.lo, acc.hi);
}

__device__ __forceinline__ void fx_store_vox(const CfgDev& cfg, void* voxel, size_t o, float v) {
    if (cfg.voxel_bf16) ((__nv_bfloat16*)voxel)[o] = __float2bfloat16_rn(v);
    else ((float*)voxel)[o] = v;
}

// NOISE: 0 = no noise, 1 = PHILOX leak + shot (the V2E2VNet wiring), 2 = generic (every feature
// decided at run time, debug maps).  CUTOFF is ignored by the generic variant.
template <int FDT, int R, int NOISE, bool CUTOFF>
__global__ void __launch_bounds__(kFxThreads, 1) k_fixpoint(const __grid_constant__ FxParams prm) {
    constexpr bool GENERIC = NOISE == 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CfgDev& cfg = prm.cfg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    const SeqDev& sq = prm.seq[seq];
    FxWork& fx = prm.fx[seq];
    const int P = cfg.P, nb = cfg.nb, T = sq.T;
    const bool leak = GENERIC ? (cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool shot = GENERIC ? (cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool philox = GENERIC ? (cfg.noise_mode == V2E_NOISE_PHILOX) : (NOISE == 1);
    const bool cutoff = GENERIC ? (cfg.cutoff_on != 0) : CUTOFF;
    const bool normalise = cfg.output_mode == V2E_OUT_VOXEL;
    const int cap = prm.cap;

    // ---- shared memory carve-up ----------------------------------------------------------------
    float2* s_lut = (float2*)smem_raw;                                   // 2048
    FrameSc* s_fr = (FrameSc*)(smem_raw + 2048);                         // 1024
    FrameGlobals* s_g = (FrameGlobals*)(smem_raw + 3072);                // 32 * 12 = 384
    unsigned* s_max = (unsigned*)(smem_raw + 3456);                      // 128: this tile, this pass
    unsigned* s_gmax = (unsigned*)(smem_raw + 3584);                     // 128: all tiles
    unsigned long long* s_acc = (unsigned long long*)(smem_raw + 3712);  // 5 * 8: tile partials
    unsigned long long* s_tot = (unsigned long long*)(smem_raw + 3752);  // 5 * 8: sequence totals
    int* s_cnt = (int*)(smem_raw + 3792);                                // n_entries, n_spill, error
    float* s_stat = (float*)(smem_raw + 3808);                           // mean, den, on
    unsigned* s_ent = (unsigned*)(smem_raw + kFxHeader);

    const int tile_px0 = tile * prm.tile_px;
    const int tile_n = max(0, min(prm.tile_px, P - tile_px0));
    unsigned* spill_list = prm.spill_list + (size_t)seq * P + tile_px0;
    float* spill_vox = prm.spill_vox + (size_t)seq * nb * P;
    float* spill_state = prm.spill_state + (size_t)seq * 3 * P;

    long long* tl = prm.timeline ? prm.timeline + (size_t)blockIdx.x * 32 : nullptr;
    FX_STAMP(0);
    load_lut(s_lut, cfg.lut);
    if (tid < T) s_fr[tid] = prm.fr[seq][tid];
    if (tid < V2E_MAX_FRAMES) {
        // the guess: last pack's maxima (a hint kept in the workspace; any value is correct)
        const unsigned h = fx.guess[tid];
        s_g[tid] = frame_globals(h ? (int)h - 1 : 8, sq.duration, sq.Tr);
        s_max[tid] = 0;
        s_gmax[tid] = 0;
    }
    if (tid < 5) { s_acc[tid] = 0ull; s_tot[tid] = 0ull; }
    if (tid < 3) s_cnt[tid] = 0;
    __syncthreads();

    FX_STAMP(1);
    // ================================ pass 1: stream ===============================================
    const bool compact_ok = T - 1 <= kFxChunk;
    for (int l0 = warp * 32; l0 < tile_n; l0 += kFxThreads) {
        const int l = l0 + lane;
        const bool valid = l < tile_n;
        const int p = tile_px0 + (valid ? l : 0);
        float base = 0.f, lp = 0.f, tm = 0.f, pos = 1.f, neg = 1.f, nlk = 0.f;
        bool small = false;
        int f = 0;
        float sb = 0.f, slp = 0.f;
        unsigned fb0 = 0, fb1 = 0, fb2 = 0;
        bool allint = true;
        if (valid) {
            pos = __ldg(sq.pos + p);
            neg = __ldg(sq.neg + p);
            if (leak) nlk = fmul(cfg.leak_rate, __ldg(sq.nra + p));
            small = is_small_px(p, cfg.W);
            if (sq.first_pack) {
                // _init (v2e_model.py:156-174,239-240)
                float L0, I0;
                log_inten(s_lut, load_frame1<FDT>(sq.frames, p), L0, I0);
                base = L0;
                lp = L0;
                tm = fadd(0.f, -sq.Tr);
            } else {
                base = sq.base[p];
                if (cutoff) lp = sq.lp[p];
                tm = sq.tmem[p];
                // pack rollover (v2e_model.py:329-330)
                if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
                if (tm < 0.f) tm = -sq.Tr;
            }
            const int nchunks = (T - 1 + kFxChunk - 1) / kFxChunk;
            for (int c = 0; c < nchunks; ++c) {
                const int n0 = 1 + c * kFxChunk;
                float x[kFxChunk];
#pragma unroll
                for (int k = 0; k < kFxChunk; ++k)
                    x[k] = n0 + k < T ? load_frame1<FDT>(sq.frames, (long long)(n0 + k) * P + p) : 0.f;
                uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
                if (philox) {
                    if (leak) wa = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 0);
                    if (shot || (leak && T - n0 > 8)) wb = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 1);
                }
                float z0 = 0.f, z1 = 0.f;
#pragma unroll
                for (int k = 0; k < kFxChunk; ++k) {
                    const int n = n0 + k;
                    if (n >= T) break;
                    const FrameSc& fs = s_fr[n];
                    const int v = __float2int_rn(x[k]);
                    const bool isint = (float)v == x[k] && (unsigned)v < 256u;
                    float2 li;
                    if (__builtin_expect(isint, 1)) li = s_lut[v];
                    else li = log_inten_slow(x[k]);
                    allint = allint && isint;
                    if (c == 0) {
                        const unsigned sh = (unsigned)(v & 0xff) << (8 * (k & 3));
                        if (k < 4) fb0 |= sh;
                        else if (k < 8) fb1 |= sh;
                        else fb2 |= sh;
                    }
                    const float lp_old = lp;
                    float nw = li.x;
                    if (cutoff) {
                        lp = lowpass(lp, li.x, li.y, small ? fs.s_small : fs.s_large);
                        nw = lp;
                    }
                    float z = 0.f;
                    if (leak) {
                        if (philox) {
                            if ((k & 1) == 0) box_muller16(k < 8 ? (&wa.x)[k >> 1] : wb.w, z0, z1);
                            z = (k & 1) ? z1 : z0;
                        } else {
                            z = __ldg(sq.leak_randn + (size_t)(n - 1) * P + p);
                        }
                    }
                    if (!f) {
                        const float b = leak ? leak_sub(base, nlk, cfg.leak_jitter, z, fs.sdt, pos) : base;
                        const float diff = fsub(nw, b);
                        const float ad = fabsf(diff);
                        const bool live = ad > 1e-6f;
                        const bool up = diff > 0.f;
                        const float C = live ? (up ? pos : neg) : 0.f;
                        bool maybe = live && ad >= fadd(C, 1e-9f);
                        if (shot) {
                            if (philox) {
                                const float thr256 = shot_thr256(up ? fs.kp : fs.kn, cfg.shot_inten_m1, li.y, C);
                                const unsigned byte = ((&wb.x)[k >> 2] >> (8 * (k & 3))) & 0xffu;
                                maybe = maybe || (float)byte < thr256;
                            } else {
                                maybe = maybe || live;
                            }
                        }
                        if (maybe) {
                            f = n;
                            sb = base;
                            slp = lp_old;
                        } else {
                            base = b;
                            if (GENERIC) {
                                const size_t o = (size_t)(n - 1) * P + p;
                                if (sq.dbg_pol) sq.dbg_pol[o] = live ? (up ? 1 : -1) : 0;
                                if (sq.dbg_cnt) sq.dbg_cnt[o] = 0;
                                if (sq.dbg_final) sq.dbg_final[o] = 0;
                                if (sq.dbg_leak_randn && leak && philox) sq.dbg_leak_randn[o] = z;
                            }
                        }
                    }
                }
            }
        }
        // ---- classify: quiet -> finished; active -> compact entry or spill -------------------------
        const bool active = valid && f != 0;
        const bool want_compact = active && compact_ok && allint;
        const unsigned cm = __ballot_sync(0xffffffffu, want_compact);
        int slot = -1;
        if (cm) {
            int s0 = 0;
            if (lane == 0) s0 = atomicAdd(&s_cnt[0], __popc(cm));
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            if (want_compact) {
                slot = s0 + __popc(cm & ((1u << lane) - 1u));
                if (slot >= cap) slot = -1;
            }
        }
        const bool spill = active && slot < 0;
        const unsigned sm = __ballot_sync(0xffffffffu, spill);
        if (sm) {
            int s0 = 0;
            if (lane == 0) s0 = atomicAdd(&s_cnt[1], __popc(sm));
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            if (spill) spill_list[s0 + __popc(sm & ((1u << lane) - 1u))] = (unsigned)l;
        }
        if (valid) {
            if (!spill) {
                if (cutoff) sq.lp[p] = lp;   // no event ever touches the low-pass state
                if (!active) {
                    sq.base[p] = base;
                    sq.tmem[p] = tm;
                }
            }
            for (int b = 0; b < nb; ++b) fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, 0.f);
            if (slot >= 0) {
                unsigned* e = s_ent + slot;
                e[FX_META * cap] = (unsigned)f | (tm == 0.f ? 32u : 0u) | ((unsigned)l << 8);
                e[FX_BASE * cap] = __float_as_uint(sb);
                e[FX_LP * cap] = __float_as_uint(slp);
                e[FX_POS * cap] = __float_as_uint(pos);
                e[FX_NEG * cap] = __float_as_uint(neg);
                e[FX_NLK * cap] = __float_as_uint(nlk);
                e[FX_F0 * cap] = fb0;
                e[FX_F1 * cap] = fb1;
                e[FX_F2 * cap] = fb2;
            }
        }
    }
    __syncthreads();
    const int n_ent = min(s_cnt[0], cap);
    const int n_spill = s_cnt[1];
    FX_STAMP(2);
    if (tl && tid == 0) { tl[28] = n_ent; tl[29] = n_spill; }

    // ================================ resolve passes ================================================
    int pass = 0;
    bool converged = false;
    for (; pass < T + 2; ++pass) {
        FxAcc acc;
        acc.nnz = 0; acc.nev = 0; acc.s1 = 0; acc.lo = 0; acc.hi = 0;
        // compact entries: inputs and voxel accumulators in shared memory
        for (int e0 = warp * 32; e0 < n_ent; e0 += kFxThreads) {
            const int e = e0 + lane;
            const bool on = e < n_ent;
            unsigned* ep = s_ent + (on ? e : 0);
            const unsigned meta = ep[FX_META * cap];
            const int f = on ? (int)(meta & 31u) : V2E_MAX_FRAMES;
            const int l = (int)(meta >> 8);
            const int p = tile_px0 + l;
            float base = __uint_as_float(ep[FX_BASE * cap]);
            float lp = __uint_as_float(ep[FX_LP * cap]);
            float tm = (meta & 32u) ? 0.f : -sq.Tr;
            const float pos = __uint_as_float(ep[FX_POS * cap]);
            const float neg = __uint_as_float(ep[FX_NEG * cap]);
            const float nlk = __uint_as_float(ep[FX_NLK * cap]);
            const unsigned fb0 = ep[FX_F0 * cap], fb1 = ep[FX_F1 * cap], fb2 = ep[FX_F2 * cap];
            float* vp = (float*)(ep + FX_V0 * cap);
            if (on)
                for (int b = 0; b < nb; ++b) vp[b * cap] = 0.f;
            const bool small = is_small_px(p, cfg.W);
            fx_resolve<FDT, R, GENERIC>(
                prm, sq, on, p, f, base, lp, tm, pos, neg, nlk, small, leak, shot, philox, cutoff,
                s_fr, s_g, s_max, vp, cap, acc, [&](int n) {
                    const int k = n - 1;
                    const unsigned w = k < 4 ? fb0 : (k < 8 ? fb1 : fb2);
                    return s_lut[(w >> (8 * (k & 3))) & 0xffu];
                });
            if (on) {
                ep[FX_OB * cap] = __float_as_uint(base);
                ep[FX_OT * cap] = __float_as_uint(tm);
            }
        }
        // spilled entries: inputs re-read from global memory, results kept in the workspace
        for (int e0 = warp * 32; e0 < n_spill; e0 += kFxThreads) {
            const int e = e0 + lane;
            const bool on = e < n_spill;
            const int l = on ? (int)spill_list[e] : 0;
            const int p = tile_px0 + l;
            float base, lp = 0.f, tm;
            const float pos = __ldg(sq.pos + p), neg = __ldg(sq.neg + p);
            const float nlk = leak ? fmul(cfg.leak_rate, __ldg(sq.nra + p)) : 0.f;
            if (sq.first_pack) {
                float L0, I0;
                log_inten(s_lut, load_frame1<FDT>(sq.frames, p), L0, I0);
                base = L0;
                lp = L0;
                tm = fadd(0.f, -sq.Tr);
            } else {
                base = sq.base[p];
                if (cutoff) lp = sq.lp[p];
                tm = sq.tmem[p];
                if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
                if (tm < 0.f) tm = -sq.Tr;
            }
            float* vp = spill_vox + p;
            if (on)
                for (int b = 0; b < nb; ++b) vp[(size_t)b * P] = 0.f;
            const bool small = is_small_px(p, cfg.W);
            fx_resolve<FDT, R, GENERIC>(
                prm, sq, on, p, 1, base, lp, tm, pos, neg, nlk, small, leak, shot, philox, cutoff,
                s_fr, s_g, s_max, vp, P, acc, [&](int n) {
                    float L, I;
                    log_inten(s_lut, load_frame1<FDT>(sq.frames, (long long)n * P + p), L, I);
                    return make_float2(L, I);
                });
            if (on) {
                spill_state[p] = base;
                spill_state[(size_t)P + p] = lp;
                spill_state[(size_t)2 * P + p] = tm;
            }
        }
        // ---- tile partials -> shared memory (integer adds: order-free) ----------------------------
        {
            unsigned long long nnz = acc.nnz, nev = acc.nev, lo = acc.lo, hi = acc.hi;
            long long s1 = acc.s1;
            for (int o = 16; o > 0; o >>= 1) {
                nnz += __shfl_xor_sync(0xffffffffu, nnz, o);
                nev += __shfl_xor_sync(0xffffffffu, nev, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                const unsigned long long olo = __shfl_xor_sync(0xffffffffu, lo, o);
                const unsigned long long ohi = __shfl_xor_sync(0xffffffffu, hi, o);
                lo += olo;
                hi += ohi + (lo < olo ? 1ull : 0ull);
            }
            if (lane == 0 && (nnz | nev)) {
                atomicAdd(&s_acc[0], nnz);
                atomicAdd(&s_acc[1], (unsigned long long)s1);
                const unsigned long long old = atomicAdd(&s_acc[2], lo);
                atomicAdd(&s_acc[3], hi + (old + lo < old ? 1ull : 0ull));
                atomicAdd(&s_acc[4], nev);
            }
        }
        __syncthreads();
        if (pass < 8) FX_STAMP(3 + 3 * pass);
        // ---- grid barrier of this sequence: publish (maxima, partials), gather everybody's --------
        unsigned* slots = &fx.slot[pass & 1][0][0];
        if (warp == 0) {
            unsigned* my = slots + tile * 32;
            unsigned w = 0;
            if (lane < 16) w = min(s_max[2 * lane], 0xffffu) | (min(s_max[2 * lane + 1], 0xffffu) << 16);
            else if (lane < 26) w = ((const unsigned*)s_acc)[lane - 16];
            if (lane < 26) my[lane] = w;
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                fx_st_release(my + 31, (unsigned)pass + 1u);
            }
        }
        if (tid < ((prm.tiles_per_seq + 31) & ~31)) {
            const bool has = tid < prm.tiles_per_seq;
            const unsigned* sl = slots + (has ? tid : 0) * 32;
            unsigned wv[26];
#pragma unroll
            for (int i = 0; i < 26; ++i) wv[i] = 0;
            if (has) {
                int spins = 0;
                while (fx_ld_relaxed(sl + 31) != (unsigned)pass + 1u)
                    if (++spins > kFxSpin) { s_cnt[2] = 1; break; }
                (void)fx_ld_acquire(sl + 31);
#pragma unroll
                for (int i = 0; i < 26; ++i) wv[i] = __ldcg(sl + i);
            }
            const int nw = (T + 1) / 2;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i >= nw) break;
                const unsigned a = __reduce_max_sync(0xffffffffu, wv[i] & 0xffffu);
                const unsigned b = __reduce_max_sync(0xffffffffu, wv[i] >> 16);
                if (lane == 0) {
                    if (a) atomicMax(&s_gmax[2 * i], a);
                    if (b) atomicMax(&s_gmax[2 * i + 1], b);
                }
            }
            unsigned long long q[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) q[i] = (unsigned long long)wv[16 + 2 * i] | ((unsigned long long)wv[17 + 2 * i] << 32);
            for (int o = 16; o > 0; o >>= 1) {
                q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
                q[1] += __shfl_xor_sync(0xffffffffu, q[1], o);
                q[4] += __shfl_xor_sync(0xffffffffu, q[4], o);
                const unsigned long long olo = __shfl_xor_sync(0xffffffffu, q[2], o);
                const unsigned long long ohi = __shfl_xor_sync(0xffffffffu, q[3], o);
                q[2] += olo;
                q[3] += ohi + (q[2] < olo ? 1ull : 0ull);
            }
            if (lane == 0) {
                atomicAdd(&s_tot[0], q[0]);
                atomicAdd(&s_tot[1], q[1]);
                const unsigned long long old = atomicAdd(&s_tot[2], q[2]);
                atomicAdd(&s_tot[3], q[3] + (old + q[2] < old ? 1ull : 0ull));
                atomicAdd(&s_tot[4], q[4]);
            }
        }
        __syncthreads();
        if (pass < 8) FX_STAMP(4 + 3 * pass);
        // ---- m == g ? ------------------------------------------------------------------------------
        const int changed = __syncthreads_or(tid >= 1 && tid < T && (int)s_gmax[tid] != s_g[tid].mx);
        if (!changed || s_cnt[2]) { converged = !changed; break; }
        if (tid < V2E_MAX_FRAMES) {
            s_g[tid] = frame_globals((int)s_gmax[tid], sq.duration, sq.Tr);
            s_max[tid] = 0;
            s_gmax[tid] = 0;
        }
        if (tid < 5) { s_acc[tid] = 0ull; s_tot[tid] = 0ull; }
        __syncthreads();
    }

    FX_STAMP(27);
    if (tl && tid == 0) tl[30] = pass + 1;
    // ================================ commit =========================================================
    if (tid == 0) {
        float mean = 0.f, den = 1.f, on = 0.f;
        const double c = (double)s_tot[0];
        if (c > 0. && normalise) {
            // utils/event_process.py:166-175
            const double a = (double)(long long)s_tot[1] * (1.0 / 8589934592.0);
            const double b = ((double)s_tot[3] * 18446744073709551616.0 + (double)s_tot[2]) *
                             (1.0 / 8589934592.0) * (1.0 / 8589934592.0);
            mean = fdiv((float)a, (float)c);
            den = fadd(sqrtf(fsub(fdiv((float)b, (float)c), fmul(mean, mean))), 1e-8f);
            on = 1.f;
        }
        s_stat[0] = mean; s_stat[1] = den; s_stat[2] = on;
        if (tile == 0) {
            *sq.num_events = s_tot[4];
            if (sq.status && (!converged || s_cnt[2])) *sq.status = V2E_STATUS_NOT_CONVERGED;
        }
    }
    if (tile == 0 && tid >= 1 && tid < V2E_MAX_FRAMES) {
        if (tid < T) {
            if (sq.max_iters) sq.max_iters[tid] = (int)s_gmax[tid];
            fx.guess[tid] = s_gmax[tid] + 1u;
        } else {
            fx.guess[tid] = 0u;
        }
    }
    __syncthreads();
    {
        const float mean = s_stat[0], den = s_stat[1];
        const bool norm = s_stat[2] != 0.f;
        for (int e = tid; e < n_ent; e += kFxThreads) {
            const unsigned* ep = s_ent + e;
            const int p = tile_px0 + (int)(ep[FX_META * cap] >> 8);
            sq.base[p] = __uint_as_float(ep[FX_OB * cap]);
            sq.tmem[p] = __uint_as_float(ep[FX_OT * cap]);
            for (int b = 0; b < nb; ++b) {
                float v = __uint_as_float(ep[(FX_V0 + b) * cap]);
                if (norm) v = fdiv(fmul(v != 0.f ? 1.f : 0.f, fsub(v, mean)), den);
                fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, v);
            }
        }
        for (int e = tid; e < n_spill; e += kFxThreads) {
            const int p = tile_px0 + (int)spill_list[e];
            sq.base[p] = spill_state[p];
            if (cutoff) sq.lp[p] = spill_state[(size_t)P + p];
            sq.tmem[p] = spill_state[(size_t)2 * P + p];
            for (int b = 0; b < nb; ++b) {
                float v = spill_vox[(size_t)b * P + p];
                if (norm) v = fdiv(fmul(v != 0.f ? 1.f : 0.f, fsub(v, mean)), den);
                fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, v);
            }
        }
    }
    FX_STAMP(31);
    // ---- leave the barrier slots zeroed: the last CTA of the sequence to get here cleans up ----------
    if (tid == 0) {
        __threadfence();
        s_cnt[0] = atomicAdd(&fx.exit_count, 1u) == (unsigned)prm.tiles_per_seq - 1u;
    }
    __syncthreads();
    if (s_cnt[0]) {
        unsigned* all = &fx.slot[0][0][0];
        for (int i = tid; i < 2 * kRaMaxTiles * 32; i += kFxThreads) all[i] = 0u;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            fx.exit_count = 0;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct FxPlan {
    int ok;
    int tiles_per_seq, tile_px, cap;
    size_t smem;
};

static int g_fx_sm = -1, g_fx_optin = 0, g_fx_coop = 0;

static int fx_device_props() {
    if (g_fx_sm >= 0) return V2E_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return V2E_ERR_NO_DEVICE;
    int major = 0, sms = 0, optin = 0, coop = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (major < 10) return V2E_ERR_NO_DEVICE;
    g_fx_optin = optin;
    g_fx_coop = coop;
    g_fx_sm = sms;
    return V2E_OK;
}

static FxPlan fx_plan(const CfgDev& cfg, int n_seq) {
    FxPlan pl{};
    if (fx_device_props() != V2E_OK || !g_fx_coop) return pl;
    if (cfg.output_mode == V2E_OUT_RAW) return pl;
    int tiles = g_fx_sm / n_seq;
    if (tiles > kRaMaxTiles) tiles = kRaMaxTiles;
    if (tiles < 1) return pl;
    int tile_px = (cfg.P + tiles - 1) / tiles;
    tile_px = (tile_px + 31) & ~31;
    if (tile_px >= (1 << 24)) return pl;
    tiles = (cfg.P + tile_px - 1) / tile_px;
    const int ewords = FX_V0 + cfg.nb;
    int cap = (int)(((size_t)g_fx_optin - kFxHeader) / (4u * ewords)) & ~31;
    if (cap > tile_px) cap = tile_px;
    if (cap < 32) return pl;
    pl.ok = 1;
    pl.tiles_per_seq = tiles;
    pl.tile_px = tile_px;
    pl.cap = cap;
    pl.smem = kFxHeader + (size_t)cap * ewords * 4;
    return pl;
}

int fixpoint_fits(const CfgDev& cfg, int frame_dtype, int n_seq) {
    (void)frame_dtype;
    const int rc = fx_device_props();
    if (rc != V2E_OK) return rc;
    return fx_plan(cfg, n_seq).ok;
}

template <int FDT, int R, int NOISE, bool CUTOFF>
static int launch_fixpoint(const FxParams& prm, const FxPlan& pl, cudaStream_t st) {
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_fixpoint<FDT, R, NOISE, CUTOFF>;
    int rc = V2E_OK;
    if (!attr_set) {
        rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          g_fx_optin));
        if (rc != V2E_OK) return rc;
        attr_set = true;
    }
    void* args[] = {(void*)&prm};
    const dim3 grid((unsigned)(pl.tiles_per_seq * prm.n_seq));
    rc = cuda_ok(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(kFxThreads), args,
                                             pl.smem, st));
    count_launch();
    return rc;
}

template <int FDT>
static int run_fixpoint_t(const CfgDev& cfg, int rounds, const v2e_pack* packs, int n_seq,
                          void* workspace, cudaStream_t st) {
    const FxPlan pl = fx_plan(cfg, n_seq);
    if (!pl.ok) return V2E_ERR_UNSUPPORTED;
    static thread_local FxParams prm;  // 16 KB: keep it off the stack
    const StepLayout lay = step_layout(cfg, n_seq);
    prm.cfg = cfg;
    prm.n_seq = n_seq;
    prm.tiles_per_seq = pl.tiles_per_seq;
    prm.tile_px = pl.tile_px;
    prm.cap = pl.cap;
    prm.fx = (FxWork*)((char*)workspace + lay.fx_off);
    // the stepwise engine's accumulator areas are unused here: borrow them for the spilled pixels
    prm.spill_vox = (float*)((char*)workspace + lay.vox_off);
    prm.spill_list = (unsigned*)((char*)workspace + lay.cntpol_off);
    prm.spill_state = (float*)((char*)workspace + lay.fxstate_off);
    static const bool want_timeline = getenv("V2E_TIMELINE") != nullptr;
    prm.timeline = want_timeline ? (long long*)((char*)workspace + lay.tl_off) : nullptr;
    bool dbg = false;
    for (int s = 0; s < n_seq; ++s) {
        fill_seq(prm.seq[s], packs[s]);
        for (int n = 1; n < packs[s].sc.num_frames; ++n)
            prm.fr[s][n] = make_frame_sc(cfg, packs[s].sc.fr[n]);
        const v2e_pack& p = packs[s];
        dbg = dbg || p.dbg_pol || p.dbg_cnt || p.dbg_final || p.dbg_leak_randn || p.dbg_shot;
    }
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool philox = cfg.noise_mode == V2E_NOISE_PHILOX;
    int noise = 2;
    if (!dbg) {
        if (!leak && !shot) noise = 0;
        else if (philox && leak && shot) noise = 1;
    }
    const bool r7 = rounds == 7;
    if (noise == 0)
        return cfg.cutoff_on ? launch_fixpoint<FDT, 10, 0, true>(prm, pl, st)
                             : launch_fixpoint<FDT, 10, 0, false>(prm, pl, st);
    if (noise == 1) {
        if (cfg.cutoff_on)
            return r7 ? launch_fixpoint<FDT, 7, 1, true>(prm, pl, st)
                      : launch_fixpoint<FDT, 10, 1, true>(prm, pl, st);
        return r7 ? launch_fixpoint<FDT, 7, 1, false>(prm, pl, st)
                  : launch_fixpoint<FDT, 10, 1, false>(prm, pl, st);
    }
    return r7 ? launch_fixpoint<FDT, 7, 2, false>(prm, pl, st)
              : launch_fixpoint<FDT, 10, 2, false>(prm, pl, st);
}

int run_fixpoint(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st) {
    switch (frame_dtype) {
        case V2E_DT_U8: return run_fixpoint_t<V2E_DT_U8>(cfg, rounds, packs, n_seq, workspace, st);
        case V2E_DT_BF16: return run_fixpoint_t<V2E_DT_BF16>(cfg, rounds, packs, n_seq, workspace, st);
        default: return run_fixpoint_t<V2E_DT_F32>(cfg, rounds, packs, n_seq, workspace, st);
    }
}

}  // namespace v2e
