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
// Statistics are accumulated as exact integers (voxel values are multiples of 2^-31 in all
// practical cases: fixed point, 64-bit sum and 128-bit sum of squares), so they do not depend
// on the order in which lanes, warps and tiles add them: results are run-to-run deterministic.
//
// Entries that do not fit the shared-memory table (more than `cap` active pixels in a tile,
// non-integer frame values, packs longer than 11 frames) are "spilled": only their index is
// kept, every pass re-reads their inputs from global memory and keeps their results in the
// workspace until commit.  Slower, same results.
#include <cstring>

#include <vector>

#include "v2e_fixpoint.cuh"

namespace v2e {

// ---- compact entries: frame-synchronous, one entry per lane ------------------------------------------
// Per frame an entry costs two shared-memory loads, two subtractions, the threshold test and the
// quotient; the event code behind it is reached by the lanes that cross (about one in eight).
template <int FDT, int R, int NOISE, bool SPEC>
static __device__ __noinline__ void fx_resolve_compact(const FxParams& prm, const SeqDev& sq,
                                                       FxSmem sm, unsigned* s_ent, int n_cold,
                                                       int n_hot, int hot_off, int* task_ctr,
                                                       int tile_px0,
                                                       bool leak_rt, bool shot_rt, bool philox_rt) {
    constexpr bool GENERIC = NOISE == 2;
    constexpr bool LEAKC = NOISE == 1;
    const bool leak = GENERIC ? leak_rt : (NOISE == 1);
    const bool shot = GENERIC ? shot_rt : (NOISE == 1);
    const bool philox = GENERIC ? philox_rt : (NOISE == 1);
    const CfgDev& cfg = prm.cfg;
    const int T = SPEC ? kFxSpecT : sq.T, P = cfg.P, nb = SPEC ? kFxSpecNb : cfg.nb;
    const int cap = SPEC ? fx_spec_cap(LEAKC) : prm.cap;
    const int nw_off = SPEC ? FX_V0 + kFxSpecNb : prm.nw_off;
    const int dl_off = SPEC ? FX_V0 + kFxSpecNb + kFxSpecT - 1 : prm.dl_off;
    const int lane = threadIdx.x & 31;
    const float Tr = sq.Tr;
    FxAcc acc;
    acc.nnz = 0; acc.nev = 0; acc.s1 = 0; acc.lo = 0; acc.hi = 0;
    // tasks of 32 entries, drawn from a queue: the hot list first, each list in the order of the
    // first active frame (sm.perm), so that a task only walks the frames its entries need and
    // the long tasks are drawn first
    const int th = (n_hot + 31) >> 5, tc = (n_cold + 31) >> 5;
    while (true) {
        int t = 0;
        if (lane == 0) t = atomicAdd(task_ctr, 1);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= th + tc) break;
        const bool probe = sm.tl && t == 0 && lane == 0;
        if (probe) sm.tl[0] = fx_now();
        int e;
        bool on;
        if (t < th) {
            const int i = t * 32 + lane;
            on = i < n_hot;
            e = sm.perm[hot_off + (on ? i : 0)];
        } else {
            const int i = (t - th) * 32 + lane;
            on = i < n_cold;
            e = sm.perm[on ? i : 0];
        }
        unsigned* ep = s_ent + (on ? e : 0);
        const unsigned meta = ep[FX_META * cap];
        const int p = tile_px0 + (int)ep[FX_IDX * cap];
        const unsigned cm = (meta >> 6) & 0x3ffu;
        float base = __uint_as_float(ep[FX_BASE * cap]);
        float tm = (meta & 32u) ? 0.f : -Tr;
        const float pos = __uint_as_float(ep[FX_POS * cap]);
        const float neg = __uint_as_float(ep[FX_NEG * cap]);
        float* vp = (float*)(ep + FX_V0 * cap);
        const float* nwp = (const float*)(ep + (nw_off - 1) * cap);   // [n] -> frame n
        const float* dlp = (const float*)(ep + (dl_off - 1) * cap);
        const int f = on ? (int)(meta & 31u) : V2E_MAX_FRAMES;
        if (on) {
#pragma unroll
            for (int b = 0; b < (SPEC ? kFxSpecNb : V2E_MAX_BINS); ++b)
                if (b < nb) vp[b * cap] = 0.f;
            fx_dbg_clear_shot<GENERIC>(sq, P, p, f, T);
        }
        if (probe) sm.tl[1] = fx_now();
        const int n_first = __reduce_min_sync(0xffffffffu, f);
#pragma unroll 1
        for (int n = n_first; n < T; ++n) {   // not unrolled: the loop body should stay in the L0 I-cache
            const bool act = n >= f;
            const float b = leak ? fsub(base, dlp[n * cap]) : base;
            const float diff = fsub(nwp[n * cap], b);
            const float ad = fabsf(diff);
            const bool live = act && ad > 1e-6f;   // v2e_model.py:394-395
            const bool up = diff > 0.f;
            const float C = live ? (up ? pos : neg) : 0.f;
            const float den = fadd(C, 1e-9f);
            const bool cross = live && ad >= den;
            if (act) base = b;
            const int cnt = cross ? (int)floorf(fdiv(ad, den)) : 0;   // :404
            if ((unsigned)cnt > sm.max[n]) atomicMax(&sm.max[n], (unsigned)cnt);
            const bool bit = shot && live && ((cm >> (n - 1)) & 1u);
            const float pol = live ? (up ? 1.f : -1.f) : 0.f;
            int fin = 0;
            if (cross || bit) {
                const FrameGlobals g = sm.g[n];
                if (!bit && g.gate) {
                    // The common event frame, inline: gate on (v2e_model.py:469), no shot draw.
                    // Iteration j of [0,c) is accepted iff ts_j - timestamp_mem > Tr (:419,:470),
                    // monotone in j, so the next accepted one is the first j that passes:
                    // j == 0, or pred(j) && !pred(j-1).  The usual outcomes (nothing / j = 0 /
                    // the closed-form estimate) are evaluated side by side without branches;
                    // anything else (estimate off, a second event in the frame) takes the loop.
                    const float t0 = sm.fr[n].t_bin_prev, step = g.ts_step;
                    const int c = min(cnt, g.mx);
                    auto ts_of = [&](int j) { return fadd(t0, fmul(step, (float)(j + 1))); };
                    auto accept = [&](float ts) {
                        tm = ts;
                        fin += 1;
                        // v2e_model.py:478-502
                        const float ti = floorf(ts);
                        const float d = fsub(ts, ti);
                        const int bb = (int)ti;
                        if (bb >= 0 && bb < nb) vp[bb * cap] = fadd(vp[bb * cap], fmul(pol, fsub(1.0f, d)));
                        if (bb >= 0 && bb + 1 < nb) vp[(bb + 1) * cap] = fadd(vp[(bb + 1) * cap], fmul(pol, d));
                    };
                    const float tc = ts_of(c - 1), t1 = ts_of(0);
                    const bool any = c > 0 && fsub(tc, tm) > Tr;      // else none is accepted
                    const bool p_first = fsub(t1, tm) > Tr;
                    int je = (int)floorf((Tr + tm - t0) * __fdividef(1.0f, step)) - 1;   // estimate
                    je = max(min(je, c - 1), 1);
                    const float tje = ts_of(je);
                    const bool ok_est = c >= 2 && fsub(tje, tm) > Tr && !(fsub(ts_of(je - 1), tm) > Tr);
                    int i = 0;
                    bool slow = any;
                    if (any && (p_first || ok_est)) {
                        const float ts = p_first ? t1 : tje;
                        i = (p_first ? 0 : je) + 1;
                        accept(ts);
                        slow = i < c && fsub(tc, ts) > Tr;            // another one fits (rare)
                    }
                    if (slow) {
                        while (i < c) {
                            if (!(fsub(tc, tm) > Tr)) break;          // none left
                            int j = i;
                            if (!(fsub(ts_of(i), tm) > Tr)) {
                                j = (int)floorf((Tr + tm - t0) * __fdividef(1.0f, step)) - 1;
                                j = min(max(j, i + 1), c - 1);
                                while (j < c - 1 && !(fsub(ts_of(j), tm) > Tr)) ++j;
                                while (j > i + 1 && fsub(ts_of(j - 1), tm) > Tr) --j;
                            }
                            accept(ts_of(j));
                            i = j + 1;
                        }
                    }
                } else {
                    const unsigned long long r = fx_event_frame<FDT, R, GENERIC>(
                        prm, sq, sm, n, p, cnt, pol, C, bit, shot, philox, tm, vp, cap);
                    fin = (int)(r >> 32);
                    tm = __uint_as_float((unsigned)r);
                }
                if (fin) {
                    acc.nev += (unsigned)fin;
                    base = fadd(base, fmul(fmul(pol, (float)fin), C));   // v2e_model.py:522
                }
            }
            if (act) fx_dbg<GENERIC>(sq, P, n, p, pol, cnt, fin);
        }
        if (probe) sm.tl[2] = fx_now();
        if (on) {
            ep[FX_OB * cap] = __float_as_uint(base);
            ep[FX_OT * cap] = __float_as_uint(tm);
#pragma unroll
            for (int b = 0; b < (SPEC ? kFxSpecNb : V2E_MAX_BINS); ++b)
                if (b < nb) fx_stat(vp[b * cap], acc.nnz, acc.s1, acc.lo, acc.hi);
        }
        if (probe) sm.tl[3] = fx_now();
    }
    if (sm.tl && lane == 0) atomicMax((unsigned long long*)&sm.tl[4], (unsigned long long)fx_now());
    fx_flush(acc, sm.acc);
    if (sm.tl && lane == 0) atomicMax((unsigned long long*)&sm.tl[5], (unsigned long long)fx_now());
}

// NOISE: 0 = no noise, 1 = PHILOX leak + shot (the V2E2VNet wiring), 2 = generic (every feature
// decided at run time, debug maps).  CUTOFF is ignored by the generic variant.
template <int FDT, int R, int NOISE, bool CUTOFF, bool SPEC>
__global__ void __launch_bounds__(kFxThreads, 1) k_fixpoint(const __grid_constant__ FxParams prm) {
    constexpr bool GENERIC = NOISE == 2;
    static_assert(!(SPEC && GENERIC), "the specialised variants have their noise wiring fixed");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CfgDev& cfg = prm.cfg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    // this sequence's launch record -> shared memory (every later access is an LDS)
    const FxSeqBlock* blk = prm.blocks ? prm.blocks + seq : &prm.one;
    {
        unsigned* d_sq = (unsigned*)(smem_raw + 3840);
        const unsigned* s_sq = (const unsigned*)&blk->sq;
        for (int i = tid; i < (int)(sizeof(SeqDev) / 4); i += kFxThreads) d_sq[i] = s_sq[i];
        unsigned* d_fr = (unsigned*)(smem_raw + 2048);
        const unsigned* s_fr0 = (const unsigned*)&blk->fr[0];
        for (int i = tid; i < (int)(sizeof(FrameSc) * V2E_MAX_FRAMES / 4); i += kFxThreads) d_fr[i] = s_fr0[i];
    }
    __syncthreads();
    const SeqDev& sq = *(const SeqDev*)(smem_raw + 3840);
    FxWork& fx = prm.fx[seq];
    const int P = cfg.P, nb = SPEC ? kFxSpecNb : cfg.nb, T = SPEC ? kFxSpecT : sq.T;
    const bool leak = GENERIC ? (cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool shot = GENERIC ? (cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool philox = GENERIC ? (cfg.noise_mode == V2E_NOISE_PHILOX) : (NOISE == 1);
    const bool cutoff = GENERIC ? (cfg.cutoff_on != 0) : CUTOFF;
    const bool normalise = cfg.output_mode == V2E_OUT_VOXEL;
    const int cap = SPEC ? fx_spec_cap(NOISE == 1) : prm.cap;
    const int nw_off = SPEC ? FX_V0 + kFxSpecNb : prm.nw_off;
    const int dl_off = SPEC ? FX_V0 + kFxSpecNb + kFxSpecT - 1 : prm.dl_off;

    // ---- shared memory carve-up ----------------------------------------------------------------
    float2* s_lut = (float2*)smem_raw;                                   // 2048
    FrameSc* s_fr = (FrameSc*)(smem_raw + 2048);                         // 1024
    FrameGlobals* s_g = (FrameGlobals*)(smem_raw + 3072);                // 32 * 12 = 384
    unsigned* s_max = (unsigned*)(smem_raw + 3456);                      // 128: this tile, this pass
    unsigned* s_gmax = (unsigned*)(smem_raw + 3584);                     // 128: all tiles
    unsigned long long* s_acc = (unsigned long long*)(smem_raw + 3712);  // 5 * 8: tile partials
    unsigned long long* s_tot = (unsigned long long*)(smem_raw + 3752);  // 5 * 8: sequence totals
    int* s_cnt = (int*)(smem_raw + 3792);   // 8: [0] entries, [1] spilled, [2] error, [5] task counter
    float* s_stat = (float*)(smem_raw + 3824);                           // mean, den, on
    unsigned short* s_perm = (unsigned short*)(smem_raw + kFxPermOff);
    int* s_hist = (int*)(smem_raw + kFxPermOff + 2 * kFxPermMax);   // 2 x 64 ints
    static_assert(kFxPermOff + 2 * kFxPermMax + 512 <= kFxHeader, "header layout");
    unsigned* s_ent = (unsigned*)(smem_raw + kFxHeader);
    FxSmem sm;
    sm.lut = s_lut; sm.fr = s_fr; sm.g = s_g; sm.max = s_max; sm.acc = s_acc; sm.perm = s_perm; sm.tl = nullptr;

    const int tile_px0 = tile * prm.tile_px;
    const int tile_n = max(0, min(prm.tile_px, P - tile_px0));
    unsigned* spill_list = prm.spill_list + (size_t)seq * P + tile_px0;
    float* spill_vox = prm.spill_vox + (size_t)seq * nb * P;
    float* spill_state = prm.spill_state + (size_t)seq * 3 * P;

    long long* tl = prm.timeline ? prm.timeline + (size_t)blockIdx.x * 32 : nullptr;
    FX_STAMP(0);
    if (prm.timeline && blockIdx.x == 0 && tid == 0)
        prm.timeline[160 * 32 + 2 * (fx.pad[0] & 127u)] = fx_now();
    load_lut(s_lut, cfg.lut);
    if (tid < V2E_MAX_FRAMES) {
        // the guess: last pack's maxima (a hint kept in the workspace; any value is correct)
        const unsigned h = fx.guess[tid];
        s_g[tid] = frame_globals(h ? (int)h - 1 : 8, sq.duration, sq.Tr);
        s_max[tid] = 0;
        s_gmax[tid] = 0;
    }
    if (tid < 5) { s_acc[tid] = 0ull; s_tot[tid] = 0ull; }
    if (tid < 8) s_cnt[tid] = 0;
    if (tid < 128) s_hist[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        int gmin = 0x7fffffff;
        for (int n = 1; n < T; ++n) gmin = min(gmin, s_g[n].mx);
        s_stat[3] = (float)max(2, gmin / 2);   // theta of the "hot" test
    }
    __syncthreads();
    const float theta = s_stat[3];

    FX_STAMP(1);
    // ================================ pass 1: stream ===============================================
    // One pixel per lane, all frames of the pack in one go; the loads of the lane's next pixel are
    // issued before the current one is computed.  Lanes past the end of the tile shadow its last
    // pixel (the warp stays converged for the ballots) and write nothing.
    const bool compact_ok = T - 1 <= kFxChunk;
    const float jit = cfg.leak_jitter;
    struct FxIn {
        float pos, neg, nra, base, lp, tm;
        float x[kFxChunk];
    };
    auto load_in = [&](int l0, FxIn& in) {
        const int p = tile_px0 + min(l0 + lane, tile_n - 1);
        in.pos = __ldg(sq.pos + p);
        in.neg = __ldg(sq.neg + p);
        in.nra = leak ? __ldg(sq.nra + p) : 0.f;
        if (sq.first_pack) {
            in.base = load_frame1<FDT>(sq.frames, p);
            in.lp = 0.f;
            in.tm = 0.f;
        } else {
            in.base = sq.base[p];
            in.lp = cutoff ? sq.lp[p] : 0.f;
            in.tm = sq.tmem[p];
        }
#pragma unroll
        for (int k = 0; k < kFxChunk; ++k)
            in.x[k] = 1 + k < T ? load_frame1<FDT>(sq.frames, (long long)(1 + k) * P + p) : 0.f;
    };
    FxIn nxt;
    if (warp * 32 < tile_n) load_in(warp * 32, nxt);
    for (int l0 = warp * 32; l0 < tile_n; l0 += kFxThreads) {
        const FxIn in = nxt;
        if (l0 + kFxThreads < tile_n) load_in(l0 + kFxThreads, nxt);
        const int l = l0 + lane;
        const bool valid = l < tile_n;
        const int p = tile_px0 + min(l, tile_n - 1);
        const float pos = in.pos, neg = in.neg;
        const float nlk = leak ? fmul(cfg.leak_rate, in.nra) : 0.f;
        const float cmax = fmaxf(pos, neg);
        const bool small = is_small_px(p, cfg.W);
        // a frame value that is not an integer in 0..255 sends the pixel to the spill path
        // (fp64 lin_log there); this loop then only ever indexes the table
        bool bad = false;
        float base, lp = 0.f, tm;
        if (sq.first_pack) {
            // _init (v2e_model.py:156-174,239-240)
            const int v = __float2int_rn(in.base);
            bad = !((float)v == in.base && (unsigned)v < 256u);
            base = s_lut[v & 255].x;
            lp = base;
            tm = fadd(0.f, -sq.Tr);
        } else {
            base = in.base;
            lp = in.lp;
            tm = in.tm;
            // pack rollover (v2e_model.py:329-330)
            if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
            if (tm < 0.f) tm = -sq.Tr;
        }
        int f = 0;
        unsigned cm = 0;
        float sb = 0.f;
        bool hot = false;
        const float base0 = base;
        float nwv[kFxChunk], dlv[kFxChunk];
        const int nchunks = (T - 1 + kFxChunk - 1) / kFxChunk;
        for (int c = 0; c < nchunks; ++c) {
            const int n0 = 1 + c * kFxChunk;
            float x[kFxChunk];
#pragma unroll
            for (int k = 0; k < kFxChunk; ++k) {
                x[k] = in.x[k];
                if (c > 0) x[k] = n0 + k < T ? load_frame1<FDT>(sq.frames, (long long)(n0 + k) * P + p) : 0.f;
            }
            uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
            if (philox) {
                if (leak) wa = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 0);
                if (shot || (leak && T - n0 > 8)) wb = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 1);
            }
            // (A) what the frames present to the pixel, whatever it does: straight-line code, the
            // only chain across frames is the low-pass state (two dependent operations per frame)
            float z0 = 0.f, z1 = 0.f;
            cm = 0;
#pragma unroll
            for (int k = 0; k < kFxChunk; ++k) {
                const int n = n0 + k;
                nwv[k] = 0.f;
                dlv[k] = 0.f;
                if (n >= T) break;
                const FrameSc& fs = s_fr[n];
                const int v = __float2int_rn(x[k]);
                bad = bad || !((float)v == x[k] && (unsigned)v < 256u);
                const float2 li = s_lut[v & 255];
                nwv[k] = li.x;
                if (cutoff) {
                    lp = lowpass(lp, li.x, li.y, small ? fs.s_small : fs.s_large);
                    nwv[k] = lp;
                }
                if (leak) {
                    float z;
                    if (philox) {
                        if ((k & 1) == 0) box_muller16(k < 8 ? (&wa.x)[k >> 1] : wb.w, z0, z1);
                        z = (k & 1) ? z1 : z0;
                    } else {
                        z = __ldg(sq.leak_randn + (size_t)(n - 1) * P + p);
                    }
                    dlv[k] = leak_dec(nlk, jit, z, fs.sdt, pos);
                    if (GENERIC && sq.dbg_leak_randn && philox && valid)
                        sq.dbg_leak_randn[(size_t)(n - 1) * P + p] = z;
                }
                // may this frame hold a shot event?  Conservative first stage: the byte against
                // the largest threshold either polarity could give (exact test in the resolve pass)
                if (shot) {
                    if (philox) {
                        const unsigned field = fx_shot_field(wb, k);
                        const float tmax = fx_shot_thr(fmaxf(fs.kp, fs.kn), cfg.shot_inten_m1, li.y, cmax);
                        // second stage too (1 pixel-frame in ~256 gets here): the same uniform the
                        // resolve pass will draw, against the same upper bound
                        if ((float)field < tmax &&
                            fx_shot_uniform<R>(sq.philox_ctr, n, p, field) * 1024.0f < tmax)
                            cm |= 1u << k;
                    } else {
                        cm |= 1u << k;
                    }
                }
            }
            // (B) the memorised brightness up to the first frame that crosses a threshold
#pragma unroll
            for (int k = 0; k < kFxChunk; ++k) {
                const int n = n0 + k;
                if (n >= T) break;
                const float b = leak ? fsub(base, dlv[k]) : base;
                const float diff = fsub(nwv[k], b);
                const float ad = fabsf(diff);
                const bool live = ad > 1e-6f;
                const bool up = diff > 0.f;
                const float C = live ? (up ? pos : neg) : 0.f;
                const float den = fadd(C, 1e-9f);
                const bool hit = !f && valid && live && (ad >= den || ((cm >> k) & 1u));
                if (hit) {
                    f = n;
                    sb = base;
                }
                // "hot": some frame is theta thresholds away from the brightness the pack starts
                // with - a bound no event count of the pixel exceeds by much.  Scheduling hint:
                // the hot entries alone usually decide the per-frame maxima (pre-pass below).
                const float d0 = fsub(nwv[k], base0);
                hot = hot || fabsf(d0) >= theta * (d0 > 0.f ? pos : neg);
                if (!f) {
                    base = b;
                    if (valid) fx_dbg<GENERIC>(sq, P, n, p, live ? (up ? 1.f : -1.f) : 0.f, 0, 0);
                }
            }
        }
        // ---- quiet -> finished; active -> compact entry or spill ------------------------------------
        // cold entries fill the table from the bottom, hot ones from the top
        int slot = -1;
        const bool want = f != 0 && compact_ok && !bad;
        const unsigned am = __ballot_sync(0xffffffffu, want);
        if (am) {
            const unsigned hm = __ballot_sync(0xffffffffu, want && hot);
            const unsigned lm = am & ~hm;
            int s0 = 0, sl = 0, sh = 0;
            if (lane == 0) {
                s0 = atomicAdd(&s_cnt[0], __popc(am));
                if (s0 + __popc(am) <= cap) {
                    if (lm) sl = atomicAdd(&s_cnt[3], __popc(lm));
                    if (hm) sh = atomicAdd(&s_cnt[4], __popc(hm));
                }
            }
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            sl = __shfl_sync(0xffffffffu, sl, 0);
            sh = __shfl_sync(0xffffffffu, sh, 0);
            if (want && s0 + __popc(am) <= cap) {
                const unsigned lt = (1u << lane) - 1u;
                slot = hot ? cap - 1 - (sh + __popc(hm & lt)) : sl + __popc(lm & lt);
            }
        }
        const bool spill = valid && (bad || (f != 0 && slot < 0));
        const unsigned spm = __ballot_sync(0xffffffffu, spill);
        if (spm) {
            int s0 = 0;
            if (lane == 0) s0 = atomicAdd(&s_cnt[1], __popc(spm));
            s0 = __shfl_sync(0xffffffffu, s0, 0);
            if (spill) spill_list[s0 + __popc(spm & ((1u << lane) - 1u))] = (unsigned)l;
        }
        if (valid) {
            if (!spill) {
                if (cutoff) sq.lp[p] = lp;   // no event ever touches the low-pass state
                if (!f) {
                    sq.base[p] = base;
                    sq.tmem[p] = tm;
                }
            }
            for (int b = 0; b < nb; ++b) fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, 0.f);
            if (slot >= 0) {
                atomicAdd(&s_hist[(hot ? 32 : 0) + f], 1);   // for the ordering by first frame
                unsigned* e = s_ent + slot;
                e[FX_META * cap] = (unsigned)f | (tm == 0.f ? 32u : 0u) | (cm << 6);
                e[FX_IDX * cap] = (unsigned)l;
                e[FX_BASE * cap] = __float_as_uint(sb);
                e[FX_POS * cap] = __float_as_uint(pos);
                e[FX_NEG * cap] = __float_as_uint(neg);
#pragma unroll
                for (int k = 0; k < kFxChunk; ++k) {
                    if (k >= T - 1) break;
                    e[(nw_off + k) * cap] = __float_as_uint(nwv[k]);
                    if (leak) e[(dl_off + k) * cap] = __float_as_uint(dlv[k]);
                }
            }
        }
    }
    __syncthreads();
    const int n_cold = s_cnt[3], n_hot = s_cnt[4];
    const int n_spill = s_cnt[1];
    // ---- order the entries by first active frame (counting sort; order within a frame is free) ---
    if (tid == 0) {
        int o = 0;
        for (int i = 0; i < 64; ++i) {   // cold frames 0..31, then hot frames 0..31
            const int c = s_hist[i];
            s_hist[64 + i] = o;
            o += c;
        }
    }
    __syncthreads();
    for (int i = tid; i < n_cold + n_hot; i += kFxThreads) {
        const bool h = i >= n_cold;
        const int e = h ? cap - 1 - (i - n_cold) : i;
        const int f = (int)(s_ent[FX_META * cap + e] & 31u);
        s_perm[atomicAdd(&s_hist[64 + (h ? 32 : 0) + f], 1)] = (unsigned short)e;
    }
    __syncthreads();
    FX_STAMP(2);
    if (tl && tid == 0) { tl[28] = n_cold + n_hot; tl[29] = n_spill; tl[26] = n_hot; }

    // ================================ resolve passes ================================================
    // round 0 is a pre-pass over the hot entries only: under last pack's guess it yields the
    // maxima those entries reach, which are nearly always the image-wide ones - the guess of the
    // first full pass.  Rounds >= 1 resolve every entry and test the guess they ran under.
    int pass = 0;
    bool converged = false;
    for (; pass < T + 3; ++pass) {
        const bool pre = pass == 0;
        sm.tl = (tl && pass < 2) ? tl + 14 + 6 * pass : nullptr;
        if (sm.tl) {
            if (tid < 6) sm.tl[tid] = 0;
            __syncthreads();
        }
        fx_resolve_compact<FDT, R, NOISE, SPEC>(prm, sq, sm, s_ent, pre ? 0 : n_cold, n_hot, n_cold,
                                                &s_cnt[5], tile_px0, leak, shot, philox);
        if (n_spill && !pre)
            fx_resolve_spill<FDT, R, GENERIC, SPEC>(prm, sq, sm, spill_list, n_spill, tile_px0, spill_vox,
                                              spill_state, leak, shot, philox, cutoff);
        __syncthreads();
        if (pass < 4) FX_STAMP(3 + 3 * pass);
        // ---- grid barrier of this sequence: accumulate (maxima, partials), arrive, wait, read ------
        // Everything a round publishes is an order-free integer atomic on the round's own words
        // (FxWork); thread 0 arrives with a release and polls the arrival count with acquires.
        {
            unsigned* gm = fx.gmax[pass];
            unsigned long long* gs = fx.stat[pass];
            if (tid >= 1 && tid < T) {
                if (s_max[tid]) atomicMax(gm + tid, s_max[tid]);
            } else if (tid >= 32 && tid < 38) {
                // nnz, sum, the 128-bit sum of squares as three 40-bit limbs, events
                const unsigned long long lo = s_acc[2], hi = s_acc[3];
                const int k = tid - 32;
                const unsigned long long v =
                    k == 0 ? s_acc[0] : k == 1 ? s_acc[1] : k == 2 ? (lo & 0xffffffffffull)
                    : k == 3 ? ((lo >> 40) | ((hi & 0xffffull) << 24)) : k == 4 ? (hi >> 16) : s_acc[4];
                if (v) atomicAdd(gs + k, v);
            }
            __syncthreads();
            if (tid == 0) {
                fx_red_release_add(&fx.arrive, 1u);
                if (tl && pass < 3) tl[5 + 3 * pass] = fx_now();
                const unsigned target = (unsigned)prm.tiles_per_seq * ((unsigned)pass + 1u);
                int spins = 0;
                while (fx_ld_acquire(&fx.arrive) < target)
                    if (++spins > kFxSpin) { s_cnt[2] = 1; break; }
            }
            __syncthreads();
            if (tid >= 1 && tid < T) {
                s_gmax[tid] = fx_ld_relaxed(gm + tid);
            } else if (tid == 32) {
                unsigned long long q[6];
#pragma unroll
                for (int k = 0; k < 6; ++k) q[k] = fx_ld_relaxed64(gs + k);
                const unsigned long long lo = q[2] + (q[3] << 40);
                s_tot[0] = q[0];
                s_tot[1] = q[1];
                s_tot[2] = lo;
                s_tot[3] = (q[3] >> 24) + (q[4] << 16) + (lo < q[2] ? 1ull : 0ull);
                s_tot[4] = q[5];
            }
        }
        __syncthreads();
        if (pass < 4) FX_STAMP(4 + 3 * pass);
        // ---- m == g ? ------------------------------------------------------------------------------
        const int changed = __syncthreads_or(tid >= 1 && tid < T && (int)s_gmax[tid] != s_g[tid].mx);
        if (!pre && (!changed || s_cnt[2])) { converged = !changed; break; }
        if (tid < V2E_MAX_FRAMES) {
            // after the pre-pass a frame no hot entry fired in keeps last pack's guess
            const int mx = (pre && s_gmax[tid] == 0) ? s_g[tid].mx : (int)s_gmax[tid];
            s_g[tid] = frame_globals(mx, sq.duration, sq.Tr);
            s_max[tid] = 0;
            s_gmax[tid] = 0;
        }
        if (tid < 5) { s_acc[tid] = 0ull; s_tot[tid] = 0ull; }
        if (tid == 0) s_cnt[5] = 0;
        __syncthreads();
    }

    FX_STAMP(27);
    if (tl && tid == 0) tl[30] = pass + 1;   // rounds run, pre-pass included
    // ================================ commit =========================================================
    if (tid == 0) {
        float mean = 0.f, den = 1.f, on = 0.f;
        const double c = (double)s_tot[0];
        if (c > 0. && normalise) {
            // utils/event_process.py:166-175
            const double a = (double)(long long)s_tot[1] * (1.0 / 2147483648.0);
            const double b = ((double)s_tot[3] * 18446744073709551616.0 + (double)s_tot[2]) *
                             (1.0 / 2147483648.0) * (1.0 / 2147483648.0);
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
        for (int i = tid; i < n_cold + n_hot; i += kFxThreads) {
            const unsigned* ep = s_ent + (i < n_cold ? i : cap - 1 - (i - n_cold));
            const int p = tile_px0 + (int)ep[FX_IDX * cap];
            sq.base[p] = __uint_as_float(ep[FX_OB * cap]);
            sq.tmem[p] = __uint_as_float(ep[FX_OT * cap]);
            for (int b = 0; b < nb; ++b) {
                float v = __uint_as_float(ep[(FX_V0 + b) * cap]);
                if (v != 0.f || !norm) {
                    if (norm) v = fdiv(fsub(v, mean), den);   // nz * (v - mean) / (std + 1e-8)
                    fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, v);
                }
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
        const int used = min(pass + 1, kFxMaxRounds);
        for (int i = tid; i < used * V2E_MAX_FRAMES; i += kFxThreads) (&fx.gmax[0][0])[i] = 0u;
        for (int i = tid; i < used * 8; i += kFxThreads) (&fx.stat[0][0])[i] = 0ull;
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            fx.exit_count = 0;
            fx.arrive = 0;
            if (prm.timeline && seq == 0) {
                prm.timeline[160 * 32 + 2 * (fx.pad[0] & 127u) + 1] = fx_now();
                fx.pad[0] += 1;
            }
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct FxPlan {
    int ok, spec;
    int tiles_per_seq, tile_px, cap, nw_off, dl_off;
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

static FxPlan fx_plan(const CfgDev& cfg, int n_seq, int T) {
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
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const int nfr = T - 1 < kFxChunk ? T - 1 : kFxChunk;
    pl.nw_off = FX_V0 + cfg.nb;
    pl.dl_off = pl.nw_off + nfr;
    const int ewords = pl.dl_off + (leak ? nfr : 0);
    int cap = (int)(((size_t)g_fx_optin - kFxHeader) / (4u * ewords)) & ~31;
    if (cap > tile_px) cap = tile_px;
    if (cap > kFxPermMax) cap = kFxPermMax;
    if (cap < 32) return pl;
    pl.ok = 1;
    // the specialised kernels assume the reference pack shape and the full B200 shared memory
    // (their table may be larger than the tile: unused capacity)
    // (and pixel pairs that never straddle a row: even width)
    pl.spec = T == kFxSpecT && cfg.nb == kFxSpecNb && g_fx_optin >= kFxSpecSmem && cfg.W % 2 == 0;
    if (pl.spec) cap = fx_spec_cap(leak);
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
    return fx_plan(cfg, n_seq, 10).ok;
}

template <int FDT, int R, int NOISE, bool CUTOFF, bool SPEC>
static int launch_fixpoint_v(const FxParams& prm, const FxPlan& pl, cudaStream_t st) {
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_fixpoint<FDT, R, NOISE, CUTOFF, SPEC>;
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

template <int FDT, int R, int NOISE, bool CUTOFF>
static int launch_fixpoint(const FxParams& prm, const FxPlan& pl, cudaStream_t st) {
    // the reference's pack shape without debug maps: the specialised kernel (v2e_fixpoint_spec.cu)
    if (NOISE != 2 && pl.spec)
        return launch_fixpoint_spec(prm, FDT, NOISE, CUTOFF, R, pl.tiles_per_seq * prm.n_seq, pl.smem,
                                    g_fx_optin, st);
    return launch_fixpoint_v<FDT, R, NOISE, CUTOFF, false>(prm, pl, st);
}

// n_chain > 1: packs[i * n_seq + s], i < n_chain, are consecutive packs of sequence s, all served by
// one launch of the specialised kernel; V2E_ERR_UNSUPPORTED (before anything is enqueued) when the
// request does not qualify - the caller then launches step by step.
template <int FDT>
static int run_fixpoint_t(const CfgDev& cfg, int rounds, const v2e_pack* packs, int n_seq,
                          void* workspace, cudaStream_t st, int n_chain, void* ev_launch) {
    const FxPlan pl = fx_plan(cfg, n_seq, packs[0].sc.num_frames);
    if (!pl.ok) return V2E_ERR_UNSUPPORTED;
    const int n_blk = (n_chain > 1 ? n_chain : 1) * n_seq;
    if (n_chain > 1) {
        if (n_chain > V2E_MAX_SEQ_PER_LAUNCH || n_blk > kFxBlkSlots || !pl.spec) return V2E_ERR_UNSUPPORTED;
        for (int b = 0; b < n_blk; ++b) {
            const v2e_pack& p = packs[b];
            const v2e_pack& q = packs[b % n_seq];     // the sequence's first pack of the chain
            if (p.sc.num_frames != packs[0].sc.num_frames || p.base_log_frame != q.base_log_frame ||
                p.lp_log_frame0 != q.lp_log_frame0 || p.timestamp_mem != q.timestamp_mem || p.pos_thres != q.pos_thres ||
                p.neg_thres != q.neg_thres || p.noise_rate != q.noise_rate ||
                p.dbg_pol || p.dbg_cnt || p.dbg_final || p.dbg_leak_randn || p.dbg_shot)
                return V2E_ERR_UNSUPPORTED;
        }
    }
    static thread_local FxParams prm;
    const StepLayout lay = step_layout(cfg, n_seq);
    prm.cfg = cfg;
    prm.n_seq = n_seq;
    prm.n_chain = n_chain > 1 ? n_chain : 1;
    prm.tiles_per_seq = pl.tiles_per_seq;
    prm.tile_px = pl.tile_px;
    prm.cap = pl.cap;
    prm.nw_off = pl.nw_off;
    prm.dl_off = pl.dl_off;
    prm.fx = (FxWork*)((char*)workspace + lay.fx_off);
    // the stepwise engine's accumulator areas are unused here: borrow them for the spilled pixels
    prm.spill_vox = (float*)((char*)workspace + lay.vox_off);
    prm.spill_list = (unsigned*)((char*)workspace + lay.cntpol_off);
    prm.spill_state = (float*)((char*)workspace + lay.fxstate_off);
    static const bool want_timeline = getenv("V2E_TIMELINE") != nullptr;
    prm.timeline = want_timeline ? (long long*)((char*)workspace + lay.tl_off) : nullptr;
    bool dbg = false;
    static thread_local std::vector<FxSeqBlock> stage_v;
    static thread_local unsigned ring_pos = 0;
    if ((int)stage_v.size() < n_blk) stage_v.resize(n_blk);
    FxSeqBlock* const stage = stage_v.data();
    for (int s = 0; s < n_blk; ++s) {
        FxSeqBlock& b = n_blk == 1 ? prm.one : stage[s];
        fill_seq(b.sq, packs[s]);
        for (int n = 1; n < packs[s].sc.num_frames; ++n)
            b.fr[n] = make_frame_sc(cfg, packs[s].sc.fr[n]);
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
    if (n_chain > 1 && noise == 2) return V2E_ERR_UNSUPPORTED;   // (nothing has been enqueued yet)
    prm.blocks = nullptr;
    prm.fr_in_params = n_blk == 1;
    if (n_chain > 1 && n_seq == 1) {
        // a chain whose packs share the stream pass's frame scalars bit for bit (evenly spaced
        // stamps: the usual case) keeps them in the kernel parameters, like a single pack
        bool same = true;
        for (int s = 1; same && s < n_chain; ++s)
            for (int n = 1; same && n < packs[0].sc.num_frames; ++n)
                same = memcmp(&stage[s].fr[n].s_large, &stage[0].fr[n].s_large, sizeof(float)) == 0 &&
                       memcmp(&stage[s].fr[n].s_small, &stage[0].fr[n].s_small, sizeof(float)) == 0 &&
                       memcmp(&stage[s].fr[n].sdt, &stage[0].fr[n].sdt, sizeof(float)) == 0;
        if (same) {
            for (int n = 1; n < packs[0].sc.num_frames; ++n) prm.one.fr[n] = stage[0].fr[n];
            prm.one.sq = stage[0].sq;   // (the planes of the sequence: the same for every pack)
            prm.fr_in_params = 1;
        }
    }
    if (n_blk > 1) {
        // the records of a multi-sequence launch or a chain go through the workspace, copied in stream order
        // (a ring of slots: the copy for launch i+1 never touches what launch i is reading)
        if (ring_pos + (unsigned)n_blk > (unsigned)kFxBlkSlots) ring_pos = 0;
        FxSeqBlock* slot = (FxSeqBlock*)((char*)workspace + lay.fxblk_off) + ring_pos;
        ring_pos += (unsigned)n_blk;
        // (pieces of at most 32 KB: small copies from pageable memory are staged by the driver at
        // once and never wait for the stream)
        const int per = (int)(32768 / sizeof(FxSeqBlock));
        for (int b0 = 0; b0 < n_blk; b0 += per) {
            const int nb_ = n_blk - b0 < per ? n_blk - b0 : per;
            const int rc = cuda_ok(cudaMemcpyAsync(slot + b0, stage + b0, sizeof(FxSeqBlock) * nb_,
                                                   cudaMemcpyHostToDevice, st));
            if (rc != V2E_OK) return rc;
        }
        prm.blocks = slot;
    }
    if (ev_launch) {   // (timing aid: after the record copy, right before the kernel)
        const int rc = cuda_ok(cudaEventRecord((cudaEvent_t)ev_launch, st));
        if (rc != V2E_OK) return rc;
    }
#ifdef V2E_FX_DEV   // development builds: one Philox depth only (fewer instantiations to compile)
#define FX_BY_ROUNDS(N, C) launch_fixpoint<FDT, 10, N, C>(prm, pl, st)
#else
#define FX_BY_ROUNDS(N, C) \
    (rounds == 7 ? launch_fixpoint<FDT, 7, N, C>(prm, pl, st) : launch_fixpoint<FDT, 10, N, C>(prm, pl, st))
#endif
    (void)rounds;
    if (noise == 0)
        return cfg.cutoff_on ? launch_fixpoint<FDT, 10, 0, true>(prm, pl, st)
                             : launch_fixpoint<FDT, 10, 0, false>(prm, pl, st);
    if (noise == 1) return cfg.cutoff_on ? FX_BY_ROUNDS(1, true) : FX_BY_ROUNDS(1, false);
    return FX_BY_ROUNDS(2, false);
#undef FX_BY_ROUNDS
}

int run_fixpoint(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st, int n_chain, void* ev_launch) {
    switch (frame_dtype) {
#ifndef V2E_FX_DEV
        case V2E_DT_U8: return run_fixpoint_t<V2E_DT_U8>(cfg, rounds, packs, n_seq, workspace, st, n_chain, ev_launch);
        case V2E_DT_BF16: return run_fixpoint_t<V2E_DT_BF16>(cfg, rounds, packs, n_seq, workspace, st, n_chain, ev_launch);
#else
        case V2E_DT_U8: case V2E_DT_BF16: return V2E_ERR_UNSUPPORTED;
#endif
        default: return run_fixpoint_t<V2E_DT_F32>(cfg, rounds, packs, n_seq, workspace, st, n_chain, ev_launch);
    }
}

}  // namespace v2e
