// v2e_fixpoint_spec.cu — the fixed-point engine specialised for the reference's pack shape
// (num_pack_frames = 10, num_bins = 5: utils/configs.py:6,18) with the two noise wirings that are
// served without debug maps: none, or PHILOX leak + shot (what V2E2VNet hard-wires,
// model_v2e2v.py:56-57).  Same algorithm, entry table, noise streams and results as the generic
// kernel in v2e_fixpoint.cu (speculate on the per-frame maxima, verify, repeat); what differs is
// how the work is laid out on the SM:
//
//   pass 1   two pixels per lane, 64-bit loads and stores, every frame of the pack in registers;
//            the frame loop is straight-line code that only decides "may anything happen at this
//            pixel" (|diff| against the smaller threshold, a zero test on the packed shot fields).
//            Quiet pixels (~85 %) are finished here.  The others get a table entry holding the
//            inputs no event can change: per frame the low-passed log intensity and the leak
//            decrement.
//   post     one entry per lane over the compact table: first active frame, brightness before
//            it, "hot" flag; counting sort by first active frame.
//   resolve  one entry per lane, voxel accumulators in registers, table addresses bumped per
//            frame; a pixel's usual event frame (one event at the frame's first time stamp,
//            whose bin and weights are per-frame constants) is a dozen instructions; per-frame
//            maxima via one warp reduction; anything unusual (shot candidates, gate off) goes
//            through the generic out-of-line code.
//   barrier  order-free integer atomics on the round's own words + one arrival counter.
#include "v2e_fixpoint.cuh"

namespace v2e {

constexpr int kSpNF = kFxSpecT - 1;   // frames 1..9 of the pack carry events
// Threads per CTA.  A 1280x720 tile (6 227 pixels) is 97.3 warp chunks of 64 pixels: 20 warps take
// them in 4.87 rounds, 24 warps in 4.05 - a fifth round for one chunk; 20 warps also leave each
// thread 96 registers (no spills in the stream loop).
#ifndef V2E_SP_THREADS
#define V2E_SP_THREADS 640
#endif
constexpr int kSpThreads = V2E_SP_THREADS;
#ifndef V2E_SP_HOT_LANES
#define V2E_SP_HOT_LANES 32
#endif
constexpr int kSpHotLanes = V2E_SP_HOT_LANES;   // entries per task of the hot list
#ifndef V2E_SP_COLD_LANES
#define V2E_SP_COLD_LANES 32
#endif
constexpr int kSpColdLanes = V2E_SP_COLD_LANES;   // entries per task of the cold list
#ifndef V2E_SP_PRE
#define V2E_SP_PRE 2
#endif
constexpr unsigned kSpStatic = 0x40000u, kSpStaticUp = 0x80000u;   // entry META flags (post pass)
#ifndef V2E_SP_BUSY
#define V2E_SP_BUSY 2
#endif
constexpr int kSpBusy = V2E_SP_BUSY;   // crossing frames after the first that make an entry "busy"
constexpr unsigned kSpShotData = 0x100000u;   // FX_SI / FX_SU hold the entry's one marked frame
constexpr int kSpPre = V2E_SP_PRE;   // pre-pass rounds (hot entries only) before the first full round
static_assert(kSpPre >= 1 && kSpPre <= 2, "FxWork::hint holds two rounds");

// Per-frame constants of the resolve loop, all derived from the frame's guessed maximum
// (v2e_model.py:405-422,469): the time axis of the frame's iterations and what an event at the
// first time stamp needs (:478-502) - its stamp, bin and bilinear weights are the same for every
// pixel of the image.  Counts are kept as floats (small integers, exact): no conversions on the
// per-pixel dependency chain.
struct SpFrameEv {
    float ts1, wl, wr, bbf;     // iteration 0: stamp, weights of bins bb and bb+1, bb as a float
    float t0, step, rstep, mxf; // time_frames[n-1], ts_step, ~1/ts_step (estimates only), max count
};

template <int FDT>
__device__ __forceinline__ float2 sp_load_frame2(const void* __restrict__ f, long long idx) {
    // idx is even and the plane is 16 B aligned
    if (FDT == V2E_DT_U8) {
        const uchar2 v = __ldg((const uchar2*)((const unsigned char*)f + idx));
        return make_float2((float)v.x, (float)v.y);
    }
    if (FDT == V2E_DT_BF16) {
        const unsigned v = __ldg((const unsigned*)((const unsigned short*)f + idx));
        return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    }
    return __ldg((const float2*)((const float*)f + idx));
}

// frame value -> table index.  Integer-valued 0..255 maps to itself; anything else (fraction,
// out of range, NaN) is flagged and the pixel takes the fp64 path of the generic code.
template <int FDT>
__device__ __forceinline__ unsigned sp_index(float x, bool& bad) {
    unsigned v = min(__float2uint_rn(x), 255u);
    if (FDT != V2E_DT_U8) bad = bad || ((float)v != x);
    return v;
}

template <int NOISE>
struct SpPx {   // one pixel after the frame loop
    float nw[kSpNF], dl[NOISE == 1 ? kSpNF : 1];
    float lp, base, base0, tm;
    unsigned w0, w1, w2;   // the packed shot fields (call B of the pixel's noise block)
    bool act, cand, bad, allcand;
};

// The frame loop of one pixel (v2e_model.py:321-322,332,350-404 up to the threshold test):
// everything the frames present to the pixel whatever it does.
// ONE: the stream pass's per-frame scalars sit in the kernel parameters (a single pack, or a chain
// whose packs all have the same ones - evenly spaced stamps): they become
// constant-bank operands of the arithmetic instead of a shared-memory load each.
// SMALL: the pixel sits on the sensing-diversity grid (even row and column), a compile-time
// choice so that the scalar is an operand and not a select.
template <int FDT, int R, int NOISE, bool CUTOFF, bool ONE, bool SMALL>
__device__ __forceinline__ void sp_stream_px(SpPx<NOISE>& o, const float (&x)[kSpNF], float x0,
                                             float pos, float neg, float nra, float base_in,
                                             float lp_in, float tm_in, unsigned p,
                                             const FxParams& prm, const SeqDev& sq,
                                             const float2* s_lut, const FrameSc* s_fr,
                                             float kmax4) {
    const CfgDev& cfg = prm.cfg;
    bool bad = false;
    float base, lp, tm;
    if (sq.first_pack) {
        // _init (v2e_model.py:156-174,239-240)
        base = s_lut[sp_index<FDT>(x0, bad)].x;
        lp = base;
        tm = fadd(0.f, -sq.Tr);
    } else {
        base = base_in;
        lp = lp_in;
        tm = tm_in;
        // pack rollover (v2e_model.py:329-330)
        if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
        if (tm < 0.f) tm = -sq.Tr;
    }
    o.base0 = base;
    // crossing a threshold needs |diff| >= C + 1e-9 for the polarity's C (v2e_model.py:403-404):
    // below the smaller of the two nothing can happen
    const float dmin = fminf(fadd(pos, 1e-9f), fadd(neg, 1e-9f));
    uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
    float nlk = 0.f;
    if (NOISE == 1) {
        wa = noise_px<R>(sq.philox_ctr, p, 0u, 0);
        wb = noise_px<R>(sq.philox_ctr, p, 0u, 1);
        nlk = fmul(cfg.leak_rate, nra);
    }
    const float jit = cfg.leak_jitter;
    bool act = false;
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int k = 0; k < kSpNF; ++k) {
        const FrameSc& fs = ONE ? prm.one.fr[k + 1] : s_fr[k + 1];
        const float2 li = s_lut[sp_index<FDT>(x[k], bad)];
        float nwk = li.x;
        if (CUTOFF) {
            lp = lowpass(lp, li.x, li.y, SMALL ? fs.s_small : fs.s_large);
            nwk = lp;
        }
        o.nw[k] = nwk;
        float b = base;
        if (NOISE == 1) {
            if ((k & 1) == 0) box_muller16(k < 8 ? (&wa.x)[k >> 1] : wb.w, z0, z1);
            const float dlk = leak_dec(nlk, jit, (k & 1) ? z1 : z0, fs.sdt, pos);
            o.dl[k] = dlk;
            b = fsub(base, dlk);
        }
        act = act || fabsf(fsub(nwk, b)) >= dmin;
        base = b;
    }
    o.lp = lp;
    o.base = base;
    o.tm = tm;
    o.act = act;
    o.bad = bad;
    o.cand = false;
    o.allcand = false;
    o.w0 = wb.x & 0x3fffffffu;
    o.w1 = wb.y & 0x3fffffffu;
    o.w2 = wb.z & 0x3fffffffu;
    if (NOISE == 1) {
        // A frame can hold a shot event only if its field is below 1024*bound (fx_shot_thr).
        // While that is <= 1 for this pixel (any practical rate) the nine tests collapse into
        // "does any 10-bit field equal zero", three instructions per word.
        o.allcand = !(kmax4 * fmaxf(pos, neg) <= 1.0f);
        const unsigned L = 0x00100401u, H = 0x20080200u;
        const unsigned zf = (((o.w0 - L) & ~o.w0) | ((o.w1 - L) & ~o.w1) | ((o.w2 - L) & ~o.w2)) & H;
        o.cand = zf != 0u || o.allcand;
    }
}

// What the iteration loop (v2e_model.py:449-475) does to one pixel-frame, for the usual outcomes -
// no event or exactly one - without walking it.  Iteration j of [0,c) carries an event; with the
// gate on (:469) it is accepted iff ts_j - timestamp_mem > Tr (:419,:470), monotone in j.
// Straight-line code:
//   one  - exactly one event, at ts (iteration 0 when `first`)
//   slow - anything else that is not "no event"
struct SpEv1 {
    bool one, slow, first;
    float ts;
};
__device__ __forceinline__ SpEv1 sp_classify(const SpFrameEv& ev, bool gate, float cntf, bool bit,
                                             float tm, float Tr) {
    const float cf = fminf(cntf, ev.mxf);
    const float tc = fadd(ev.t0, fmul(ev.step, cf));        // stamp of iteration c-1
    const bool first = !gate || fsub(ev.ts1, tm) > Tr;      // iteration 0 passes
    // else the first iteration that can pass, from an estimate: iteration j passes iff
    // step * (j + 1) > Tr + timestamp_mem - t0 =: X, so the first one is floor(X / step).
    // X / step sits next to an integer whenever the maxima of consecutive frames are commensurable
    // (the usual case), so two neighbours are tried, each with the exact predicate.
    float jl = floorf((Tr + tm - ev.t0) * ev.rstep - 0.5f);
    jl = fmaxf(fminf(jl, cf - 1.f), 1.f);
    const float tsa = fadd(ev.t0, fmul(ev.step, jl));         // iteration jl-1
    const float tsb = fadd(ev.t0, fmul(ev.step, jl + 1.f));   // iteration jl
    const float tsc = fadd(ev.t0, fmul(ev.step, jl + 2.f));   // iteration jl+1
    const bool pb = fsub(tsb, tm) > Tr;
    const bool any = cf >= 1.f && (!gate || fsub(tc, tm) > Tr);
    const bool est = cf >= 2.f && !(fsub(tsa, tm) > Tr) &&
                     (pb || (jl + 2.f <= cf && fsub(tsc, tm) > Tr));
    SpEv1 r;
    r.first = first;
    r.ts = first ? ev.ts1 : (pb ? tsb : tsc);
    const float nxt = first ? 1.f : (pb ? jl + 1.f : jl + 2.f);   // iterations consumed
    // one more after it?  (gate off: every iteration is an event)
    const bool more = nxt < cf && (!gate || fsub(tc, r.ts) > Tr);
    r.slow = bit || (any && ((!first && !est) || more));
    r.one = any && !r.slow;
    if (r.slow && !bit && gate && any && !first && !est) {
        // the estimate missed (rare): walk to the first iteration that passes, with the exact
        // predicate - it exists (`any`: the last one passes) and the predicate is monotone
        float j = 1.f;   // iteration index; iteration 0 was refused (`first`)
        float ts = fadd(ev.t0, fmul(ev.step, 2.f));
        while (j + 1.f < cf && !(fsub(ts, tm) > Tr)) {
            j += 1.f;
            ts = fadd(ev.t0, fmul(ev.step, j + 1.f));
        }
        // one more after it?  (the last iteration decides)
        if (!(j + 1.f < cf && fsub(tc, ts) > Tr)) {
            r.slow = false;
            r.one = true;
            r.ts = ts;
        }
    }
    return r;
}

// What the exact shot test of a marked frame needs when the entry does not carry it (several marks
// survived the post pass, or not the first one): the frame's inten01 and uniform, fetched on the
// spot.  A marked frame has field 0 (entries whose every frame is a candidate are not served
// here), so the uniform is fx_shot_uniform(.., field 0).  False: not an integer-valued frame.
template <int FDT, int R>
__device__ __forceinline__ bool sp_mark_data(const SeqDev& sq, const float2* s_lut, int P, int n, int p,
                                             float& I, float& u) {
    const float x = load_frame1<FDT>(sq.frames, (long long)n * P + p);
    const int v = __float2int_rn(x);
    if (!((float)v == x && (unsigned)v < 256u)) return false;
    I = s_lut[v].y;
    const uint4 w = noise_words<R>(sq.philox_ctr, (unsigned)n, (unsigned)p, 1);
    u = (0.f + (float)w.x * (1.0f / 4294967296.0f)) * (1.0f / 1024.0f);
    return true;
}

// A frame whose shot candidate is real (the exact first-stage test passed), without the generic
// iteration loop (iterate_pixel, shot_mode 2) in the usual cases:
//   kind 0 - the frame is as if unmarked: the first firing falls beyond the frame's iterations, or
//            on an iteration that carries an event of the pixel's own anyway;
//   kind 1 - after the pixel's own events (iterations [0, cnt), possibly none) the shot fires once,
//            at stamp ts of a later iteration (the refractory gate is the caller's);
//   kind 2 - anything else (a second firing within the frame, arguments outside the small-argument
//            branch of geometric_skip): the generic code.  r.ts then says why (development builds).
// sa: this frame's (rate/2)*dt / num_iters (emulator_utils.py:174), in the precision the reference
// uses for it (trap T5).  Same arithmetic and draws as iterate_pixel, hence the same results.
struct SpShot {
    int kind;
    float ts;
    float g, p_it;   // iteration of the first firing and the per-iteration probability (kinds 1, 2 with ts >= 1)
};

// The iteration loop of a frame with shot firings, for the cases sp_shot_lean does not settle in
// closed form (a second firing within the frame, events of the pixel's own that the refractory
// gate does not reduce to one): iterate_pixel's shot_mode-2 walk over the iterations that can
// carry an event - [0, cnt) and the firings - under the same preconditions as sp_shot_lean
// (p_it < 1e-3, at most 24 iterations: a follow-up uniform >= 0.03 puts the next firing beyond the
// frame without forming the logarithms).  Small and out of line: it runs a few times per pack, and
// the generic fx_event_frame / fx_events pair costs microseconds of instruction fetch each time.
// Returns (final << 32) | bits(timestamp_mem); voxel accumulators at vp[b * vstride].
template <int R>
static __device__ __noinline__ unsigned long long sp_shot_loop(unsigned long long ctr, int n, int p,
                                                               float t0, float step, float Tr, int gate,
                                                               int mx, int cnt, int first_shot, float p_it,
                                                               float pol, float tm, float* vp, int vstride) {
    int next = first_shot, fin = 0;
    unsigned follow = 2;
    for (int i = cnt == 0 ? next : 0; i < mx;) {
        const int cur = i;
        bool m = cnt >= i + 1;
        if (i == next) {
            m = true;
            const uint4 w = noise_words<R>(ctr, (unsigned)n, (unsigned)p, follow);
            follow = follow < 0x7fu ? follow + 1 : follow;
            const float u2 = (float)(w.x >> 8) * (1.0f / 16777216.0f);
            next = 0x3fffffff;
            if (u2 < 0.03f) next = i + 1 + (int)fminf(geometric_skip_small(u2, p_it), 1.0e9f);
        }
        i = cur + 1;
        if (i >= cnt && i < next) i = next;
        const float ts = fadd(t0, fmul(step, (float)(cur + 1)));
        if (gate) {   // v2e_model.py:469-472
            m = fsub(m ? ts : 0.f, tm) > Tr;
            if (m) tm = ts;
        }
        if (!m) continue;
        fin += 1;
        const float ti = floorf(ts);   // v2e_model.py:478-502
        const float d = fsub(ts, ti);
        const int b = (int)ti;
        if (b >= 0 && b < kFxSpecNb) vp[(size_t)b * vstride] = fadd(vp[(size_t)b * vstride], fmul(pol, fsub(1.0f, d)));
        if (b >= 0 && b + 1 < kFxSpecNb) vp[(size_t)(b + 1) * vstride] = fadd(vp[(size_t)(b + 1) * vstride], fmul(pol, d));
    }
    return ((unsigned long long)(unsigned)fin << 32) | (unsigned long long)__float_as_uint(tm);
}
template <int R>
__device__ __forceinline__ SpShot sp_shot_lean(const CfgDev& cfg, const SeqDev& sq, double sa, bool f64,
                                               const SpFrameEv& ev, int n, int p, bool up, float C,
                                               float cntf, float I, float u) {
    SpShot r;
    r.kind = 2;
    r.ts = 0.f;
    r.g = 0.f;
    r.p_it = 0.f;
    if (!(ev.mxf >= 1.f) || !(ev.mxf <= 24.f)) return r;
    const float fac = fadd(fmul(cfg.shot_inten_m1, I), 1.0f);
    const float pp = fmul(up ? cfg.rn_pos : cfg.rn_neg, C);
    float p_it;
    if (f64) p_it = (float)__dmul_rn(__dmul_rn(sa, (double)fac), (double)pp);
    else p_it = fmul(fmul((float)sa, fac), pp);
    if (!(p_it > 0.f && p_it < 1e-3f && u < 0.03f)) return r;
    const float g = geometric_skip_small(u, p_it);   // iteration of the first firing
    if (!(g < ev.mxf)) {
        r.kind = 0;
        return r;
    }
    r.g = g;
    r.p_it = p_it;
    r.ts = 1.f;
    // the next firing, from a fresh uniform (drawn when the loop reaches the first one).  It lies
    // beyond the frame whenever the uniform is not tiny: log1p(-0.03) / log1p(-p) > 30 iterations
    // for p < 1e-3 and a frame has at most 24 here, so those logarithms are never formed.
    const uint4 w = noise_words<R>(sq.philox_ctr, (unsigned)n, (unsigned)p, 2u);
    const float u2 = (float)(w.x >> 8) * (1.0f / 16777216.0f);
    if (!(u2 >= 0.03f)) {
        r.ts = 2.f;
        if (!(g + 1.f + geometric_skip_small(u2, p_it) >= ev.mxf)) return r;
    }
    if (g < cntf) {
        r.kind = 0;
        return r;
    }
    r.kind = 1;
    r.ts = fadd(ev.t0, fmul(ev.step, g + 1.f));   // stamp of iteration g (:419)
    return r;
}

template <int FDT, int R, int NOISE, bool CUTOFF, bool ONE>
__global__ void __launch_bounds__(kSpThreads, 1) k_fixpoint_spec(const __grid_constant__ FxParams prm) {
    constexpr bool LEAK = NOISE == 1, SHOT = NOISE == 1;
    constexpr int T = kFxSpecT, nb = kFxSpecNb;
    constexpr int cap = fx_spec_cap(LEAK);
    constexpr int nw_off = FX_V0 + nb, dl_off = nw_off + kSpNF;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const CfgDev& cfg = prm.cfg;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    // A launch serves one pack of n_seq sequences, or a chain of n_chain consecutive packs of each
    // of them (the caller's pack loop, test.py:115-121, in one launch: every CTA keeps its tile,
    // the state planes carry over in place, and neither the launch gap nor the set-up is paid
    // again).  Consecutive packs of a chain alternate between two sets of barrier words, so that
    // the last CTA to leave a pack can zero its set while the others are already in the next.
    for (int pk = 0; pk < prm.n_chain; ++pk) {
    // this pack's launch record -> shared memory
    const FxSeqBlock* blk = prm.blocks ? prm.blocks + (size_t)pk * prm.n_seq + seq : &prm.one;
    {
        unsigned* d_sq = (unsigned*)(smem_raw + 3840);
        const unsigned* s_sq = (const unsigned*)&blk->sq;
        for (int i = tid; i < (int)(sizeof(SeqDev) / 4); i += kSpThreads) d_sq[i] = s_sq[i];
        unsigned* d_fr = (unsigned*)(smem_raw + 2048);
        const unsigned* s_fr0 = (const unsigned*)&blk->fr[0];
        for (int i = tid; i < (int)(sizeof(FrameSc) * T / 4); i += kSpThreads) d_fr[i] = s_fr0[i];
    }
    __syncthreads();
    const SeqDev& sq = *(const SeqDev*)(smem_raw + 3840);
    // what every pack of this launch shares (the sequence's planes) is read from the kernel
    // parameters when they hold it (ONE): constant-bank operands instead of shared-memory loads
    // that would be repeated after every table store
    const SeqDev& sqc = ONE ? prm.one.sq : sq;
    FxWork& fx = prm.fx[seq + (pk & 1) * prm.n_seq];
    FxWork& fx0 = prm.fx[seq];   // holds the hint for the launch that follows
    const int P = cfg.P;
    const bool normalise = cfg.output_mode == V2E_OUT_VOXEL;

    // ---- shared memory carve-up (header layout of v2e_fixpoint.cu) --------------------------------
    float2* s_lut = (float2*)smem_raw;                                   // 2048
    FrameSc* s_fr = (FrameSc*)(smem_raw + 2048);                         // 10 * 32
    SpFrameEv* s_ev = (SpFrameEv*)(smem_raw + 2368);                     // 10 * 32
    FrameGlobals* s_g = (FrameGlobals*)(smem_raw + 3072);                // 32 * 12 = 384
    unsigned* s_max = (unsigned*)(smem_raw + 3456);                      // 128: this tile, this round
    unsigned* s_gmax = (unsigned*)(smem_raw + 3584);                     // 128: all tiles
    unsigned long long* s_acc = (unsigned long long*)(smem_raw + 3712);  // 5 * 8: tile partials
    unsigned long long* s_tot = (unsigned long long*)(smem_raw + 3752);  // 5 * 8: sequence totals
    int* s_cnt = (int*)(smem_raw + 3792);   // [0] slot counter, [1] spilled, [2] error, [3] entries, [5] task counter
    float* s_stat = (float*)(smem_raw + 3824);                           // mean, den, on, theta
    int* s_sc = (int*)(smem_raw + 2944);    // 2 x 16: single-event entries per frame, ON / OFF
    double* s_sa = (double*)(smem_raw + 3200);   // 10 doubles in the unused tail of s_g (frames >= T)
    unsigned short* s_perm = (unsigned short*)(smem_raw + kFxPermOff);
    int* s_hist = (int*)(smem_raw + kFxPermOff + 2 * kFxPermMax);        // 2 x 64 ints
    unsigned* s_ent = (unsigned*)(smem_raw + kFxHeader);
    FxSmem sm;
    sm.lut = s_lut; sm.fr = s_fr; sm.g = s_g; sm.max = s_max; sm.acc = s_acc; sm.perm = s_perm; sm.tl = nullptr;

    const int tile_px0 = tile * prm.tile_px;
    const int tile_n = max(0, min(prm.tile_px, P - tile_px0));   // even (P and tile_px are)
    unsigned* spill_list = prm.spill_list + (size_t)seq * P + tile_px0;
    float* spill_vox = prm.spill_vox + (size_t)seq * nb * P;
    float* spill_state = prm.spill_state + (size_t)seq * 3 * P;

    // development aid (V2E_TIMELINE=1): stamps are kept in shared memory and written out at the end
    long long* tl_out = prm.timeline ? prm.timeline + (size_t)blockIdx.x * 32 : nullptr;
    long long* tl = tl_out ? (long long*)(smem_raw + 2688) : nullptr;
    if (tl && tid < 32) tl[tid] = 0;
    __syncthreads();
    FX_STAMP(0);
    if (prm.timeline && blockIdx.x == 0 && tid == 0)
        prm.timeline[160 * 32 + 2 * (fx0.pad[0] & 127u)] = fx_now();
    if (pk == 0) load_lut(s_lut, cfg.lut);
    auto set_globals = [&](int n, int mx) {
        const FrameGlobals g = frame_globals(mx, sq.duration, sq.Tr);
        s_g[n] = g;
        SpFrameEv e;
        e.t0 = s_fr[n].t_bin_prev;
        e.step = g.ts_step;
        e.rstep = __fdividef(1.0f, g.ts_step);
        e.mxf = (float)mx;
        e.ts1 = fadd(e.t0, g.ts_step);   // iteration 0: step * 1 is exact (:419)
        const float ti = floorf(e.ts1);
        const float d = fsub(e.ts1, ti);
        e.bbf = ti;
        e.wl = fsub(1.0f, d);
        e.wr = d;
        s_ev[n] = e;
        if (SHOT && mx > 0)   // iterate_pixel's a64 / a32 (emulator_utils.py:174, trap T5)
            s_sa[n] = ((sq.shot_f64_mask >> n) & 1u) ? s_fr[n].shot_c / (double)mx
                                                    : (double)fdiv((float)s_fr[n].shot_c, (float)mx);
    };
    if (tid < V2E_MAX_FRAMES) {
        // the guess: last pack's maxima (a hint kept in the workspace; any value is correct)
        // (later packs of a chain: the maxima every CTA gathered at the end of the last one)
        const unsigned h = pk == 0 ? fx0.guess[tid] : (tid < T ? s_gmax[tid] + 1u : 0u);
        if (tid < T) set_globals(tid, h ? (int)h - 1 : 8);
        s_max[tid] = 0;
        s_gmax[tid] = 0;
    }
    if (tid < 5) { s_acc[tid] = 0ull; s_tot[tid] = 0ull; }
    if (tid < 8) s_cnt[tid] = tid == 3 ? 0x7fffffff : 0;
    if (tid < 128) s_hist[tid] = 0;
    if (tid < 32) s_sc[tid] = 0;
    __syncthreads();
    if (tid == 0) {
        int gmin = 0x7fffffff;
        float kmax = 0.f;
        for (int n = 1; n < T; ++n) {
            gmin = min(gmin, s_g[n].mx);
            kmax = fmaxf(kmax, fmaxf(s_fr[n].kp, s_fr[n].kn));
        }
        s_stat[3] = (float)max(2, gmin / 2);   // theta of the "hot" test
        // 1024 * bound <= kmax4 * C: fx_shot_thr with its intensity factor 1 + m1 * inten01,
        // inten01 in (0, 1], replaced by its largest value
        s_stat[2] = kmax * 4.0f * (1.0f + fmaxf(cfg.shot_inten_m1, 0.f));
    }
    __syncthreads();
    const float theta = s_stat[3];
    const float kmax4 = s_stat[2];

    FX_STAMP(1);
    // ================================ pass 1: stream ===============================================
    // row and column of the warp's chunk, carried along instead of divided out per pixel
    int ch_y, ch_x;
    {
        const int q = tile_px0 + warp * 64;
        ch_y = q / cfg.W;
        ch_x = q - ch_y * cfg.W;
    }
    const int ch_dy = (kSpThreads * 2) / cfg.W, ch_dx = (kSpThreads * 2) - ch_dy * cfg.W;
    for (int c0 = warp * 64; c0 < tile_n; c0 += kSpThreads * 2) {
        const bool valid = c0 + 2 * lane < tile_n;
        const int l = min(c0 + 2 * lane, tile_n - 2);   // lanes past the end shadow the last pair
        const int p = tile_px0 + l;
        const float2 pos2 = __ldg((const float2*)(sqc.pos + p));
        const float2 neg2 = __ldg((const float2*)(sqc.neg + p));
        float2 nra2 = make_float2(0.f, 0.f), base2 = nra2, lp2 = nra2, tm2 = nra2, x02 = nra2;
        if (LEAK) nra2 = __ldg((const float2*)(sqc.nra + p));
        if (sq.first_pack) {
            x02 = sp_load_frame2<FDT>(sq.frames, p);
        } else {
            base2 = *(const float2*)(sqc.base + p);
            if (CUTOFF) lp2 = *(const float2*)(sqc.lp + p);
            tm2 = *(const float2*)(sqc.tmem + p);
        }
        float2 x2[kSpNF];
#pragma unroll
        for (int k = 0; k < kSpNF; ++k) x2[k] = sp_load_frame2<FDT>(sq.frames, (long long)(k + 1) * P + p);
        // sensing-diversity pattern: even row AND even column (v2e_model.py:194,211); W is even,
        // so only the pair's first pixel can sit on an even column
        int yrow = ch_y, xcol = ch_x + (l - c0);
        while (xcol >= cfg.W) { xcol -= cfg.W; ++yrow; }
        const bool small0 = ((yrow | xcol) & 1) == 0;
        ch_y += ch_dy;
        ch_x += ch_dx;
        if (ch_x >= cfg.W) { ch_x -= cfg.W; ++ch_y; }

        float lp_out[2], base_out[2], tm_out[2];
        bool quiet[2], spilled[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float x[kSpNF];
#pragma unroll
            for (int k = 0; k < kSpNF; ++k) x[k] = j ? x2[k].y : x2[k].x;
            SpPx<NOISE> o;
            // (a row holds a whole number of warp chunks at the usual widths, so the branch on
            // the first pixel's grid membership is warp-uniform)
            if (j == 0 && CUTOFF && small0)
                sp_stream_px<FDT, R, NOISE, CUTOFF, ONE, true>(o, x, x02.x, pos2.x, neg2.x, nra2.x, base2.x,
                                                               lp2.x, tm2.x, (unsigned)p, prm, sq, s_lut,
                                                               s_fr, kmax4);
            else
                sp_stream_px<FDT, R, NOISE, CUTOFF, ONE, false>(
                    o, x, j ? x02.y : x02.x, j ? pos2.y : pos2.x, j ? neg2.y : neg2.x, j ? nra2.y : nra2.x,
                    j ? base2.y : base2.x, j ? lp2.y : lp2.x, j ? tm2.y : tm2.x, (unsigned)(p + j), prm, sq,
                    s_lut, s_fr, kmax4);
            // ---- quiet -> finished; maybe active -> table entry (or spill when the table is full)
            const bool want = valid && (o.act || o.cand) && !o.bad;
            const unsigned am = __ballot_sync(0xffffffffu, want);
            int slot = -1;
            if (am) {
                int s0 = 0;
                if (lane == 0) {
                    // allocations only grow: the first one that does not fit ends the table where it
                    // starts, and nothing after it fits either
                    s0 = atomicAdd(&s_cnt[0], __popc(am));
                    if (s0 + __popc(am) > cap) atomicMin(&s_cnt[3], s0);
                }
                s0 = __shfl_sync(0xffffffffu, s0, 0);
                if (want && s0 + __popc(am) <= cap) slot = s0 + __popc(am & ((1u << lane) - 1u));
            }
            const bool spill = valid && (o.bad || (want && slot < 0));
            const unsigned spm = __ballot_sync(0xffffffffu, spill);
            if (spm) {
                int s0 = 0;
                if (lane == 0) s0 = atomicAdd(&s_cnt[1], __popc(spm));
                s0 = __shfl_sync(0xffffffffu, s0, 0);
                if (spill) spill_list[s0 + __popc(spm & ((1u << lane) - 1u))] = (unsigned)(l + j);
            }
            if (slot >= 0) {
                unsigned cm = 0;
                unsigned* e = s_ent + slot;
                if (SHOT && o.cand) {
                    if (o.allcand) {
                        cm = 0x1ffu;
                    } else {
                        float xm = 0.f;   // frame value of the first marked frame
#pragma unroll
                        for (int k = kSpNF - 1; k >= 0; --k) {
                            const unsigned w = k < 3 ? o.w0 : (k < 6 ? o.w1 : o.w2);
                            if (((w >> (10 * (k % 3))) & 1023u) == 0u) { cm |= 1u << k; xm = x[k]; }
                        }
                        // inten01 of that frame, for the post pass (an entry with one mark left)
                        bool bad2 = false;
                        e[FX_SI * cap] = __float_as_uint(s_lut[sp_index<FDT>(xm, bad2)].y);
                        // marked entries are listed (in the space of the entry order, which is built
                        // later): their marks are refined before the post pass, a few lanes per warp
                        s_perm[atomicAdd(&s_cnt[7], 1)] = (unsigned short)slot;
                    }
                }
                // first active frame 0: found by the post pass
                e[FX_META * cap] = (o.tm == 0.f ? 32u : 0u) | (cm << 6) | (o.allcand ? 0x20000u : 0u);
                e[FX_IDX * cap] = (unsigned)(l + j);
                e[FX_BASE * cap] = __float_as_uint(o.base0);
                e[FX_POS * cap] = __float_as_uint(j ? pos2.y : pos2.x);
                e[FX_NEG * cap] = __float_as_uint(j ? neg2.y : neg2.x);
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    e[(nw_off + k) * cap] = __float_as_uint(o.nw[k]);
                    if (LEAK) e[(dl_off + k) * cap] = __float_as_uint(o.dl[k]);
                }
            }
            lp_out[j] = o.lp;
            base_out[j] = o.base;
            tm_out[j] = o.tm;
            spilled[j] = spill;
            quiet[j] = valid && !spill && slot < 0;
        }
        if (valid) {
            // no event ever touches the low-pass state: final for every pixel that is not spilled
            // (the spill path recomputes a pixel from its stored state)
            if (CUTOFF) {
                if (!spilled[0] && !spilled[1]) {
                    *(float2*)(sqc.lp + p) = make_float2(lp_out[0], lp_out[1]);
                } else {
                    if (!spilled[0]) sqc.lp[p] = lp_out[0];
                    if (!spilled[1]) sqc.lp[p + 1] = lp_out[1];
                }
            }
            // (an active pixel's words are written again by the commit; nothing reads them before)
            if (!spilled[0] && !spilled[1]) {
                *(float2*)(sqc.base + p) = make_float2(base_out[0], base_out[1]);
                *(float2*)(sqc.tmem + p) = make_float2(tm_out[0], tm_out[1]);
            } else {
                if (quiet[0]) { sqc.base[p] = base_out[0]; sqc.tmem[p] = tm_out[0]; }
                if (quiet[1]) { sqc.base[p + 1] = base_out[1]; sqc.tmem[p + 1] = tm_out[1]; }
            }
            // zero voxel columns; the commit overwrites those of the active pixels
            // (the pointer is read once: a store through it could alias the record it comes from)
            void* const vox = sq.voxel;
#pragma unroll
            for (int b = 0; b < nb; ++b) {
                if (cfg.voxel_bf16) *(unsigned*)((__nv_bfloat16*)vox + (size_t)b * P + p) = 0u;
                else *(float2*)((float*)vox + (size_t)b * P + p) = make_float2(0.f, 0.f);
            }
        }
    }
    __syncthreads();
    FX_STAMP(8);
    const int n_ent = min(s_cnt[3], s_cnt[0]);
    const int n_spill = s_cnt[1];
    // ================================ post pass over the table =====================================
    // Per entry: first active frame f (exact test, v2e_model.py:394-404), brightness before it,
    // "hot" flag (some frame is theta thresholds away from the brightness the entry starts with -
    // a bound no event count of the pixel exceeds by much; a scheduling hint: the hot entries
    // alone usually decide the per-frame maxima), and the STATIC entries, whose fate does not
    // depend on the maxima at all:
    //   * nothing happens (the stream test was conservative, or its shot marks went away);
    //   * exactly one event: a single frame crosses, by one threshold, it is the pixel's first
    //     event of the pack (timestamp_mem = -Tr: iteration 0 always passes the gate, and the
    //     frame's maximum is >= 1 because this very pixel counts 1), and with the brightness
    //     stepped by that event (:522) no later frame crosses.  Three entries in five.
    // Only the event's time stamp - the frame's first, hence its voxel weights - depends on the
    // guess, and it is the same for all of them: static entries never enter a resolve round;
    // they are counted per frame and polarity, their statistics follow from the counts, and the
    // commit writes their columns from the final frame constants.
    if (SHOT) {
        // ---- marked entries first: second stage of the shot test against the same upper bound.  A
        // marked frame has field 0, so its uniform is u2 / 1024 with u2 the draw the exact test will
        // use (fx_event_frame); four marks in five go away here instead of costing a resolve task an
        // out-of-line call every round.  The list is short (one pixel in a hundred): it is dealt
        // across the warps, so that no warp walks more than a few of these divergent chains.
        const int n_mark = s_cnt[7];
        constexpr int nwarps = kSpThreads / 32;
        for (int j = lane * nwarps + warp; j < n_mark; j += 32 * nwarps) {
            unsigned* ep = s_ent + s_perm[j];
            unsigned meta = ep[FX_META * cap];
            const unsigned cm0 = (meta >> 6) & 0x3ffu;
            unsigned cm = cm0;
            float u_first = 0.f;   // the uniform of the entry's first mark
            const int p = tile_px0 + (int)ep[FX_IDX * cap];
            const float cmax = fmaxf(__uint_as_float(ep[FX_POS * cap]), __uint_as_float(ep[FX_NEG * cap]));
            for (unsigned m = cm0; m != 0u; m &= m - 1u) {
                const int k = __ffs((int)m) - 1;
                const float tmax = (fmaxf(s_fr[k + 1].kp, s_fr[k + 1].kn) * 4.0f *
                                    (1.0f + fmaxf(cfg.shot_inten_m1, 0.f))) * cmax;
                // (fx_shot_uniform, inline: a call here would park the loop's registers in local
                // memory, an L2 round trip each with the L1 carved out as shared memory)
                const uint4 w = noise_words<R>(sq.philox_ctr, (unsigned)(k + 1), (unsigned)p, 1);
                const float u = (0.f + (float)w.x * (1.0f / 4294967296.0f)) * (1.0f / 1024.0f);
                if (!(u * 1024.0f < tmax)) cm &= ~(1u << k);
                if (k == __ffs((int)cm0) - 1) u_first = u;
            }
            meta = (meta & ~(0x3ffu << 6)) | (cm << 6);
            if (cm != 0u && cm == (cm0 & (0u - cm0))) {   // one mark left, and it is the first one
                // the usual survivor: one marked frame, the entry's first mark - the one whose
                // inten01 pass 1 left in FX_SI.  With the frame's uniform next to it, the exact test
                // (emulator_utils.py:173-188) runs inline in the resolve rounds instead of through an
                // out-of-line call each time.
                ep[FX_SU * cap] = __float_as_uint(u_first);
                meta |= kSpShotData;
            }
            ep[FX_META * cap] = meta;
        }
        __syncthreads();
        FX_STAMP(23);
    }
    for (int i = tid; i < n_ent; i += kSpThreads) {
        unsigned* ep = s_ent + i;
        const unsigned meta = ep[FX_META * cap];
        const unsigned cm = SHOT ? (meta >> 6) & 0x3ffu : 0u;
        const float pos = __uint_as_float(ep[FX_POS * cap]), neg = __uint_as_float(ep[FX_NEG * cap]);
        // One chain: up to the first active frame c is the brightness the pixel really has (no event
        // yet: base less the leak, :366-371); from there on it is the brightness it would have if
        // every crossing gave exactly one event (:522 with final = 1) - what the static test and the
        // busy hint need.  Before the first active frame a crossing IS the first active frame, so the
        // chain steps at every crossing, whatever came before.  What happened where is kept as one
        // bit per frame and sorted out after the loop.
        float c = __uint_as_float(ep[FX_BASE * cap]);
        const float base0 = c;
        const float denp = fadd(pos, 1e-9f), denn = fadd(neg, 1e-9f), nneg = -neg;
        const float tp = theta * pos, tn = theta * neg;
        unsigned xmask = 0u, umask = 0u, amask = 0u;   // crossing / diff > 0 / a polarity (|diff| > 1e-6)
        float sb = c, adf = 0.f, denf = 1.f;           // at the first active frame (so far: none)
        bool seen = false, hot = false;
#pragma unroll
        for (int k = 0; k < kSpNF; ++k) {
            const float nw = __uint_as_float(ep[(nw_off + k) * cap]);
            const float b = LEAK ? fsub(c, __uint_as_float(ep[(dl_off + k) * cap])) : c;
            const float diff = fsub(nw, b);
            const float ad = fabsf(diff);
            const bool up = diff > 0.f;
            const float den = up ? denp : denn;
            const bool cross = ad >= den;   // den > 1e-6: the tolerance mask (:394) holds
            if (!seen) { sb = c; adf = ad; denf = den; }
            if (cross) xmask |= 1u << k;
            if (up) umask |= 1u << k;
            if (SHOT && ad > 1e-6f) amask |= 1u << k;
            // the first active frame: a crossing, or a shot candidate with a polarity (:394-404,:447)
            seen = seen || cross || (SHOT && ((cm >> k) & 1u) && ad > 1e-6f);
            c = cross ? fadd(b, up ? pos : nneg) : b;   // :522 with one event: base + pol * C
            const float d0 = fsub(nw, base0);
            hot = hot || fabsf(d0) >= (d0 > 0.f ? tp : tn);
        }
        const unsigned act = xmask | (cm & amask);
        const bool idle = act == 0u;
        int f = idle ? 31 : __ffs((int)act);
        const unsigned later = idle ? 0u : (xmask >> f);   // crossings after the first active frame
        const bool up1 = !idle && ((umask >> (f - 1)) & 1u);
        // exactly one event: the first active frame crosses by one threshold, the pixel has no shot
        // mark and no event earlier in the pack (:469-472), and nothing crosses afterwards
        const bool one = !idle && ((xmask >> (f - 1)) & 1u) && adf < fadd(denf, denf) && cm == 0u && !(meta & 32u);
        // busy entries (an event in three frames or more) join the hot list: its frame-synchronous
        // straight-line loop beats the event-driven one when most frames hold an event
        hot = hot || __popc(later) >= kSpBusy;
        if (idle) sb = c;   // nothing happens: the entry only carries its state
        const bool stat = idle || (one && later == 0u);
        ep[FX_META * cap] = meta | (unsigned)f | (hot ? 0x10000u : 0u) | (stat ? kSpStatic : 0u) |
                            (up1 ? kSpStaticUp : 0u);
        ep[FX_BASE * cap] = __float_as_uint(sb);
        if (stat) {
            ep[FX_OB * cap] = __float_as_uint(c);
            if (!idle) atomicAdd(&s_sc[(up1 ? 0 : 16) + f], 1);
        } else {
            atomicAdd(&s_hist[(hot ? 32 : 0) + f], 1);
        }
    }
    __syncthreads();
    FX_STAMP(14);
    // ---- order the listed entries by first active frame (counting sort; order within a frame is free)
    if (tid == 0) {
        int o = 0;
        for (int i = 0; i < 64; ++i) {   // cold frames 0..31, then hot frames 0..31
            const int c = s_hist[i];
            s_hist[64 + i] = o;
            o += c;
            if (i == 31) s_cnt[4] = o;   // cold entries
        }
        s_cnt[6] = o;                    // entries that take part in the resolve rounds
    }
    if (tid >= 1 && tid < T) s_max[tid] = (s_sc[tid] + s_sc[16 + tid]) > 0 ? 1u : 0u;
    __syncthreads();
    FX_STAMP(15);
    const int n_cold = s_cnt[4], n_hot = s_cnt[6] - n_cold;
    for (int i = tid; i < n_ent; i += kSpThreads) {
        const unsigned meta = s_ent[FX_META * cap + i];
        if (meta & kSpStatic) continue;
        s_perm[atomicAdd(&s_hist[64 + ((meta & 0x10000u) ? 32 : 0) + (int)(meta & 31u)], 1)] = (unsigned short)i;
    }
    __syncthreads();
    FX_STAMP(2);
    if (tl && tid == 0) { tl[28] = n_ent; tl[29] = n_spill; tl[26] = n_hot; tl[25] = n_cold; }

    // ================================ resolve rounds ================================================
    // round 0 is a pre-pass over the hot entries only: under last pack's guess it yields the
    // maxima those entries reach, which are nearly always the image-wide ones - the guess of the
    // first full round.  Rounds >= 1 resolve every entry and test the guess they ran under.
    const float Tr = sq.Tr;
    int pass = 0;
    bool converged = false;
    for (; pass < T + 3 + kSpPre; ++pass) {
        const bool pre = pass < kSpPre;
        if (pre) {
            // The pre-pass only has to produce a good guess, and every warp of the grid waits for
            // its dependent chain of frames: one hot entry per lane, straight-line code, no voxel
            // accumulation, shot candidates ignored, every event frame taken as the usual single
            // event (gate on) or the plain count (gate off).
            for (int i0 = warp * 32; i0 < n_hot; i0 += kSpThreads) {
                const int i = i0 + lane;
                const bool on = i < n_hot;
                const unsigned* ep = s_ent + s_perm[n_cold + (on ? i : 0)];
                const unsigned meta = ep[FX_META * cap];
                float base = __uint_as_float(ep[FX_BASE * cap]);
                float tm = (meta & 32u) ? 0.f : -Tr;
                const float pos = __uint_as_float(ep[FX_POS * cap]);
                const float neg = __uint_as_float(ep[FX_NEG * cap]);
                const float denp = fadd(pos, 1e-9f), denn = fadd(neg, 1e-9f);
                const int f = on ? (int)(meta & 31u) : V2E_MAX_FRAMES;
                // fully unrolled: the table reads and the frame constants leave the chain
                float nwv[kSpNF], dlv[kSpNF], cmax[kSpNF];
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    nwv[k] = __uint_as_float(ep[(nw_off + k) * cap]);
                    dlv[k] = LEAK ? __uint_as_float(ep[(dl_off + k) * cap]) : 0.f;
                }
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    const int n = k + 1;
                    const bool act = n >= f;
                    const SpFrameEv ev = s_ev[n];
                    const bool gate = s_g[n].gate != 0;
                    const float b = LEAK ? fsub(base, dlv[k]) : base;
                    const float diff = fsub(nwv[k], b);
                    const float ad = fabsf(diff);
                    const bool up = diff > 0.f;
                    const float den = up ? denp : denn;
                    const bool cross = act && ad >= den;
                    if (act) base = b;
                    // event_counts (:404); an approximate quotient will do for a hint
                    const float cntf = cross ? floorf(__fdividef(ad, den)) : 0.f;
                    cmax[k] = cntf;
                    const float cf = fminf(cntf, ev.mxf);
                    const float tc = fadd(ev.t0, fmul(ev.step, cf));
                    const bool first = fsub(ev.ts1, tm) > Tr;
                    float jl = floorf((Tr + tm - ev.t0) * ev.rstep - 0.5f);
                    jl = fmaxf(fminf(jl, cf - 1.f), 1.f);
                    const float tsb = fadd(ev.t0, fmul(ev.step, jl + 1.f));
                    const float tsc = fadd(ev.t0, fmul(ev.step, jl + 2.f));
                    const float ts = first ? ev.ts1 : (fsub(tsb, tm) > Tr ? tsb : tsc);
                    const float C = up ? pos : neg;
                    const bool hit = cf >= 1.f && (!gate || fsub(tc, tm) > Tr);
                    const float step_c = gate ? (up ? C : -C) : fmul(fmul(up ? 1.f : -1.f, cf), C);
                    if (hit && gate) tm = ts;
                    if (hit) base = fadd(base, step_c);
                }
                // the warp's maxima: nine reductions back to back, then lane k publishes frame k+1
                unsigned mine = 0u;
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    const unsigned wmax = __reduce_max_sync(0xffffffffu, (unsigned)(int)cmax[k]);
                    if (lane == k) mine = wmax;
                }
                if (lane < kSpNF && mine) atomicMax(&s_max[lane + 1], mine);
            }
        } else {
            // tasks drawn from a queue: the hot list first (short tasks of kSpHotLanes entries:
            // a hot entry has events in most frames and a warp walks the union of its lanes'
            // paths, so fewer lanes per warp shorten the chain every round waits for), then the
            // cold list in tasks of 32; each list in the order of the first active frame, so that
            // a task only walks the frames its entries need
            const int th = (n_hot + kSpHotLanes - 1) / kSpHotLanes, tc = (n_cold + kSpColdLanes - 1) / kSpColdLanes;
            while (true) {
                int t = 0;
                if (lane == 0) t = atomicAdd(&s_cnt[5], 1);
                t = __shfl_sync(0xffffffffu, t, 0);
                if (t >= th + tc) break;
                int e;
                bool on;
                if (t < th) {
                    const int i = t * kSpHotLanes + lane;
                    on = lane < kSpHotLanes && i < n_hot;
                    e = s_perm[n_cold + (on ? i : 0)];
                } else {
                    const int i = (t - th) * kSpColdLanes + lane;
                    on = lane < kSpColdLanes && i < n_cold;
                    e = s_perm[on ? i : 0];
                }
#ifdef SP_EXP_STAMPS
                const long long task_t0 = fx_now();
                long long slow_t = 0;
#endif
                unsigned* ep = s_ent + e;
                const unsigned meta = ep[FX_META * cap];
                const int p = tile_px0 + (int)ep[FX_IDX * cap];
                const unsigned cm = SHOT ? (meta >> 6) & 0x1ffu : 0u;
                float base = __uint_as_float(ep[FX_BASE * cap]);
                float tm = (meta & 32u) ? 0.f : -Tr;
                const float pos = __uint_as_float(ep[FX_POS * cap]);
                const float neg = __uint_as_float(ep[FX_NEG * cap]);
                const float denp = fadd(pos, 1e-9f), denn = fadd(neg, 1e-9f);
                const int f = on ? (int)(meta & 31u) : V2E_MAX_FRAMES;
                float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f, v4 = 0.f;
                unsigned nev = 0;
                const int n_first = __reduce_min_sync(0xffffffffu, f);
                const unsigned* nwp = ep + (nw_off + n_first - 1) * cap;
                const bool has_cm = SHOT && __any_sync(0xffffffffu, cm != 0u);
                // shot candidates, several events in one frame, estimate off: the generic code
                // (out of line), accumulators via the entry's voxel words.  A macro, not a lambda:
                // a closure that is not inlined would pin base, tm and the accumulators to local
                // memory for the whole loop.
#ifdef SP_EXP_STAMPS
#define SP_T0 const long long t0_ = fx_now();
#define SP_T1 if (tl && pass == kSpPre) { slow_t += fx_now() - t0_; atomicMax((unsigned long long*)&tl[22], (unsigned long long)(fx_now() - t0_)); atomicAdd((unsigned long long*)&tl[21], (unsigned long long)(fx_now() - t0_)); }
#else
#define SP_T0
#define SP_T1
#endif
#define SP_SLOW_FRAME(n_, cntf_, up_, C_, bit_)                                                        \
    do {                                                                                               \
        SP_T0                                                                                          \
        float* vp_ = (float*)(ep + FX_V0 * cap);                                                       \
        vp_[0] = v0; vp_[cap] = v1; vp_[2 * cap] = v2; vp_[3 * cap] = v3; vp_[4 * cap] = v4;           \
        const unsigned long long r_ = fx_event_frame<FDT, R, false>(                                   \
            prm, sq, sm, (n_), p, (int)(cntf_), (up_) ? 1.f : -1.f, (C_), (bit_), SHOT, NOISE == 1, tm, \
            vp_, cap);                                                                                 \
        const int fin_ = (int)(r_ >> 32);                                                              \
        tm = __uint_as_float((unsigned)r_);                                                            \
        v0 = vp_[0]; v1 = vp_[cap]; v2 = vp_[2 * cap]; v3 = vp_[3 * cap]; v4 = vp_[4 * cap];           \
        if (fin_) base = fadd(base, fmul(fmul((up_) ? 1.f : -1.f, (float)fin_), (C_))); /* :522 */     \
        nev += (unsigned)fin_;                                                                         \
        SP_T1                                                                                          \
    } while (0)
                // a shot candidate whose inten01 and uniform are at hand: straight to the iteration
                // loop (what fx_event_frame does once it has decided `cand`)
#define SP_SHOT_FRAME(n_, cntf_, up_, C_, I_, u_)                                                       \
    do {                                                                                               \
        SP_T0                                                                                          \
        float* vp_ = (float*)(ep + FX_V0 * cap);                                                       \
        vp_[0] = v0; vp_[cap] = v1; vp_[2 * cap] = v2; vp_[3 * cap] = v3; vp_[4 * cap] = v4;           \
        const FrameGlobals g_ = s_g[(n_)];                                                             \
        const unsigned long long r_ = fx_events<R>(                                                    \
            &prm.cfg, &sq, &s_fr[(n_)], g_.mx, g_.ts_step, g_.gate, (n_), p, (up_) ? 1.f : -1.f, (C_), \
            min((int)(cntf_), g_.mx), (I_), tm, g_.mx > 0 ? 2 : 0, (u_), vp_, cap);                    \
        const int fin_ = (int)(r_ >> 32);                                                              \
        tm = __uint_as_float((unsigned)r_);                                                            \
        v0 = vp_[0]; v1 = vp_[cap]; v2 = vp_[2 * cap]; v3 = vp_[3 * cap]; v4 = vp_[4 * cap];           \
        if (fin_) base = fadd(base, fmul(fmul((up_) ? 1.f : -1.f, (float)fin_), (C_))); /* :522 */     \
        nev += (unsigned)fin_;                                                                         \
        SP_T1                                                                                          \
    } while (0)
                // the same frame through the small shot loop (first firing and probability known)
#define SP_SHOT_LOOP(n_, cntf_, up_, C_, gfirst_, pit_)                                                 \
    do {                                                                                               \
        SP_T0                                                                                          \
        float* vp_ = (float*)(ep + FX_V0 * cap);                                                       \
        vp_[0] = v0; vp_[cap] = v1; vp_[2 * cap] = v2; vp_[3 * cap] = v3; vp_[4 * cap] = v4;           \
        const FrameGlobals g_ = s_g[(n_)];                                                             \
        const unsigned long long r_ = sp_shot_loop<R>(                                                 \
            sq.philox_ctr, (n_), p, s_fr[(n_)].t_bin_prev, g_.ts_step, Tr, g_.gate, g_.mx,             \
            min((int)(cntf_), g_.mx), (int)(gfirst_), (pit_), (up_) ? 1.f : -1.f, tm, vp_, cap);       \
        const int fin_ = (int)(r_ >> 32);                                                              \
        tm = __uint_as_float((unsigned)r_);                                                            \
        v0 = vp_[0]; v1 = vp_[cap]; v2 = vp_[2 * cap]; v3 = vp_[3 * cap]; v4 = vp_[4 * cap];           \
        if (fin_) base = fadd(base, fmul(fmul((up_) ? 1.f : -1.f, (float)fin_), (C_))); /* :522 */     \
        nev += (unsigned)fin_;                                                                         \
        SP_T1                                                                                          \
    } while (0)
                const bool shot_data = SHOT && (meta & kSpShotData);
                const float shot_I = shot_data ? __uint_as_float(ep[FX_SI * cap]) : 0.f;
                const float shot_u = shot_data ? __uint_as_float(ep[FX_SU * cap]) : 0.f;
                if (t < th) {
                    // ---- hot entries: an event in most frames, the chain every round waits for;
                    // straight-line code for every lane, predicated updates
#pragma unroll 1
                    for (int n = n_first; n < T; ++n, nwp += cap) {
                        const bool act = n >= f;
                        const SpFrameEv ev = s_ev[n];
                        const bool gate = s_g[n].gate != 0;
                        const float b = LEAK ? fsub(base, __uint_as_float(nwp[kSpNF * cap])) : base;
                        const float diff = fsub(__uint_as_float(nwp[0]), b);
                        const float ad = fabsf(diff);
                        const bool up = diff > 0.f;
                        const float den = up ? denp : denn;
                        const bool cross = act && ad >= den;   // den > 1e-6: the tolerance mask (:394) holds
                        const bool bit = has_cm && act && ((cm >> (n - 1)) & 1u) && ad > 1e-6f;
                        if (act) base = b;
                        const float cntf = cross ? floorf(fdiv(ad, den)) : 0.f;   // event_counts (:404)
                        const unsigned wmax = __reduce_max_sync(0xffffffffu, (unsigned)(int)cntf);
                        if (wmax > s_max[n] && lane == 0) atomicMax(&s_max[n], wmax);
                        const float C = up ? pos : neg;
                        // the marked frame of an entry with shot data: the exact test, inline
                        bool shot_ev = false, bit_g = bit;
                        float fI = shot_I, fu = shot_u;
                        bool have = bit && shot_data;
                        if (has_cm && __any_sync(0xffffffffu, bit && !shot_data && !(meta & 0x20000u))) {
                            if (bit && !shot_data && !(meta & 0x20000u))
                                have = sp_mark_data<FDT, R>(sq, s_lut, P, n, p, fI, fu);
                        }
                        if (have) {
                            const float thr = fx_shot_thr(up ? s_fr[n].kp : s_fr[n].kn, cfg.shot_inten_m1, fI, C);
                            shot_ev = fu * 1024.0f < thr;
                            bit_g = false;
                        }
                        SpEv1 r = sp_classify(ev, gate, cntf, bit_g, tm, Tr);
                        float xts = -1.f;   // stamp of a shot event that follows the pixel's own events
                        if (__any_sync(0xffffffffu, r.slow || shot_ev)) {
                            if (shot_ev) {
                                const SpShot sh = sp_shot_lean<R>(cfg, sq, s_sa[n], (sq.shot_f64_mask >> n) & 1u,
                                                                  ev, n, p, up, C, cntf, fI, fu);
                                // (kind 0: r stands - it was formed as for an unmarked frame)
                                if (sh.kind == 1 && !r.slow) {
                                    xts = sh.ts;
                                } else if (sh.kind != 0) {
                                    r.slow = false;
                                    r.one = false;
#ifdef SP_EXP_STAMPS
                                    if (tl && pass == kSpPre) atomicAdd((unsigned long long*)&tl[20], 1ull << (16 * (sh.kind == 2 ? (int)sh.ts : 1)));
#endif
                                    if (sh.kind == 2 && sh.ts < 1.f) SP_SHOT_FRAME(n, cntf, up, C, fI, fu);
                                    else SP_SHOT_LOOP(n, cntf, up, C, sh.g, sh.p_it);
                                }
                            }
                            if (r.slow) {
                                r.one = false;
#ifdef SP_EXP_STAMPS
                                if (tl && pass == kSpPre) atomicAdd((unsigned long long*)&tl[20], 1ull << (bit ? 48 : 56));
#endif
                                SP_SLOW_FRAME(n, cntf, up, C, bit);
                            }
                        }
                        const float base_pre = base;
                        // exactly one event, at ts (:478-502); pol * w is w or -w
                        const float bbf = r.first ? ev.bbf : floorf(r.ts);
                        float wr = r.first ? ev.wr : fsub(r.ts, bbf);
                        float wl = r.first ? ev.wl : fsub(1.0f, wr);
                        if (!up) { wl = -wl; wr = -wr; }
                        if (r.one && bbf == 0.f) v0 = fadd(v0, wl);
                        if (r.one && bbf == 1.f) v1 = fadd(v1, wl);
                        if (r.one && bbf == 2.f) v2 = fadd(v2, wl);
                        if (r.one && bbf == 3.f) v3 = fadd(v3, wl);
                        if (r.one && bbf == 4.f) v4 = fadd(v4, wl);
                        if (r.one && bbf == 0.f) v1 = fadd(v1, wr);
                        if (r.one && bbf == 1.f) v2 = fadd(v2, wr);
                        if (r.one && bbf == 2.f) v3 = fadd(v3, wr);
                        if (r.one && bbf == 3.f) v4 = fadd(v4, wr);
                        if (r.one) {
                            if (gate) tm = r.ts;
                            base = fadd(base, up ? C : -C);   // :522: (pol * 1) * C
                            nev += 1u;
                        }
                        if (__any_sync(0xffffffffu, xts >= 0.f)) {
                            // the shot event, after the pixel's own (:449-502 at a later iteration)
                            if (xts >= 0.f && (!gate || fsub(xts, tm) > Tr)) {
                                if (gate) tm = xts;
                                const float xb = floorf(xts);
                                float xr = fsub(xts, xb);
                                float xl = fsub(1.0f, xr);
                                if (!up) { xl = -xl; xr = -xr; }
                                if (xb == 0.f) { v0 = fadd(v0, xl); v1 = fadd(v1, xr); }
                                else if (xb == 1.f) { v1 = fadd(v1, xl); v2 = fadd(v2, xr); }
                                else if (xb == 2.f) { v2 = fadd(v2, xl); v3 = fadd(v3, xr); }
                                else if (xb == 3.f) { v3 = fadd(v3, xl); v4 = fadd(v4, xr); }
                                else if (xb == 4.f) { v4 = fadd(v4, xl); }
                                // :522 with final = 1 or 2: (pol * final) * C, 2C being exact
                                const float st = r.one ? fadd(C, C) : C;
                                base = fadd(base_pre, up ? st : -st);
                                nev += 1u;
                            }
                        }
                    }
                } else {
                    // ---- cold entries: a few events per entry, most frames quiet.  Event-driven:
                    // every lane scans its own frames to its next crossing (a handful of
                    // instructions per frame), then the lanes that found one handle it together,
                    // each at its own frame - a warp pays the event code once per event of its
                    // busiest entry instead of once per frame.
                    int n = f;   // this lane's frame cursor (>= T: nothing left)
                    const unsigned* nwq = ep + (nw_off - 1) * cap;   // [n * cap]: frame n
                    while (true) {
                        float ad = 0.f, den = 1.f;
                        bool up = false, bit = false;
                        while (n < T) {
                            const float b = LEAK ? fsub(base, __uint_as_float(nwq[(n + kSpNF) * cap])) : base;
                            const float diff = fsub(__uint_as_float(nwq[n * cap]), b);
                            ad = fabsf(diff);
                            up = diff > 0.f;
                            den = up ? denp : denn;
                            base = b;
                            bit = has_cm && ((cm >> (n - 1)) & 1u) && ad > 1e-6f;
                            if (ad >= den || bit) break;   // den > 1e-6: the tolerance mask (:394) holds
                            ++n;
                        }
                        if (!__any_sync(0xffffffffu, n < T)) break;
                        if (n < T) {
                            const bool cross = ad >= den;
                            float xts = -1.f;   // stamp of a shot event that follows the pixel's own events
                            float xg = 0.f, xp = 0.f;   // ... its iteration and the per-iteration probability
                            float fI = shot_I, fu = shot_u;
                            bool have = bit && shot_data;
                            if (bit && !shot_data && !(meta & 0x20000u))
                                have = sp_mark_data<FDT, R>(sq, s_lut, P, n, p, fI, fu);
                            if (have) {
                                // a marked frame: the exact test, inline
                                const float C = up ? pos : neg;
                                const float thr = fx_shot_thr(up ? s_fr[n].kp : s_fr[n].kn, cfg.shot_inten_m1, fI, C);
                                if (fu * 1024.0f < thr) {
                                    float cq = 0.f;
                                    if (cross) cq = ad < fadd(den, den) ? 1.f : floorf(fdiv(ad, den));
                                    if ((unsigned)(int)cq > s_max[n]) atomicMax(&s_max[n], (unsigned)(int)cq);
                                    const SpFrameEv evs = s_ev[n];
                                    const SpShot sh = sp_shot_lean<R>(cfg, sq, s_sa[n], (sq.shot_f64_mask >> n) & 1u,
                                                                      evs, n, p, up, C, cq, fI, fu);
                                    if (sh.kind == 2) {
#ifdef SP_EXP_STAMPS
                                        if (tl && pass == kSpPre) atomicAdd((unsigned long long*)&tl[20], (1ull << (16 * (int)sh.ts)) << 8);
#endif
                                        if (sh.ts < 1.f) SP_SHOT_FRAME(n, cq, up, C, fI, fu);
                                        else SP_SHOT_LOOP(n, cq, up, C, sh.g, sh.p_it);
                                        ++n;
                                        continue;
                                    }
                                    if (sh.kind == 1) { xts = sh.ts; xg = sh.g; xp = sh.p_it; }
                                    // (kind 0: as if unmarked)
                                }
                                bit = false;
                                if (!cross && xts < 0.f) { ++n; continue; }   // not a candidate after all, and no crossing
                            }
                            // event_counts (:404): 1 below twice the denominator (exact), else the quotient
                            const bool big = cross && !(ad < fadd(den, den));
                            float cntf = cross ? 1.f : 0.f;
                            if (big) cntf = floorf(fdiv(ad, den));
                            if ((unsigned)(int)cntf > s_max[n]) atomicMax(&s_max[n], (unsigned)(int)cntf);
                            const SpFrameEv ev = s_ev[n];
                            const bool gate = s_g[n].gate != 0;
                            const float C = up ? pos : neg;
                            SpEv1 r;
                            r.one = false; r.slow = false; r.first = true; r.ts = ev.ts1;
                            bool simple = false;
                            if (cross || bit) {
                                // iteration 0 passes and none after it does (with the gate on nothing
                                // passes within Tr of it: the last iteration, c-1, decides)
                                simple = !bit && ev.mxf >= 1.f;
                                if (gate) {
                                    simple = simple && fsub(ev.ts1, tm) > Tr;
                                    if (big) {
                                        const float tc = fadd(ev.t0, fmul(ev.step, fminf(cntf, ev.mxf)));
                                        simple = simple && !(fsub(tc, ev.ts1) > Tr);
                                    }
                                } else {
                                    simple = simple && !big;
                                }
                                r.one = simple;
                                if (!simple) r = sp_classify(ev, gate, cntf, bit, tm, Tr);
                            }
#ifdef SP_EXP_STAMPS
                            if (tl && pass == kSpPre)
                                atomicAdd((unsigned long long*)&tl[24], r.slow ? (bit ? 1ull : (1ull << 16)) : (simple ? (1ull << 48) : (1ull << 32)));
#endif
                            if (r.slow) {
                                // (a pending shot event goes through the generic code with the rest)
                                if (xts >= 0.f) SP_SHOT_LOOP(n, cntf, up, C, xg, xp);
                                else SP_SLOW_FRAME(n, cntf, up, C, bit);
                            } else {
                                const float base_pre = base;
                                if (r.one) {
                                    // exactly one event, at ts (:478-502); pol * w is w or -w
                                    if (gate) tm = r.ts;
                                    const float bbf = r.first ? ev.bbf : floorf(r.ts);
                                    float wr = r.first ? ev.wr : fsub(r.ts, bbf);
                                    float wl = r.first ? ev.wl : fsub(1.0f, wr);
                                    if (!up) { wl = -wl; wr = -wr; }
                                    if (bbf == 0.f) { v0 = fadd(v0, wl); v1 = fadd(v1, wr); }
                                    else if (bbf == 1.f) { v1 = fadd(v1, wl); v2 = fadd(v2, wr); }
                                    else if (bbf == 2.f) { v2 = fadd(v2, wl); v3 = fadd(v3, wr); }
                                    else if (bbf == 3.f) { v3 = fadd(v3, wl); v4 = fadd(v4, wr); }
                                    else if (bbf == 4.f) { v4 = fadd(v4, wl); }
                                    base = fadd(base, up ? C : -C);   // :522: (pol * 1) * C
                                    nev += 1u;
                                }
                                if (xts >= 0.f && (!gate || fsub(xts, tm) > Tr)) {
                                    // the shot event, after the pixel's own (:449-502 at a later iteration)
                                    if (gate) tm = xts;
                                    const float xb = floorf(xts);
                                    float xr = fsub(xts, xb);
                                    float xl = fsub(1.0f, xr);
                                    if (!up) { xl = -xl; xr = -xr; }
                                    if (xb == 0.f) { v0 = fadd(v0, xl); v1 = fadd(v1, xr); }
                                    else if (xb == 1.f) { v1 = fadd(v1, xl); v2 = fadd(v2, xr); }
                                    else if (xb == 2.f) { v2 = fadd(v2, xl); v3 = fadd(v3, xr); }
                                    else if (xb == 3.f) { v3 = fadd(v3, xl); v4 = fadd(v4, xr); }
                                    else if (xb == 4.f) { v4 = fadd(v4, xl); }
                                    // :522 with final = 1 or 2: (pol * final) * C, 2C being exact
                                    const float st = r.one ? fadd(C, C) : C;
                                    base = fadd(base_pre, up ? st : -st);
                                    nev += 1u;
                                }
                            }
                            ++n;
                        }
                    }
                }
#ifdef SP_EXP_STAMPS
                if (tl && pass == kSpPre) {
                    // the longest task of the round: duration, queue index, time in generic calls
                    for (int o = 16; o > 0; o >>= 1) slow_t = max(slow_t, __shfl_xor_sync(0xffffffffu, slow_t, o));
                    if (lane == 0) {
                        const unsigned long long d = (unsigned long long)(fx_now() - task_t0);
                        atomicMax((unsigned long long*)&tl[16], (d << 32) | ((unsigned long long)(slow_t >> 4) << 8) | (unsigned)t);
                        tl[17] = th; tl[18] = tc;
                        if (t == 0) tl[19] = fx_now();
                    }
                }
#endif
                if (on) {
                    ep[FX_OB * cap] = __float_as_uint(base);
                    ep[FX_OT * cap] = __float_as_uint(tm);
                    ep[(FX_V0 + 0) * cap] = __float_as_uint(v0);
                    ep[(FX_V0 + 1) * cap] = __float_as_uint(v1);
                    ep[(FX_V0 + 2) * cap] = __float_as_uint(v2);
                    ep[(FX_V0 + 3) * cap] = __float_as_uint(v3);
                    ep[(FX_V0 + 4) * cap] = __float_as_uint(v4);
                }
                // statistics of the task's voxel columns (exact integers, order-free): summed over
                // the warp and added to the tile's partials here, so that no accumulator lives
                // across the task loop and no separate sweep follows the last task
                {
                    FxAcc a;
                    a.nnz = 0; a.nev = nev; a.s1 = 0; a.lo = 0; a.hi = 0;
                    fx_stat(v0, a.nnz, a.s1, a.lo, a.hi);
                    fx_stat(v1, a.nnz, a.s1, a.lo, a.hi);
                    fx_stat(v2, a.nnz, a.s1, a.lo, a.hi);
                    fx_stat(v3, a.nnz, a.s1, a.lo, a.hi);
                    fx_stat(v4, a.nnz, a.s1, a.lo, a.hi);
                    fx_flush(a, s_acc);
                }
            }
#ifdef SP_EXP_STAMPS
            if (tl && pass == kSpPre && lane == 0) atomicMax((unsigned long long*)&tl[15], (unsigned long long)fx_now());
#endif
            if (tid >= 64 && tid < 64 + 2 * kSpNF) {
                // the single-event entries: per frame cp ON and cn OFF events, all at the frame's
                // first stamp, i.e. values +-wl in bin bb and +-wr in bin bb+1 (if there is one);
                // the same exact integers a sweep would add for each of them.  One thread per
                // (frame, side).
                const int n = 1 + (tid - 64) % kSpNF, side = (tid - 64) / kSpNF;
                const long long cp = s_sc[n], cn = s_sc[16 + n];
                const unsigned long long c = (unsigned long long)(cp + cn);
                const SpFrameEv ev = s_ev[n];
                const float w = side ? ev.wr : ev.wl;
                if (c && side == 0) atomicAdd(&s_acc[4], c);   // events
                if (c && w != 0.f && !(side && !(ev.bbf + 1.f < (float)nb))) {
                    const int ip = __float2int_rz(w);
                    const int fp = __float2int_rz((w - (float)ip) * 2147483648.0f);
                    const unsigned long long k = ((unsigned long long)ip << 31) + (unsigned long long)fp;   // fx_stat's k, w > 0
                    const unsigned long long a2lo = k * k, a2hi = __umul64hi(k, k);
                    const unsigned long long lo = a2lo * c;
                    const unsigned long long hi = a2hi * c + __umul64hi(a2lo, c);
                    atomicAdd(&s_acc[0], c);
                    atomicAdd(&s_acc[1], (unsigned long long)((long long)k * (cp - cn)));
                    const unsigned long long old = atomicAdd(&s_acc[2], lo);
                    atomicAdd(&s_acc[3], hi + (old + lo < old ? 1ull : 0ull));
                }
            }
        }
        if (n_spill && !pre)
            fx_resolve_spill<FDT, R, false, true>(prm, sq, sm, spill_list, n_spill, tile_px0, spill_vox,
                                                  spill_state, LEAK, SHOT, NOISE == 1, CUTOFF);
        __syncthreads();
        if (pass < 4) FX_STAMP(3 + 3 * pass);
        if (pre) {
            // ---- pre-pass barrier: one self-validating word per tile (FxWork::hint) ----------------
            unsigned long long* slots = fx.hint[pass];
            if (tid == 0) {
                unsigned long long w = (unsigned long long)(pass + 1) << 54;
#pragma unroll
                for (int n = 1; n < T; ++n) w |= (unsigned long long)min(s_max[n], 63u) << (6 * (n - 1));
                atomicOr(slots + tile, w);   // (an atomic reaches the L2 at once; a plain store lingers)
            }
            if (warp == 0) {
                // one warp polls (a poller per tile in every CTA floods the L2 with requests)
                // (a lane's words are loaded together, then checked: one L2 round trip per sweep,
                // not one per word)
                constexpr int kPer = (kRaMaxTiles + 31) / 32;
                unsigned long long w[kPer];
                const unsigned long long tag = (unsigned long long)(pass + 1);
                int spins = 0;
                while (true) {
                    bool ok = true;
#pragma unroll
                    for (int j = 0; j < kPer; ++j) {
                        const int i = lane + 32 * j;
                        w[j] = i < prm.tiles_per_seq ? fx_ld_relaxed64(slots + i) : (tag << 54);
                    }
#pragma unroll
                    for (int j = 0; j < kPer; ++j) ok = ok && (w[j] >> 54) == tag;
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (++spins > kFxSpin) {
                        s_cnt[2] = 1;
#pragma unroll
                        for (int j = 0; j < kPer; ++j) if ((w[j] >> 54) != tag) w[j] = 0ull;
                        break;
                    }
                }
                unsigned m[kSpNF];
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    m[k] = 0u;
#pragma unroll
                    for (int j = 0; j < kPer; ++j) m[k] = max(m[k], (unsigned)(w[j] >> (6 * k)) & 63u);
                }
#pragma unroll
                for (int k = 0; k < kSpNF; ++k) {
                    const unsigned r = __reduce_max_sync(0xffffffffu, m[k]);
                    if (lane == 0) s_gmax[k + 1] = r;
                }
            }
        } else {
        // ---- grid barrier of this sequence: accumulate (maxima, partials), arrive, wait, read ------
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
                const unsigned target = (unsigned)prm.tiles_per_seq * (unsigned)(pass - kSpPre + 1);
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
        if (tid >= 1 && tid < T) {
            // after the pre-pass a frame no hot entry fired in keeps last pack's guess
            const int mx = (pre && s_gmax[tid] == 0) ? s_g[tid].mx : (int)s_gmax[tid];
            set_globals(tid, mx);
        }
        if (tid < V2E_MAX_FRAMES) {
            s_max[tid] = (tid >= 1 && tid < T && (s_sc[tid] + s_sc[16 + tid]) > 0) ? 1u : 0u;
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
            if (pk == prm.n_chain - 1) fx0.guess[tid] = s_gmax[tid] + 1u;
        } else {
            if (pk == prm.n_chain - 1) fx0.guess[tid] = 0u;
        }
    }
    __syncthreads();
    // ---- leave the barrier words zeroed: the last CTA of the sequence to report cleans up ----------
    // The CTA is done with this pack's barrier words: warp 0 reports that (and cleans up if the
    // CTA is the sequence's last to get here) while the other warps commit - the atomic's round
    // trip is off the path every pack ends with.  (Warp 0: its lane 0 is the thread that arrives
    // at the grid barriers of the next pack, so its fence after the clean-up orders the zeroed
    // words before that arrival.)
    constexpr int kCommitThreads = kSpThreads - 32;
    const int ctid = tid - 32;
    if (warp == 0) {
        unsigned last = 0;
        if (lane == 0) {
            __threadfence();
            last = atomicAdd(&fx.exit_count, 1u) == (unsigned)prm.tiles_per_seq - 1u;
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
            const int used = min(pass + 1, kFxMaxRounds);
            for (int i = lane; i < used * V2E_MAX_FRAMES; i += 32) (&fx.gmax[0][0])[i] = 0u;
            for (int i = lane; i < used * 8; i += 32) (&fx.stat[0][0])[i] = 0ull;
            for (int i = lane; i < kSpPre * kRaMaxTiles; i += 32) (&fx.hint[0][0])[i] = 0ull;
            __syncwarp();
            if (lane == 0) {
                __threadfence();
                fx.exit_count = 0;
                fx.arrive = 0;
                if (prm.timeline && seq == 0) {
                    prm.timeline[160 * 32 + 2 * (fx0.pad[0] & 127u) + 1] = fx_now();
                    prm.timeline[160 * 32 + 256 + (fx0.pad[0] & 127u)] = pass + 1;   // rounds, pre-pass included
                    fx0.pad[0] += 1;
                }
            }
        }
    } else {
        const float mean = s_stat[0], den = s_stat[1];
        const bool norm = s_stat[2] != 0.f;
        for (int i = ctid; i < n_ent; i += kCommitThreads) {
            const unsigned* ep = s_ent + i;
            const unsigned meta = ep[FX_META * cap];
            const int p = tile_px0 + (int)ep[FX_IDX * cap];
            sqc.base[p] = __uint_as_float(ep[FX_OB * cap]);
            float vv[nb];
            if (meta & kSpStatic) {
                // static entries: the one event (if any) sits at its frame's first stamp
                float tm = (meta & 32u) ? 0.f : -Tr;
                const int f = (int)(meta & 31u);
#pragma unroll
                for (int b = 0; b < nb; ++b) vv[b] = 0.f;
                if (f < T) {
                    const SpFrameEv ev = s_ev[f];
                    if (s_g[f].gate) tm = ev.ts1;
                    const bool up = (meta & kSpStaticUp) != 0u;
                    const float wl = up ? ev.wl : -ev.wl, wr = up ? ev.wr : -ev.wr;
#pragma unroll
                    for (int b = 0; b < nb; ++b) {
                        if (ev.bbf == (float)b) vv[b] = fadd(0.f, wl);
                        if (ev.bbf + 1.f == (float)b) vv[b] = fadd(0.f, wr);
                    }
                }
                sqc.tmem[p] = tm;
            } else {
                sqc.tmem[p] = __uint_as_float(ep[FX_OT * cap]);
#pragma unroll
                for (int b = 0; b < nb; ++b) vv[b] = __uint_as_float(ep[(FX_V0 + b) * cap]);
            }
#pragma unroll
            for (int b = 0; b < nb; ++b) {
                float v = vv[b];
                if (v != 0.f || !norm) {
                    if (norm) v = fdiv(fsub(v, mean), den);   // nz * (v - mean) / (std + 1e-8)
                    fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, v);
                }
            }
        }
        for (int e = ctid; e < n_spill; e += kCommitThreads) {
            const int p = tile_px0 + (int)spill_list[e];
            sq.base[p] = spill_state[p];
            if (CUTOFF) sq.lp[p] = spill_state[(size_t)P + p];
            sq.tmem[p] = spill_state[(size_t)2 * P + p];
            for (int b = 0; b < nb; ++b) {
                float v = spill_vox[(size_t)b * P + p];
                if (norm) v = fdiv(fmul(v != 0.f ? 1.f : 0.f, fsub(v, mean)), den);
                fx_store_vox(cfg, sq.voxel, (size_t)b * P + p, v);
            }
        }
    }
    FX_STAMP(31);
    if (tl && pk == min(2, prm.n_chain - 1)) {   // (a chain reports its third pack: one in steady state)
        __syncthreads();
        if (tid < 32) tl_out[tid] = tl[tid];
    }
    __syncthreads();   // (the next pack of the chain re-initialises the header)
    }
}

// ---- host side -----------------------------------------------------------------------------------
template <int FDT, int R, int NOISE, bool CUTOFF, bool ONE>
static int launch_spec_1(const FxParams& prm, int grid, size_t smem, int optin, cudaStream_t st) {
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_fixpoint_spec<FDT, R, NOISE, CUTOFF, ONE>;
    int rc = V2E_OK;
    if (!attr_set) {
        rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        if (rc != V2E_OK) return rc;
        attr_set = true;
    }
    void* args[] = {(void*)&prm};
    rc = cuda_ok(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)grid), dim3(kSpThreads),
                                             args, smem, st));
    count_launch();
    return rc;
}

template <int FDT, int R, int NOISE, bool CUTOFF>
static int launch_spec_v(const FxParams& prm, int grid, size_t smem, int optin, cudaStream_t st) {
    return prm.fr_in_params ? launch_spec_1<FDT, R, NOISE, CUTOFF, true>(prm, grid, smem, optin, st)
                            : launch_spec_1<FDT, R, NOISE, CUTOFF, false>(prm, grid, smem, optin, st);
}

template <int FDT>
static int launch_spec_t(const FxParams& prm, int noise, bool cutoff, int rounds, int grid,
                         size_t smem, int optin, cudaStream_t st) {
#ifdef V2E_FX_DEV   // development builds: one Philox depth only (fewer instantiations to compile)
#define SP_BY_ROUNDS(C) launch_spec_v<FDT, 10, 1, C>(prm, grid, smem, optin, st)
#else
#define SP_BY_ROUNDS(C)                                                           \
    (rounds == 7 ? launch_spec_v<FDT, 7, 1, C>(prm, grid, smem, optin, st)        \
                 : launch_spec_v<FDT, 10, 1, C>(prm, grid, smem, optin, st))
#endif
    (void)rounds;
    if (noise == 0)
        return cutoff ? launch_spec_v<FDT, 10, 0, true>(prm, grid, smem, optin, st)
                      : launch_spec_v<FDT, 10, 0, false>(prm, grid, smem, optin, st);
    return cutoff ? SP_BY_ROUNDS(true) : SP_BY_ROUNDS(false);
#undef SP_BY_ROUNDS
}

int launch_fixpoint_spec(const FxParams& prm, int frame_dtype, int noise, bool cutoff, int rounds,
                         int grid, size_t smem, int optin, cudaStream_t st) {
    switch (frame_dtype) {
#ifndef V2E_FX_DEV
        case V2E_DT_U8: return launch_spec_t<V2E_DT_U8>(prm, noise, cutoff, rounds, grid, smem, optin, st);
        case V2E_DT_BF16: return launch_spec_t<V2E_DT_BF16>(prm, noise, cutoff, rounds, grid, smem, optin, st);
#else
        case V2E_DT_U8: case V2E_DT_BF16: return V2E_ERR_UNSUPPORTED;
#endif
        default: return launch_spec_t<V2E_DT_F32>(prm, noise, cutoff, rounds, grid, smem, optin, st);
    }
}

}  // namespace v2e
