// v2e_fixpoint.cuh — declarations shared by the two translation units of the fixed-point engine:
// v2e_fixpoint.cu (generic kernel: any pack shape, injected noise, debug maps; host side) and
// v2e_fixpoint_spec.cu (the kernel specialised for the reference's pack shape).
#pragma once
#include <cstdlib>

#include "v2e_engine.cuh"

namespace v2e {

#ifndef V2E_FX_THREADS
#define V2E_FX_THREADS 768
#endif
constexpr int kFxThreads = V2E_FX_THREADS;
constexpr int kFxWarps = kFxThreads / 32;
constexpr int kFxChunk = 9;         // frames per unrolled chunk = frames per Philox block
constexpr int kFxHeader = 10240;    // LUT, frame tables, reduction scratch, entry order, histogram
constexpr int kFxPermOff = 4096;    // byte offset of the entry order (u16[kFxPermMax]) in the header
constexpr int kFxPermMax = 2816;    // most entries a tile's table may hold; 512 B of histogram follow
constexpr int kFxSpin = 1 << 25;    // barrier poll limit (a lost CTA must not hang the GPU)
// Specialised variants (SPEC): the reference's pack shape (10 frames, 5 bins) with everything the
// address arithmetic needs known at compile time, for the 227 KB of shared memory a B200 CTA gets.
#ifndef V2E_FX_SPEC_SMEM
#define V2E_FX_SPEC_SMEM 232448
#endif
constexpr int kFxSpecT = 10, kFxSpecNb = 5, kFxSpecSmem = V2E_FX_SPEC_SMEM;

// entry words (structure of arrays: word w of entry e lives at s_ent[w * cap + e]).  After the
// fixed words come num_bins voxel accumulators, then per frame k = n-1 < 10 the two inputs of the
// pixel that no event can change: nw (the low-passed log intensity the frame presents,
// v2e_model.py:375) and, with leak on, dl (the leak decrement of that frame,
// emulator_utils.py:117-125).
enum {
    FX_META = 0,   // [4:0] first active frame f, [5] timestamp_mem == 0 (else -Tr), [15:6] frames
                   // that may hold a shot-noise event (conservative first-stage test)
    FX_IDX,        // tile-local pixel index
    FX_BASE,       // base_log_frame before frame f
    FX_POS, FX_NEG,
    FX_OB, FX_OT,  // results: base_log_frame, timestamp_mem
    FX_SI, FX_SU,  // specialised kernel: an entry's one shot-candidate frame - inten01 of that
                   // frame and the frame's uniform (post pass; valid when META says so)
    FX_V0          // results: num_bins voxel accumulators
};

// one sequence's launch record: device descriptor + per-frame scalars
struct FxSeqBlock {
    SeqDev sq;
    FrameSc fr[V2E_MAX_FRAMES];
};
// launch records in the workspace: a circular buffer of kFxBlkSlots records; a launch takes
// n_chain * n_seq consecutive ones (copied in stream order, so the copy for a launch never
// overtakes the launch before it; consecutive launches use disjoint ranges whenever two fit)
static_assert(sizeof(FxSeqBlock) <= 1280 && kFxBlkSlots * 1280 <= kFxBlkBytes,
              "workspace ring too small (v2e_engine.cuh, kFxBlkBytes)");

// Kernel parameters stay well under 4 KB: the launch path for larger parameter blocks costs
// ~20 us per launch on this driver (measured: scripts/clip_probe.py), more than a barrier pass.
struct FxParams {
    CfgDev cfg;
    int n_seq;
    int tiles_per_seq;
    int tile_px;       // pixels per tile, multiple of 32
    int cap;           // compact entries per tile, multiple of 32
    int nw_off, dl_off;  // entry word offsets of the per-frame inputs
    FxWork* fx;        // [n_seq]
    unsigned* spill_list;  // [n_seq][P]   tile-local indices of spilled pixels, per tile region
    float* spill_vox;      // [n_seq][nb][P] voxel accumulators of spilled pixels
    float* spill_state;    // [n_seq][3][P]  base, lp, tmem results of spilled pixels
    long long* timeline;   // development aid (V2E_TIMELINE=1): [cta][32] globaltimer stamps
    int fr_in_params;  // one.fr and the plane pointers of one.sq are valid for every pack of this launch
    int n_chain;       // consecutive packs of every sequence served by this launch (records [pack][seq])
    const FxSeqBlock* blocks;   // n_seq > 1 or a chain: records in the workspace (copied there in stream order)
    FxSeqBlock one;             // n_seq == 1: the record travels with the parameters
};
static_assert(sizeof(FxParams) < 4000, "keep the kernel parameters on the fast launch path");
static_assert(sizeof(SeqDev) <= 240, "SeqDev must fit its shared-memory slot");

__host__ __device__ constexpr int fx_spec_ewords(bool leak) {
    return FX_V0 + kFxSpecNb + (kFxSpecT - 1) * (leak ? 2 : 1);
}
__host__ __device__ constexpr int fx_spec_cap(bool leak) {
    return ((kFxSpecSmem - kFxHeader) / (4 * fx_spec_ewords(leak))) & ~31;
}
static_assert(fx_spec_cap(false) <= kFxPermMax && fx_spec_cap(true) <= kFxPermMax, "entry order table");

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

__device__ __forceinline__ unsigned long long fx_ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fx_st_relaxed64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void fx_red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ long long fx_now() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define FX_STAMP(i) do { if (tl && tid == 0) tl[(i)] = fx_now(); } while (0)

// ---- per-pixel noise stream of this engine --------------------------------------------------------
// Block c covers frames 9c+1 .. 9c+9 of pixel p with two Philox calls:
//   A: 4 words -> 8 normals (Box-Muller pairs), frames +1..+8
//   B: .x .y .z -> nine 10-bit fields (three per word, bits 0-9 / 10-19 / 20-29), first stage of
//      the shot test of frames +1..+9: a frame is a candidate only if field < 1024*bound, about
//      one pixel-frame in 5000 at the reference's rates; .w -> normal of frame +9
// A pack of T <= 10 frames is one block: two calls per pixel and pack.
// (The lock-step engines index their stage-0 stream by (4-pixel group, frame); a pixel that is
// resolved on its own would waste 3/4 of every group call.  The second-stage streams are shared.)
template <int R>
__device__ __forceinline__ uint4 noise_px(unsigned long long ctr64, unsigned p, unsigned c,
                                          unsigned which) {
    return philox4x32<R>(
        make_uint4(p, c | ((0x80u + which) << 8), (unsigned)ctr64, (unsigned)(ctr64 >> 32)));
}

// first stage of the shot test: frame +k+1 of the block owns the 10-bit field k of call B
__device__ __forceinline__ unsigned fx_shot_field(const uint4& wb, int k) {
    const unsigned w = k < 3 ? wb.x : (k < 6 ? wb.y : wb.z);
    return (w >> (10 * (k % 3))) & 1023u;
}
// 1024 * (upper bound of the probability that any iteration of the frame fires): shot_thr256
// (v2e_device.cuh) rescaled to the field width, exactly
__device__ __forceinline__ float fx_shot_thr(float k_side, float shot_m1, float I, float C) {
    return shot_thr256(k_side, shot_m1, I, C) * 4.0f;
}
// the frame's uniform u = (field + u2)/1024, u2 drawn only if the field alone cannot exclude
// u < bound (field >= 1024*bound  =>  u >= bound)
template <int R>
static __device__ __noinline__ float fx_shot_uniform(unsigned long long ctr64, int n, int p,
                                                     unsigned field) {
    const uint4 w = noise_words<R>(ctr64, (unsigned)n, (unsigned)p, 1);
    return ((float)field + (float)w.x * (1.0f / 4294967296.0f)) * (1.0f / 1024.0f);
}

// exact integer statistics of one voxel value: v = k * 2^-31.  Voxel values are sums of a few
// pol*(1-d), pol*d with d a fraction of an f32 time stamp in [0, num_bins): multiples of 2^-31
// whenever |v| >= 2^-8 (finer bits, if any, are truncated - the same way in every order).
// integer part and fraction are split exactly (Sterbenz), so no 64-bit conversion is needed.
__device__ __forceinline__ void fx_stat(float v, unsigned& nnz, long long& s1,
                                        unsigned long long& lo, unsigned long long& hi) {
    if (v != 0.f) {
        nnz += 1;
        const int ip = __float2int_rz(v);                         // |v| < 2^31 always (event counts)
        const int fp = __float2int_rz((v - (float)ip) * 2147483648.0f);
        const long long k = ((long long)ip << 31) + (long long)fp;
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


// leak decrement of one frame: (sdt*((leak*nra)*(1 - jit*z)))*pos   (emulator_utils.py:117-125)
__device__ __forceinline__ float leak_dec(float lk_nra, float jit, float z, float sdt, float pos) {
    return fmul(fmul(sdt, fmul(lk_nra, fsub(1.0f, fmul(jit, z)))), pos);
}

// per-CTA view of the shared-memory header
struct FxSmem {
    float2* lut;
    FrameSc* fr;
    FrameGlobals* g;
    unsigned* max;
    unsigned long long* acc;
    const unsigned short* perm;   // entries ordered by first active frame: cold list, then hot list
    long long* tl;   // development aid: timeline slots of this CTA for the current round (or null)
};

// tile partials -> shared memory (integer adds: order-free)
__device__ __forceinline__ void fx_flush(const FxAcc& acc, unsigned long long* s_acc) {
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
    if ((threadIdx.x & 31) == 0 && (nnz | nev)) {
        atomicAdd(&s_acc[0], nnz);
        atomicAdd(&s_acc[1], (unsigned long long)s1);
        const unsigned long long old = atomicAdd(&s_acc[2], lo);
        atomicAdd(&s_acc[3], hi + (old + lo < old ? 1ull : 0ull));
        atomicAdd(&s_acc[4], nev);
    }
}

// ---- event frame of one pixel under the current guess: the general case, out of line ---------------
// (shot-noise candidates, injected draws, gate off).  cnt/pol/C: diff_count results of frame n;
// bit: the frame may hold a shot event (I is read back from the frame only then).  Everything
// goes in and out by value (an address-taken local would live in local memory, and with the
// whole L1 carved out as shared memory that is an L2 round trip per access).
// Returns (final_evts_frame << 32) | bits(timestamp_mem).
template <int FDT, int R, bool GENERIC>
static __device__ __noinline__ unsigned long long fx_event_frame(
    const FxParams& prm, const SeqDev& sq, const FxSmem sm, int n, int p, int cnt, float pol,
    float C, bool bit, bool shot, bool philox, float tm, float* vp, int vstride) {
    const CfgDev& cfg = prm.cfg;
    const FrameSc& fs = sm.fr[n];
    const FrameGlobals g = sm.g[n];
    bool cand = false;
    float u = 0.f, I = 0.f;
    if (shot && bit && pol != 0.f) {
        float L;
        log_inten(sm.lut, load_frame1<FDT>(sq.frames, (long long)n * cfg.P + p), L, I);
        if (philox) {
            const int k = (n - 1) % kFxChunk;
            const uint4 wb = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)((n - 1) / kFxChunk), 1);
            const unsigned field = fx_shot_field(wb, k);
            const float thr = fx_shot_thr(pol > 0.f ? fs.kp : fs.kn, cfg.shot_inten_m1, I, C);
            if ((float)field < thr) {
                u = fx_shot_uniform<R>(sq.philox_ctr, n, p, field);
                cand = u * 1024.0f < thr;
            }
        } else {
            cand = true;   // injected draws: every pol != 0 pixel walks the iterations
        }
    }
    if (cnt > 0 || cand) {
        const int mode = (cand && g.mx > 0) ? (philox ? 2 : 1) : 0;
        return fx_events<R>(&prm.cfg, &sq, &fs, g.mx, g.ts_step, g.gate, n, p, pol, C,
                            min(cnt, g.mx), I, tm, mode, u, vp, vstride);
    }
    return (unsigned long long)__float_as_uint(tm);
}

template <bool GENERIC>
__device__ __forceinline__ void fx_dbg(const SeqDev& sq, int P, int n, int p, float pol, int cnt,
                                       int fin) {
    if (GENERIC) {
        const size_t o = (size_t)(n - 1) * P + p;
        if (sq.dbg_pol) sq.dbg_pol[o] = (signed char)pol;
        if (sq.dbg_cnt) sq.dbg_cnt[o] = cnt;
        if (sq.dbg_final) sq.dbg_final[o] = fin;
    }
}

template <bool GENERIC>
__device__ __forceinline__ void fx_dbg_clear_shot(const SeqDev& sq, int P, int p, int f, int T) {
    // an earlier pass (under another guess) may have marked other iterations
    if (GENERIC && sq.dbg_shot)
        for (int n = f; n < T; ++n)
            for (int i = 0; i < sq.dbg_shot_imax; ++i)
                sq.dbg_shot[((size_t)(n - 1) * sq.dbg_shot_imax + i) * P + p] = 0;
}

// ---- spilled pixels: inputs re-read from global memory, results kept in the workspace -------------
// Frame-synchronous, one pixel per lane, everything recomputed from the pack's first frame.
template <int FDT, int R, bool GENERIC, bool SPEC>
static __device__ __noinline__ void fx_resolve_spill(const FxParams& prm, const SeqDev& sq,
                                                     FxSmem sm, const unsigned* spill_list,
                                                     int n_spill, int tile_px0, float* spill_vox,
                                                     float* spill_state, bool leak, bool shot,
                                                     bool philox, bool cutoff) {
    const CfgDev& cfg = prm.cfg;
    const int T = SPEC ? kFxSpecT : sq.T, P = cfg.P, nb = SPEC ? kFxSpecNb : cfg.nb;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FxAcc acc;
    acc.nnz = 0; acc.nev = 0; acc.s1 = 0; acc.lo = 0; acc.hi = 0;
    for (int e0 = warp * 32; e0 < n_spill; e0 += (int)blockDim.x) {
        const int e = e0 + lane;
        const bool on = e < n_spill;
        const int p = tile_px0 + (on ? (int)spill_list[e] : 0);
        float base, lp = 0.f, tm;
        const float pos = __ldg(sq.pos + p), neg = __ldg(sq.neg + p);
        const float nlk = leak ? fmul(cfg.leak_rate, __ldg(sq.nra + p)) : 0.f;
        if (sq.first_pack) {
            float L0, I0;
            log_inten(sm.lut, load_frame1<FDT>(sq.frames, p), L0, I0);
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
        if (on) {
            for (int b = 0; b < nb; ++b) vp[(size_t)b * P] = 0.f;
            fx_dbg_clear_shot<GENERIC>(sq, P, p, 1, T);
        }
        const bool small = is_small_px(p, cfg.W);
        const int nchunks = (T - 1 + kFxChunk - 1) / kFxChunk;
        for (int c = 0; c < nchunks; ++c) {
            const int n0 = 1 + c * kFxChunk;
            uint4 wa = make_uint4(0, 0, 0, 0), wb = make_uint4(0, 0, 0, 0);
            if (philox && leak) {
                wa = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 0);
                if (T - n0 > 8) wb = noise_px<R>(sq.philox_ctr, (unsigned)p, (unsigned)c, 1);
            }
            float z0 = 0.f, z1 = 0.f;
#pragma unroll
            for (int k = 0; k < kFxChunk; ++k) {
                const int n = n0 + k;
                if (n >= T) break;
                const FrameSc& fs = sm.fr[n];
                if (leak && philox && (k & 1) == 0) box_muller16(k < 8 ? (&wa.x)[k >> 1] : wb.w, z0, z1);
                int cnt = 0;
                if (on) {
                    float L, I;
                    log_inten(sm.lut, load_frame1<FDT>(sq.frames, (long long)n * P + p), L, I);
                    float nw = L;
                    if (cutoff) {
                        lp = lowpass(lp, L, I, small ? fs.s_small : fs.s_large);
                        nw = lp;
                    }
                    if (leak) {
                        const float z = philox ? ((k & 1) ? z1 : z0)
                                               : __ldg(sq.leak_randn + (size_t)(n - 1) * P + p);
                        base = fsub(base, leak_dec(nlk, cfg.leak_jitter, z, fs.sdt, pos));
                        if (GENERIC && sq.dbg_leak_randn && philox)
                            sq.dbg_leak_randn[(size_t)(n - 1) * P + p] = z;
                    }
                    const PixelA a = diff_count(nw, base, pos, neg);
                    cnt = a.cnt;
                    if ((unsigned)cnt > sm.max[n]) atomicMax(&sm.max[n], (unsigned)cnt);
                    const unsigned long long r = fx_event_frame<FDT, R, GENERIC>(
                        prm, sq, sm, n, p, cnt, a.pol, a.C, true, shot, philox, tm, vp, P);
                    const int fin = (int)(r >> 32);
                    tm = __uint_as_float((unsigned)r);
                    if (fin) {
                        acc.nev += (unsigned)fin;
                        base = fadd(base, fmul(fmul(a.pol, (float)fin), a.C));   // v2e_model.py:522
                    }
                    fx_dbg<GENERIC>(sq, P, n, p, a.pol, cnt, fin);
                }
            }
        }
        if (on) {
            spill_state[p] = base;
            spill_state[(size_t)P + p] = lp;
            spill_state[(size_t)2 * P + p] = tm;
            for (int b = 0; b < nb; ++b) fx_stat(vp[(size_t)b * P], acc.nnz, acc.s1, acc.lo, acc.hi);
        }
    }
    fx_flush(acc, sm.acc);
}


__device__ __forceinline__ void fx_store_vox(const CfgDev& cfg, void* voxel, size_t o, float v) {
    if (cfg.voxel_bf16) ((__nv_bfloat16*)voxel)[o] = __float2bfloat16_rn(v);
    else ((float*)voxel)[o] = v;
}

// v2e_fixpoint_spec.cu
int launch_fixpoint_spec(const FxParams& prm, int frame_dtype, int noise, bool cutoff, int rounds,
                         int grid, size_t smem, int optin, cudaStream_t st);

}  // namespace v2e
