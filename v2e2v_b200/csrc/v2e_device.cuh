// v2e_device.cuh — device-side building blocks shared by the stepwise and persistent engines.
//
// Bit-exactness contract (DESIGN.md "Numerics"): every f32 operation the reference
// performs as a separate ATen op is one correctly-rounded instruction here (__fmul_rn /
// __fadd_rn / __fdiv_rn are never contracted into FMAs, independent of --fmad).  Code that only
// feeds the in-kernel RNG (PHILOX mode) is free to use FMAs and fast intrinsics.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/v2e2v_b200.h"

namespace v2e {

// ---- per-sequence device descriptor (passed by value inside the kernel parameter block) -------
struct SeqDev {
    const void* frames;
    float* base;
    float* lp;
    float* tmem;
    const float* pos;
    const float* neg;
    const float* nra;
    void* voxel;
    float* raw_events;
    long long raw_capacity;
    unsigned long long* num_events;
    int* status;
    int* max_iters;
    const float* leak_randn;
    const float* shot_rand;
    int shot_imax;
    int dbg_shot_imax;
    signed char* dbg_pol;
    int* dbg_cnt;
    int* dbg_final;
    float* dbg_leak_randn;
    unsigned char* dbg_shot;
    unsigned long long philox_ctr;  // splitmix64(seed) + pack offset: the (z,w) counter words
    // pack scalars
    float Tr, duration, bins_m1;
    int first_pack;
    unsigned shot_f64_mask;
    int T;
};

struct FrameSc {  // per-frame scalars, index n (32 bytes)
    float t_bin_prev, s_large, s_small, sdt;
    double shot_c;
    float kp, kn;  // first-stage shot factors 256*1.0001*f32(shot_c)*f32(1/nominal) (ON, OFF)
};

struct CfgDev {
    int H, W, P, nb;
    int cutoff_on, leak_on, shot_on, noise_mode;
    int output_mode, voxel_bf16;
    int frames_integer;
    float leak_rate, leak_jitter, rn_pos, rn_neg, shot_inten_m1;
    const float* lut;
};

// ---- exact helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// lin_log / rescale for an arbitrary (non-integer) frame value: v2e/emulator_utils.py:13-45.
// Out of line: integer-valued frames (every reader of the reference produces them) never come
// here.
static __device__ __noinline__ float2 log_inten_slow(float xf) {
    const double x = (double)xf;
    const double f = (1. / 20.) * 2.995732273553991;  // (1./threshold)*math.log(threshold)
    double y = (x <= 20.) ? x * f : log(x);
    y = rint(y * 1e8) / 1e8;
    return make_float2((float)y, fdiv(fadd(xf, 20.f), 275.f));
}

// floor(ad/den) for ad >= den (the ~4 % of pixel-frames that cross a threshold).  Out of line to
// keep the per-pixel fast path small.
static __device__ __noinline__ int count_div(float ad, float den) {
    return (int)floorf(fdiv(ad, den));
}

// LUT in shared memory: float2 {lin_log(v), (v+20)/275}.
__device__ __forceinline__ void load_lut(float2* s_lut, const float* __restrict__ g_lut) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
        s_lut[i] = make_float2(__ldg(g_lut + i), __ldg(g_lut + 256 + i));
}

// Frame value -> (L, I).  Integer-valued 0..255 -> LUT (exact: the LUT holds the reference
// functions' own outputs); anything else -> fp64 slow path.
__device__ __forceinline__ void log_inten(const float2* s_lut, float x, float& L, float& I) {
    const int v = __float2int_rn(x);
    float2 e;
    if (__builtin_expect((float)v == x && (unsigned)v < 256u, 1)) e = s_lut[v];
    else e = log_inten_slow(x);
    L = e.x;
    I = e.y;
}

// ---- frame loads --------------------------------------------------------------------------------
template <int FDT>
__device__ __forceinline__ float load_frame1(const void* __restrict__ f, long long idx) {
    if (FDT == V2E_DT_U8) return (float)__ldg((const unsigned char*)f + idx);
    if (FDT == V2E_DT_BF16)
        return __uint_as_float(((unsigned)__ldg((const unsigned short*)f + idx)) << 16);
    return __ldg((const float*)f + idx);
}

template <int FDT>
__device__ __forceinline__ float4 load_frame4(const void* __restrict__ f, long long idx) {
    // idx is a multiple of 4 and the plane is 16B (f32) / 4B (u8) / 8B (bf16) aligned
    if (FDT == V2E_DT_U8) {
        const uchar4 v = __ldg((const uchar4*)((const unsigned char*)f + idx));
        return make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
    }
    if (FDT == V2E_DT_BF16) {
        const uint2 v = __ldg((const uint2*)((const unsigned short*)f + idx));
        return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                           __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
    }
    return __ldg((const float4*)((const float*)f + idx));
}

// ---- Philox4x32 (Salmon et al., SC'11), R rounds ------------------------------------------------
// The key is a library constant (all per-emulator entropy sits in the 128-bit counter), so the
// key schedule folds into immediates: a round is 2 IMAD.WIDE + 2 three-input XORs.
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr) {
    unsigned k0 = 0xA4093822u, k1 = 0x299F31D0u;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        // (mul.wide.u32 spelled out: from the C expression (u64)a * M nvcc 12.9 also emits an add
        // of a zero uniform register to each high word - 20 instructions per call)
        unsigned long long p0, p1;
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(ctr.x), "r"(0xD2511F53u));
        asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(ctr.z), "r"(0xCD9E8D57u));
        ctr = make_uint4((unsigned)(p1 >> 32) ^ ctr.y ^ k0, (unsigned)p1,
                         (unsigned)(p0 >> 32) ^ ctr.w ^ k1, (unsigned)p0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return ctr;
}

// Noise words: a pure function of (counter base, frame n, index, stream), hence identical in both
// engines.  ctr64 = splitmix64(seed) + pack counter (host side, fill_seq).
//   stream 0, index = 4-pixel group G: .x/.y -> leak normals of pixels (0,1)/(2,3) of the group,
//                                      .z    -> one byte per pixel: first stage of the shot test
//   stream 1.., index = pixel p:       .x    -> second-stage / follow-up shot uniforms (rare)
template <int R>
__device__ __forceinline__ uint4 noise_words(unsigned long long ctr64, unsigned n, unsigned index,
                                             unsigned stream) {
    return philox4x32<R>(
        make_uint4(index, n | (stream << 8), (unsigned)ctr64, (unsigned)(ctr64 >> 32)));
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Two standard normals from one 32-bit word (Box-Muller): u1 = (lo16+1)/65536 in (0,1],
// u2 = hi16/65536.  r = sqrt(-2 ln u1) = sqrt(-2 ln2 * (log2(lo16+1) - 16)).
__device__ __forceinline__ void box_muller16(unsigned w, float& z0, float& z1) {
    const float a = (float)((w & 0xffffu) + 1u);
    const float r = sqrt_approx(fmaf(lg2_approx(a), -1.3862943611198906f, 22.18070977791825f));
    const float t = (float)(w >> 16) * (6.283185307179586f / 65536.0f);
    z0 = r * __cosf(t);
    z1 = r * __sinf(t);
}

// ---- shot noise in PHILOX mode: distribution-equivalent skip-ahead ------------------------------
// The reference draws one uniform per (pixel, iteration) and fires with probability p per
// iteration (emulator_utils.py:173-205).  P(any of ni iterations fires) = 1-(1-p)^ni <= ni*p
// = shot_c*fac*(thr/nominal), which does not depend on ni, so the "anything at all?" test can
// run before the per-frame max is known.  One uniform u decides: u >= bound -> nothing (exact);
// u < bound -> candidate, and the same u gives the first firing iteration by inverting the
// geometric law (exact); later firings use fresh uniforms.
// 256*bound for one pixel = (k_side * fac) * C with k_side = 256*1.0001*f32(shot_c)*rn_side (host,
// make_frame_sc; the 1e-4 relative slack covers rounding) and fac = (0.25-1)*inten01 + 1.
__device__ __forceinline__ float shot_thr256(float k_side, float shot_m1, float I, float C) {
    return (k_side * fmaf(shot_m1, I, 1.0f)) * C;
}
// The uniform is u = (byte + u2)/256 with u2 drawn (shot_uniform) only if the byte alone cannot
// exclude u < bound:  byte >= 256*bound  =>  u >= bound.  Candidate iff 256*u < 256*bound.

template <int R>
static __device__ __noinline__ float shot_uniform(unsigned long long ctr64, int n, int p,
                                                  unsigned byte) {
    const uint4 w = noise_words<R>(ctr64, (unsigned)n, (unsigned)p, 1);
    return ((float)byte + (float)w.x * (1.0f / 4294967296.0f)) * (1.0f / 256.0f);
}

// Geometric skip: index of the first success of iid Bernoulli(p) trials, from one uniform:
// floor(log1p(-u)/log1p(-p)).  Both arguments are tiny in practice (u < bound ~ 5e-4), where the
// series -log1p(-x) = x + x^2/2 + x^3/3 + x^4/4 (relative error < x^4) is exact to f32 rounding.
// the small-argument branch on its own (u < 0.03, 0 < p < 0.03), for callers that stay inline
__device__ __forceinline__ float geometric_skip_small(float u, float p) {
    const float lu = u * fmaf(u, fmaf(u, fmaf(u, 0.25f, 0.33333334f), 0.5f), 1.0f);
    const float lq = p * fmaf(p, fmaf(p, fmaf(p, 0.25f, 0.33333334f), 0.5f), 1.0f);
    return floorf(__fdividef(lu, lq));
}
static __device__ __noinline__ int geometric_skip(float u, float p) {
    if (!(p > 0.f)) return 0x7fffffff;
    if (p >= 1.f) return 0;
    float g;
    if (u < 0.03f && p < 0.03f) {
        g = geometric_skip_small(u, p);
    } else {
        g = (float)floor(log1p(-(double)u) / log1p(-(double)p));
    }
    return g > 2.0e9f ? 0x7fffffff : (int)g;
}

// ---- sensing-diversity pattern ------------------------------------------------------------------
// pixels at even row AND even column use (ps, qs): v2e_model.py:194,211; emulator_utils.py:87-89
__device__ __forceinline__ bool is_small_px(int p, int W) {
    const int y = p / W, x = p - y * W;
    return ((y | x) & 1) == 0;
}

// ---- per-pixel phase A: low-pass, leak, diff, polarity, count ------------------------------------
// lp update (emulator_utils.py:75-93).  s = +inf encodes eps = 1 (q <= 0 branch).
__device__ __forceinline__ float lowpass(float lp, float L, float I, float s) {
    float eps = fmul(I, s);           // I > 0, s = +inf -> +inf -> clamped to 1
    eps = fminf(eps, 1.0f);           // torch.clamp(max=1)
    return fadd(fmul(fsub(1.0f, eps), lp), fmul(eps, L));
}

// base - (sdt*((leak*nra)*(1 - jit*z)))*pos   (emulator_utils.py:117-125)
__device__ __forceinline__ float leak_sub(float base, float lk_nra, float jit, float z, float sdt,
                                          float pos) {
    const float clr = fmul(lk_nra, fsub(1.0f, fmul(jit, z)));
    return fsub(base, fmul(fmul(sdt, clr), pos));
}

// diff / polarity / threshold / count (v2e_model.py:375,394-404).
//   pol in {-1,0,+1}; C = threshold of the polarity side (0 if pol == 0); cnt = event_counts.
// floor(|diff|/(C+1e-9)) == 0  <=>  |diff| < C+1e-9 (exact: the quotient of two floats a < b is
// at most 1-2^-24, which is representable), so the division is only evaluated on a crossing.
struct PixelA {
    float pol, C;
    int cnt;
};
__device__ __forceinline__ PixelA diff_count(float nw, float base, float pos, float neg) {
    PixelA r;
    const float diff = fsub(nw, base);
    const float ad = fabsf(diff);
    const bool live = ad > 1e-6f;  // tolerance mask (:394-395); NaN -> not live, like masked_fill
    const bool up = diff > 0.f;
    r.pol = live ? (up ? 1.f : -1.f) : 0.f;
    r.C = live ? (up ? pos : neg) : 0.f;
    const float den = fadd(r.C, 1e-9f);
    // one code path for every lane: the quotient is always formed (the division is ~10
    // instructions; a divergent "only on a crossing" branch costs a warp more than that)
    const int q = (int)floorf(fdiv(ad, den));
    r.cnt = (live && ad >= den) ? q : 0;
    return r;
}

// ---- per-frame globals (v2e_model.py:405-417,469) -----------------------------------------------
struct FrameGlobals {
    int mx;
    float ts_step;
    int gate;
};

__device__ __forceinline__ FrameGlobals frame_globals(int mx, float duration, float Tr) {
    FrameGlobals g;
    g.mx = mx;
    const int ni = mx > 0 ? mx : 1;
    g.ts_step = fmul(__frcp_rn((float)ni), duration);  // reciprocal()*duration (trap T6)
    g.gate = Tr > g.ts_step;
    return g;
}

// ---- per-pixel phase B: the iteration loop (v2e_model.py:449-522) --------------------------------
// VoxAdd: void(int bin, float val); Emit: void(float ts, float pol).
// Returns final_evts_frame for this pixel (post-refractory event count incl. shot events).
//   pol/C/cnt: phase-A results; tm: timestamp_mem (updated in place)
//   shot_mode: 0 none, 1 injected (reads sq.shot_rand), 2 philox candidate with uniform u
template <int R, class VoxAdd, class Emit>
__device__ __forceinline__ int iterate_pixel(const CfgDev& cfg, const SeqDev& sq, const FrameSc& fs,
                                             const FrameGlobals& g, int n, int p, float pol,
                                             float C, int cnt, float I, float& tm, int shot_mode,
                                             float u, VoxAdd vox_add, Emit emit) {
    int final_cnt = 0;
    if (shot_mode == 0 && g.gate) {
        // No shot noise at this pixel and the gate is on (v2e_model.py:469): iterations [0,cnt)
        // all carry an event and the gate refuses most of them.  Iteration j is accepted iff
        //     fsub(fadd(t0, fmul(step, j+1)), tm) > Tr                       (:419, :469-470)
        // and every operation is monotone non-decreasing in j, so the accepted iterations after
        // the last accepted one form a suffix: jump to its first element from an estimate and
        // fix it up with the exact predicate (evaluated as the reference does).  One loop trip
        // per ACCEPTED event, the same code for every lane (no divergent variants).
        const float t0 = fs.t_bin_prev, step = g.ts_step, Tr = sq.Tr;
        const float rstep = __fdividef(1.0f, step);
        auto pred = [&](int j) { return fsub(fadd(t0, fmul(step, (float)(j + 1))), tm) > Tr; };
        int i = 0;
        while (i < cnt) {
            // (j+1)*step + t0 - tm > Tr  <=>  j+1 > (Tr + tm - t0)/step   (estimate only)
            int j = (int)floorf((Tr + tm - t0) * rstep) - 1;
            j = min(max(j, i), cnt);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (j < cnt && !pred(j)) ++j;
            while (j < cnt && !pred(j)) ++j;   // estimate too low by more than 3 (never seen)
            while (j > i && pred(j - 1)) --j;  // estimate too high
            if (j >= cnt) break;
            const float ts = fadd(t0, fmul(step, (float)(j + 1)));
            tm = ts;
            final_cnt += 1;
            i = j + 1;
            if (cfg.output_mode == V2E_OUT_RAW) {
                emit(ts, pol);
            } else {  // v2e_model.py:478-502
                const float ti = floorf(ts);
                const float d = fsub(ts, ti);
                const int b = (int)ti;
                if (b >= 0 && b < cfg.nb) vox_add(b, fmul(pol, fsub(1.0f, d)));
                if (b >= 0 && b + 1 < cfg.nb) vox_add(b + 1, fmul(pol, d));
            }
        }
        return final_cnt;
    }
    int last = cnt;  // iterations [0,last) can produce events
    double thr64 = 0.;
    float thr32 = 0.f, p_it = 0.f;
    int next_shot = 0x7fffffff;
    unsigned follow = 2;
    bool f64 = false;
    const bool up = pol > 0.f;
    if (shot_mode && g.mx > 0 && pol != 0.f) {
        // emulator_utils.py:173-188: factor = ((rate/2)*dt/ni) * ((0.25-1)*inten01 + 1);
        // ON fires when rand > 1 - factor*ppp, OFF when rand < factor*npp (f32 or f64, trap T5)
        const float fac = fadd(fmul(cfg.shot_inten_m1, I), 1.0f);
        const float pp = fmul(up ? cfg.rn_pos : cfg.rn_neg, C);
        const int ni = g.mx;  // > 0 here, so num_iters == mx
        f64 = (sq.shot_f64_mask >> n) & 1u;
        if (f64) {
            // (rate/2)*dt / num_iters in f64 (emulator_utils.py:174, trap T5)
            const double a64 = fs.shot_c / (double)ni;
            const double q = __dmul_rn(__dmul_rn(a64, (double)fac), (double)pp);
            thr64 = up ? __dsub_rn(1.0, q) : q;
            p_it = (float)q;
        } else {
            const float a32 = fdiv((float)fs.shot_c, (float)ni);
            const float q = fmul(fmul(a32, fac), pp);
            thr32 = up ? fsub(1.0f, q) : q;
            p_it = q;
        }
        if (shot_mode == 1) {
            last = g.mx;
        } else {
            next_shot = geometric_skip(u, p_it);
            if (next_shot < g.mx) last = g.mx;  // walk to mx (rare path)
        }
    }
    if (last > g.mx) last = g.mx;
    // PHILOX candidates only visit the iterations that can fire: [0,cnt) and the shot indices
    int i = (shot_mode == 2 && cnt == 0) ? next_shot : 0;
    for (; i < last;) {
        const int cur = i;
        bool m = cnt >= i + 1;
        if (shot_mode == 1 && pol != 0.f) {
            const float r =
                __ldg(sq.shot_rand + ((long long)(n - 1) * sq.shot_imax + i) * cfg.P + p);
            bool s;
            if (f64) s = up ? ((double)r > thr64) : ((double)r < thr64);
            else s = up ? (r > thr32) : (r < thr32);
            m = m || s;
        } else if (shot_mode == 2 && i == next_shot) {
            if (sq.dbg_shot && i < sq.dbg_shot_imax)
                sq.dbg_shot[((long long)(n - 1) * sq.dbg_shot_imax + i) * cfg.P + p] = 1;
            m = true;
            const uint4 w = noise_words<R>(sq.philox_ctr, (unsigned)n,
                                           (unsigned)p, follow);
            follow = follow < 0x7fu ? follow + 1 : follow;
            const int gap = geometric_skip((float)(w.x >> 8) * (1.0f / 16777216.0f), p_it);
            next_shot = gap >= 0x7fffffff - i - 1 ? 0x7fffffff : i + 1 + gap;
        }
        i = cur + 1;
        if (shot_mode == 2 && i >= cnt && i < next_shot) i = next_shot;
        const float ts = fadd(fs.t_bin_prev, fmul(g.ts_step, (float)(cur + 1)));
        if (g.gate) {  // v2e_model.py:469-472
            const float since = fsub(m ? ts : 0.f, tm);
            m = since > sq.Tr;
            if (m) tm = ts;
        }
        if (!m) continue;
        final_cnt += 1;
        if (cfg.output_mode == V2E_OUT_RAW) {
            emit(ts, pol);
        } else {  // v2e_model.py:478-502
            const float ti = floorf(ts);
            const float d = fsub(ts, ti);
            const int b = (int)ti;
            if (b >= 0 && b < cfg.nb) vox_add(b, fmul(pol, fsub(1.0f, d)));
            if (b >= 0 && b + 1 < cfg.nb) vox_add(b + 1, fmul(pol, d));
        }
    }
    return final_cnt;
}

}  // namespace v2e
