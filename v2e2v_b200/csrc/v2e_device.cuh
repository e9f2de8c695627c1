// v2e_device.cuh — device-side building blocks shared by the stepwise and persistent engines.
//
// Bit-exactness contract (DESIGN.md "Numerics"): every f32 operation the reference
// performs as a separate ATen op is one correctly-rounded instruction here (__fmul_rn /
// __fadd_rn / __fdiv_rn are never contracted into FMAs, independent of --fmad).
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
    unsigned long long philox_seed;
    unsigned long long philox_offset;
    // pack scalars
    float Tr, duration, bins_m1;
    int first_pack;
    unsigned shot_f64_mask;
    int T;
};

struct FrameSc {  // per-frame scalars, index n
    float t_bin_prev, s_large, s_small, sdt;
    double shot_c;
};

struct CfgDev {
    int H, W, P, nb;
    int cutoff_on, leak_on, shot_on, noise_mode;
    int output_mode, voxel_bf16;
    float leak_rate, leak_jitter, rn_pos, rn_neg, shot_inten_m1;
    const float* lut;
};

// ---- exact helpers ----------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }

// lin_log for an arbitrary (non-integer) frame value: v2e/emulator_utils.py:13-37 in fp64.
static __device__ __noinline__ float lin_log_slow(float xf) {
    const double x = (double)xf;
    const double f = (1. / 20.) * 2.995732273553991;  // (1./threshold)*math.log(threshold)
    double y = (x <= 20.) ? x * f : log(x);
    y = rint(y * 1e8) / 1e8;
    return (float)y;
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
    if (__builtin_expect((float)v == x && (unsigned)v < 256u, 1)) {
        const float2 e = s_lut[v];
        L = e.x;
        I = e.y;
    } else {
        L = lin_log_slow(x);
        I = fdiv(fadd(x, 20.f), 275.f);
    }
}

// ---- frame loads (4 consecutive pixels) -------------------------------------------------------
template <int FDT>
__device__ __forceinline__ float load_frame1(const void* __restrict__ f, long long idx) {
    if (FDT == V2E_DT_U8) return (float)__ldg((const unsigned char*)f + idx);
    if (FDT == V2E_DT_BF16) return __bfloat162float(((const __nv_bfloat16*)f)[idx]);
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
        const uint2 v = __ldg((const uint2*)((const __nv_bfloat16*)f + idx));
        return make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                           __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
    }
    return __ldg((const float4*)((const float*)f + idx));
}

// ---- Philox4x32 (Salmon et al., SC'11), R rounds ------------------------------------------------
template <int R>
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}

// Noise for the 4-pixel group G of frame n (engine-independent: a pure function of
// (seed, pack offset, n, G)).  z[4]: standard normals (Box-Muller on 16+16 bit uniforms),
// u[4]: 32-bit uniforms as raw words.  `stream` selects independent draws (0: per-frame draw,
// >= 1: follow-up shot draws of the rare path).
template <int R>
__device__ __forceinline__ uint4 noise_words(unsigned long long seed, unsigned long long offset,
                                             unsigned n, unsigned G, unsigned stream) {
    const uint4 ctr = make_uint4(G, n | (stream << 8), (unsigned)offset, (unsigned)(offset >> 32));
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    return philox4x32<R>(ctr, key);
}

// Two standard normals from one 32-bit word: u1 = (lo16+1)/65536 in (0,1], u2 = hi16/65536.
__device__ __forceinline__ void box_muller16(unsigned w, float& z0, float& z1) {
    const float u1 = (float)((w & 0xffffu) + 1u) * (1.0f / 65536.0f);
    const float u2 = (float)(w >> 16) * (1.0f / 65536.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    z0 = r * c;
    z1 = r * s;
}

__device__ __forceinline__ float u32_to_unit(unsigned w) {  // [0,1) with 24 significant bits
    return (float)(w >> 8) * (1.0f / 16777216.0f);
}

// ---- sensing-diversity pattern ------------------------------------------------------------------
// pixels at even row AND even column use (ps, qs): v2e_model.py:194,211; emulator_utils.py:87-89
__device__ __forceinline__ bool is_small_px(int p, int W) {
    const int y = p / W, x = p - y * W;
    return ((y | x) & 1) == 0;
}

// ---- per-pixel phase A: low-pass, leak, diff, polarity, count ------------------------------------
struct PixelA {
    float diff;  // tolerance-masked diff (v2e_model.py:394-395)
    int cnt;     // event_counts (v2e_model.py:404)
};

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

__device__ __forceinline__ PixelA diff_count(float nw, float base, float pos, float neg) {
    PixelA r;
    float diff = fsub(nw, base);
    if (!(fabsf(diff) > 1e-6f)) diff = 0.f;
    r.diff = diff;
    const float C = diff > 0.f ? pos : (diff < 0.f ? neg : 0.f);
    const float den = fadd(C, 1e-9f);
    const float ad = fabsf(diff);
    // floor(ad/den) == 0  <=>  ad < den (exact), so divide only when needed
    r.cnt = (ad >= den) ? (int)floorf(fdiv(ad, den)) : 0;
    return r;
}

// ---- per-pixel phase B: the iteration loop (v2e_model.py:449-522) --------------------------------
struct FrameGlobals {  // derived from the per-frame global max (v2e_model.py:405-417,469)
    int mx;
    float ts_step;
    bool gate;
    double a64;
    float a32;
    bool f64;
};

__device__ __forceinline__ FrameGlobals frame_globals(int mx, float duration, float Tr,
                                                      double shot_c, bool f64) {
    FrameGlobals g;
    g.mx = mx;
    const int ni = mx > 0 ? mx : 1;
    g.ts_step = fmul(__frcp_rn((float)ni), duration);  // reciprocal()*duration (trap T6)
    g.gate = Tr > g.ts_step;
    g.f64 = f64;
    g.a64 = shot_c / (double)ni;
    g.a32 = fdiv((float)shot_c, (float)ni);
    return g;
}


// Shot-noise thresholds of one pixel (emulator_utils.py:173-188), f32 or f64 per trap T5.
struct ShotThr {
    double on64, off64;
    float on32, off32;
    float p_on, p_off;  // per-iteration firing probabilities (PHILOX mode)
};

__device__ __forceinline__ ShotThr shot_thresholds(const CfgDev& cfg, const FrameGlobals& g,
                                                   float I, float pos, float neg) {
    ShotThr t;
    const float fac = fadd(fmul(cfg.shot_inten_m1, I), 1.0f);
    const float ppp = fmul(cfg.rn_pos, pos);
    const float npp = fmul(cfg.rn_neg, neg);
    if (g.f64) {
        const double snf = __dmul_rn(g.a64, (double)fac);
        t.on64 = __dsub_rn(1.0, __dmul_rn(snf, (double)ppp));
        t.off64 = __dmul_rn(snf, (double)npp);
        t.on32 = 2.f; t.off32 = -1.f;
        t.p_on = (float)(snf * (double)ppp);
        t.p_off = (float)t.off64;
    } else {
        const float snf = fmul(g.a32, fac);
        t.on32 = fsub(1.0f, fmul(snf, ppp));
        t.off32 = fmul(snf, npp);
        t.on64 = 2.; t.off64 = -1.;
        t.p_on = fmul(snf, ppp);
        t.p_off = t.off32;
    }
    return t;
}

// Conservative bound on P(any shot event in this pixel-frame) = 1-(1-p)^ni <= ni*p
// = shot_c*fac*ppp: independent of ni, so the candidate test runs before the global max is known.
__device__ __forceinline__ float shot_any_bound(const CfgDev& cfg, float shot_c32, float I,
                                                float thr, float rn) {
    const float fac = fmaf(cfg.shot_inten_m1, I, 1.0f);
    return shot_c32 * fac * (rn * thr) * 1.0001f;  // +1e-4 relative slack over rounding
}

// Geometric skip: index of the first success of iid Bernoulli(p) trials, from one uniform.
__device__ __forceinline__ int geometric_skip(float u, float p) {
    if (!(p > 0.f)) return 0x7fffffff;
    if (p >= 1.f) return 0;
    const double g = floor(log1p(-(double)u) / log1p(-(double)p));
    return g > 2.0e9 ? 0x7fffffff : (int)g;
}

// The per-pixel iteration loop.  VoxAdd: void(int bin, float val); Emit: void(float ts, float pol).
// Returns final_evts_frame for this pixel (post-refractory event count incl. shot events).
//   cnt/diff: phase-A results; tm: timestamp_mem (updated in place)
//   shot_mode: 0 none, 1 injected (reads sq.shot_rand), 2 philox (first_shot = first shot
//   iteration index or INT_MAX; follow-ups drawn from the `stream >= 1` words)
template <int R, class VoxAdd, class Emit>
__device__ __forceinline__ int iterate_pixel(const CfgDev& cfg, const SeqDev& sq, const FrameSc& fs,
                                             const FrameGlobals& g, int n, int p, float diff,
                                             int cnt, float I, float pos, float neg, float& tm,
                                             int shot_mode, unsigned shot_word, VoxAdd vox_add,
                                             Emit emit) {
    const float pol = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    int final_cnt = 0;
    int last = cnt;  // iterations [0,last) can produce events
    ShotThr thr;
    int next_shot = 0x7fffffff;
    unsigned follow = 1;
    float p_it = 0.f;
    if (shot_mode && g.mx > 0 && pol != 0.f) {
        thr = shot_thresholds(cfg, g, I, pos, neg);
        if (shot_mode == 1) {
            last = g.mx;
        } else {
            p_it = pol > 0.f ? thr.p_on : thr.p_off;
            next_shot = geometric_skip(u32_to_unit(shot_word), p_it);
            if (next_shot < g.mx) last = g.mx;  // conservatively walk to mx (rare path)
        }
    } else if (shot_mode == 1 && g.mx > 0) {
        last = cnt;  // pol == 0: shot events are masked out (v2e_model.py:447)
    }
    if (last > g.mx) last = g.mx;
    for (int i = 0; i < last; ++i) {
        bool m = cnt >= i + 1;
        if (shot_mode == 1 && pol != 0.f) {
            const float r = __ldg(sq.shot_rand + ((long long)(n - 1) * sq.shot_imax + i) * cfg.P + p);
            bool on, off;
            if (g.f64) { on = (double)r > thr.on64; off = (double)r < thr.off64; }
            else { on = r > thr.on32; off = r < thr.off32; }
            const bool s = (on && pol > 0.f) || (off && pol < 0.f);
            if (sq.dbg_shot && i < sq.dbg_shot_imax)
                sq.dbg_shot[((long long)(n - 1) * sq.dbg_shot_imax + i) * cfg.P + p] = s;
            m = m || s;
        } else if (shot_mode == 2 && i == next_shot) {
            if (sq.dbg_shot && i < sq.dbg_shot_imax)
                sq.dbg_shot[((long long)(n - 1) * sq.dbg_shot_imax + i) * cfg.P + p] = 1;
            m = true;
            const uint4 w = noise_words<R>(sq.philox_seed, sq.philox_offset, (unsigned)n,
                                           (unsigned)p, follow++);
            const int gap = geometric_skip(u32_to_unit(w.x), p_it);
            next_shot = gap >= 0x7fffffff - i - 1 ? 0x7fffffff : i + 1 + gap;
        }
        const float ts = fadd(fs.t_bin_prev, fmul(g.ts_step, (float)(i + 1)));
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
