// v2e_runahead.cu — the "run-ahead" engine: ONE cooperative launch per pack, streaming decoupled
// from the per-frame global maximum.
//
// Why.  The only cross-pixel coupling of the emulator is the image-wide max event count of a
// frame (v2e/v2e_model.py:405-406): it sets the event time stamps of that frame (:416-422) and
// through the refractory gate (:469-472) the number of events a pixel really emits, hence its
// memorised brightness (:522).  So frame n+1 of a pixel THAT FIRED in frame n cannot be evaluated
// before a grid-wide reduction of frame n has completed: T-1 dependent grid barriers per pack.
// But ~97 % of all pixel-frames fire nothing, and for those nothing depends on the maximum.
//
// How.  Each CTA owns a contiguous tile of pixels of one sequence and splits its warps in two
// roles that run concurrently:
//   * STREAM warps (26): per-pixel state in registers, time-step over the frames of the pack at
//     memory speed, never wait for anybody.  A pixel that crosses a threshold (or is a shot-noise
//     candidate) at frame n is PARKED: its state at frame n goes to its state planes, its park
//     frame to a byte map in shared memory, its count to the tile's running max of frame n; the
//     streaming registers of that pixel are poisoned (NaN thresholds: it never fires again) and
//     streaming goes on.
//   * RESOLVE warps (6): for n = 1..T-1: wait until the tile has streamed frame n, publish the
//     tile max of frame n and wait for the other tiles (one 64-bit word per frame: arrivals |
//     max), then take the pixels parked at n, one per thread: event iteration with refractory
//     gate, voxel accumulation (shared memory), memory update, and catch-up through the frames
//     n+1.. (fetched ahead of time, while the barrier is in flight) until the pixel crosses
//     again (re-park) or the pack ends.
// The dependent chain of T-1 grid barriers thus carries only the few parked pixels, and runs
// concurrently with the streaming of later frames.  The epilogue (all warps) reduces the
// statistics of the un-normalised grid over all tiles (utils/event_process.py:166-174), then
// normalises and stores the voxel tile.  CTAs of different sequences never wait for each other.
// Results are bit-identical to the stepwise engine (same per-pixel arithmetic, same noise words).
#include <cstdlib>

#include "v2e_engine.cuh"

namespace v2e {

#ifndef V2E_RA_L2PREFETCH
#define V2E_RA_L2PREFETCH 0
#endif
constexpr int kRaThreads = 1024;
constexpr int kRaStreamWarps = 25;
constexpr int kRaResolveWarps = kRaThreads / 32 - kRaStreamWarps;  // 7
constexpr int kRaStreamThreads = kRaStreamWarps * 32;              // 800
constexpr int kRaResolveThreads = kRaResolveWarps * 32;            // 224
constexpr int kRaPPT = 8;                                          // pixels per stream thread
constexpr int kRaGroups = kRaPPT / 4;
constexpr int kRaTileMaxPx = kRaStreamThreads * kRaPPT;            // 6400
constexpr int kRaWarps = kRaThreads / 32;
// threshold / leak planes in shared memory: [slot e of a group][group], fixed stride so that the
// slot offset is an immediate in the load instruction
constexpr int kRaSlotStride = kRaStreamThreads * kRaGroups;  // groups per tile (capacity)
constexpr int kRaParkDone = 0xff;  // park byte of a pixel the resolve warps carried to the end
static_assert(V2E_MAX_FRAMES < kRaParkDone, "park byte holds the frame index");

// control block in shared memory
struct RaCtl {
    int stream_cnt[V2E_MAX_FRAMES];  // stream warps that have finished frame n
    int lmax[V2E_MAX_FRAMES];        // running max event count of this tile in frame n
    int mx;                          // released image-wide max of the current round
    int listn[2];                    // entries in the park list of round n: listn[n & 1]
    int nonint;                      // a frame value of this tile is not an integer in 0..255
    int last;                        // this CTA is the last one of its sequence to leave
    int handover;                    // the stream warps are done: the next round starts phase 2
    int phase2;                      // first round of phase 2 (T: none)
    float stat[4];                   // mean, den, normalise?
    double dred[kRaWarps * 4];
};
constexpr int kRaLutOff = 0;                                  // float2[256]
constexpr int kRaFrOff = 2048;                                // FrameSc[V2E_MAX_FRAMES]
constexpr int kRaCtlOff = kRaFrOff + V2E_MAX_FRAMES * 32;     // RaCtl
constexpr int kRaHeaderBytes = 5120;
static_assert(sizeof(FrameSc) == 32, "FrameSc layout");
static_assert(kRaCtlOff + sizeof(RaCtl) <= kRaHeaderBytes, "header too small");

struct RaParams {
    CfgDev cfg;
    int n_seq;
    int tiles_per_seq;
    int tile_groups;  // 4-pixel groups per tile
    int use_bytes;    // the frame values of the whole pack are kept as bytes in shared memory
    SeqWork* work;
    long long* timeline;  // development aid (V2E_TIMELINE=1): [cta][frame][8] clock64 stamps
    SeqDev seq[V2E_MAX_SEQ_PER_LAUNCH];
    FrameSc fr[V2E_MAX_SEQ_PER_LAUNCH][V2E_MAX_FRAMES];
};

// ---- small PTX helpers --------------------------------------------------------------------------
__device__ __forceinline__ unsigned ra_smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ unsigned ra_ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ra_st_relaxed(unsigned* p, unsigned v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ra_ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ra_st_relaxed64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// state words exchanged between the two roles of one CTA through global memory: strong (L2)
// accesses ordered by fence.cta + the shared-memory park byte / counters
__device__ __forceinline__ float ra_ldg_strong(const float* p) {
    float v;
    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ra_ldg_strong4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void ra_stg_strong(float* p, float v) {
    asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void ra_red_add(float* p, float v) {
    asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ int ra_lds_volatile(const int* p) {
    int v;
    asm volatile("ld.volatile.shared.s32 %0, [%1];" : "=r"(v) : "r"(ra_smem_u32(p)) : "memory");
    return v;
}
// release/acquire fence at CTA scope (__threadfence_block() is the much heavier fence.sc.cta)
__device__ __forceinline__ void ra_fence_cta() {
    asm volatile("fence.acq_rel.cta;" ::: "memory");
}
__device__ __forceinline__ void ra_named_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ float4 ra_lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float ra_lds_f1(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 ra_lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}

// Frame value -> LUT index.  Integer-valued 0..255 is the fast path (every reader of the
// reference produces such frames); `ok` is false otherwise.  Uses only full-rate FADD/IADD
// (no F2I/I2F): x + 1.5*2^23 puts round(x) into the low mantissa bits.
__device__ __forceinline__ int ra_index(float x, bool& ok) {
    const float y = __fadd_rn(x, 12582912.0f);
    const int v = __float_as_int(y) - 0x4B400000;
    ok = (__fsub_rn(y, 12582912.0f) == x) && ((unsigned)v < 256u);
    return v;
}
// exact float of a byte without I2F
__device__ __forceinline__ float ra_byte_to_float(unsigned b) {
    return __fsub_rn(__uint_as_float(0x4B000000u | b), 8388608.0f);
}

// four consecutive frame pixels as they sit in memory
template <int FDT> struct RaVec { float4 v; };
template <> struct RaVec<V2E_DT_U8> { unsigned v; };
template <> struct RaVec<V2E_DT_BF16> { uint2 v; };
template <int FDT>
__device__ __forceinline__ RaVec<FDT> ra_load4(const void* __restrict__ f, long long idx) {
    RaVec<FDT> r;
    if constexpr (FDT == V2E_DT_U8) r.v = __ldg((const unsigned*)((const unsigned char*)f + idx));
    else if constexpr (FDT == V2E_DT_BF16) r.v = __ldg((const uint2*)((const unsigned short*)f + idx));
    else r.v = __ldg((const float4*)((const float*)f + idx));
    return r;
}
// element e of the vector: LUT index, validity, and the value itself (for the slow path)
template <int FDT>
__device__ __forceinline__ int ra_elem(const RaVec<FDT>& x, int e, bool& ok, float& xf) {
    if constexpr (FDT == V2E_DT_U8) {
        const unsigned v = (x.v >> (8 * e)) & 0xffu;
        ok = true;
        xf = 0.f;
        return (int)v;
    } else if constexpr (FDT == V2E_DT_BF16) {
        const unsigned w = e < 2 ? x.v.x : x.v.y;
        xf = __uint_as_float((e & 1) ? (w & 0xffff0000u) : (w << 16));
        return ra_index(xf, ok);
    } else {
        xf = e == 0 ? x.v.x : (e == 1 ? x.v.y : (e == 2 ? x.v.z : x.v.w));
        return ra_index(xf, ok);
    }
}

// second stage of the PHILOX shot-candidate test, shared by both roles (device.cuh, "shot noise
// in PHILOX mode"): candidate iff u*256 < thr256 with u = (byte + u2)/256
template <int R>
__device__ __forceinline__ bool ra_shot_candidate(unsigned long long ctr, int n, int p,
                                                  unsigned byte, float thr256, float& u) {
    if (!(ra_byte_to_float(byte) < thr256)) return false;
    u = shot_uniform<R>(ctr, n, p, byte);
    return u * 256.0f < thr256;
}

// a parked pixel in the registers of a resolve thread
struct RaEntry {
    int lidx;    // tile-local pixel index; < 0: empty
    bool small;  // a (ps, qs) pixel: even row and even column
    float b, l, tm;
};

// NOISE: 0 = no noise, 1 = PHILOX leak + shot (the V2E2VNet wiring), 2 = generic: every feature
// flag is read at run time, debug outputs and injected noise are honoured (tests, odd configs).
template <int FDT, int R, int NOISE, bool CUTOFF, bool INTF>
__global__ void __launch_bounds__(kRaThreads, 1) k_runahead(const __grid_constant__ RaParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool GENERIC = NOISE == 2;
    const CfgDev& cfg = prm.cfg;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    const int tiles = prm.tiles_per_seq;
    const SeqDev& sq = prm.seq[seq];
    SeqWork& wk = prm.work[seq];
    const int P = cfg.P;
    const int nb = cfg.nb;
    const int T = sq.T;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    const bool leak = GENERIC ? (cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool shot = GENERIC ? (cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool philox = GENERIC ? (cfg.noise_mode == V2E_NOISE_PHILOX) : (NOISE == 1);
    const bool cutoff = GENERIC ? (cfg.cutoff_on != 0) : CUTOFF;

    // ---- tile geometry ------------------------------------------------------------------------
    const int total_groups = P >> 2;  // P % 4 == 0 is a precondition of this engine
    const int g0 = tile * prm.tile_groups;
    int ngroups = total_groups - g0;
    ngroups = ngroups > prm.tile_groups ? prm.tile_groups : (ngroups < 0 ? 0 : ngroups);
    const int tile_px0 = g0 * 4;
    const int tile_px = ngroups * 4;
    const int tpa = prm.tile_groups * 4;  // allocated pixels per tile (multiple of 16)

    // ---- shared memory carve-up -----------------------------------------------------------------
    float2* s_lut = (float2*)(smem_raw + kRaLutOff);
    FrameSc* s_fr = (FrameSc*)(smem_raw + kRaFrOff);
    RaCtl* ctl = (RaCtl*)(smem_raw + kRaCtlOff);
    float2* s_pn = (float2*)(smem_raw + kRaHeaderBytes);                 // [4][kRaSlotStride] {pos, neg}
    float* s_lkn = (float*)(s_pn + 4 * kRaSlotStride);                   // [4][kRaSlotStride] if leak
    float* s_base = s_lkn + (leak ? 4 * kRaSlotStride : 0);              // [tpa] state of parked pixels
    float* s_lp = s_base + tpa;                                          // [tpa] if cutoff
    unsigned char* s_park = (unsigned char*)(s_lp + (cutoff ? tpa : 0)); // [tpa]
    unsigned short* s_list = (unsigned short*)(s_park + tpa);            // [tpa]
    unsigned char* s_byte = (unsigned char*)(s_list + tpa);              // [T-1][tpa] if use_bytes
    const bool use_bytes = prm.use_bytes != 0;
    float* const g_vox = (float*)sq.voxel;  // un-normalised accumulators live in the output grid
    const unsigned a_lut = ra_smem_u32(s_lut);
    const size_t felem = FDT == V2E_DT_U8 ? 1 : (FDT == V2E_DT_BF16 ? 2 : 4);

    const bool is_stream = warp < kRaStreamWarps;
    const int ts = tid;  // stream thread index (stream warps come first)
    long long* tl = prm.timeline ? prm.timeline + (size_t)blockIdx.x * V2E_MAX_FRAMES * 8 : nullptr;
    if (tl && tid == 0) {
        tl[0] = clock64();
        for (int n = 1; n < V2E_MAX_FRAMES; ++n) tl[n * 8 + 2] = tl[n * 8 + 5] = 0;
        for (int n = 16 * 8; n < V2E_MAX_FRAMES * 8; ++n) tl[n] = 0;
    }

    // ---- stream threads: issue the loads of the per-pixel planes before anything else ----------
    float4 ld_pos[kRaGroups], ld_neg[kRaGroups], ld_nra[kRaGroups], ld_b[kRaGroups],
        ld_l[kRaGroups], ld_t[kRaGroups];
    bool valid[kRaGroups];
#pragma unroll
    for (int g = 0; g < kRaGroups; ++g) {
        const int gl = g * kRaStreamThreads + ts;
        valid[g] = is_stream && gl < ngroups;
        ld_pos[g] = ld_neg[g] = make_float4(1, 1, 1, 1);
        ld_nra[g] = ld_b[g] = ld_l[g] = ld_t[g] = make_float4(0, 0, 0, 0);
        if (valid[g]) {
            const int p0 = tile_px0 + gl * 4;
            ld_pos[g] = __ldg((const float4*)(sq.pos + p0));
            ld_neg[g] = __ldg((const float4*)(sq.neg + p0));
            if (leak) ld_nra[g] = __ldg((const float4*)(sq.nra + p0));
            if (sq.first_pack) {
                ld_b[g] = load_frame4<FDT>(sq.frames, p0);  // frame 0 (raw values)
            } else {
                ld_b[g] = __ldcs((const float4*)(sq.base + p0));
                if (cutoff) ld_l[g] = __ldcs((const float4*)(sq.lp + p0));
                ld_t[g] = __ldcs((const float4*)(sq.tmem + p0));
            }
        }
    }
    // pull this tile's slice of every later frame into L2 (frame data depends on no state)
    static_assert(true, "");
    if (V2E_RA_L2PREFETCH && tid == kRaStreamThreads && ngroups > 0) {
        const unsigned char* src = (const unsigned char*)sq.frames + (size_t)tile_px0 * felem;
        const unsigned bytes = (unsigned)(tile_px * felem);
        if ((((size_t)P * felem) & 15) == 0)
            for (int n = 2; n < T; ++n)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(
                                 src + (size_t)n * P * felem),
                             "r"(bytes)
                             : "memory");
    }

    // ---- common initialisation of shared memory -------------------------------------------------
    load_lut(s_lut, cfg.lut);
    for (int i = tid; i < T; i += kRaThreads) s_fr[i] = prm.fr[seq][i];
    for (int i = tid; i < (int)(sizeof(RaCtl) / 4); i += kRaThreads) ((int*)ctl)[i] = 0;
    {
        uint4* p4 = (uint4*)s_park;
        for (int i = tid; i < tpa / 16; i += kRaThreads) p4[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (tl && tid == 0) tl[1] = clock64();

    long long nev = 0;
    // =============================================================================================
    // helpers of the dependent chain (used by the resolve warps, and by every warp in phase 2)
    // =============================================================================================
    unsigned* slots = wk.ra_slot;
    const float jit = cfg.leak_jitter;
    const float shot_m1 = cfg.shot_inten_m1;

    // index of a tile-local pixel in the [slot][group] planes
    auto slot_of = [&](int lidx) { return (lidx & 3) * kRaSlotStride + (lidx >> 2); };
    // (L, I) of frame m for a parked pixel: from the byte the stream warps left in shared memory
    // when they passed frame m, or (frame not streamed yet / arbitrary float frames / no room
    // for the byte planes) from global memory
    int streamed = 0;  // frames whose bytes are known to be complete
    bool bytes_ok = use_bytes;
    auto fetch_li = [&](const RaEntry& e, int m, float& L, float& I) {
        if (bytes_ok && m > streamed) {
            if (ra_lds_volatile(&ctl->stream_cnt[m]) >= kRaStreamWarps) {
                ra_fence_cta();
                streamed = m;
            }
            if (ra_lds_volatile(&ctl->nonint)) bytes_ok = false;
        }
        if (bytes_ok && m <= streamed) {
            const unsigned v = s_byte[(size_t)(m - 1) * tpa + e.lidx];
            const float2 li = ra_lds_f2(a_lut + 8u * v);
            L = li.x;
            I = li.y;
        } else {
            log_inten(s_lut, load_frame1<FDT>(sq.frames, (long long)m * P + tile_px0 + e.lidx), L, I);
        }
    };
    // state of a pixel parked at frame n: base / lp from shared memory (left there by the stream
    // thread that parked it, or by the resolve thread that carried it on), timestamp_mem from its
    // plane (a strong load, issued while the grid barrier is in flight)
    auto load_entry = [&](RaEntry& e, int lidx, int n) {
        (void)n;
        e.lidx = lidx;
        const int p = tile_px0 + lidx;
        e.tm = ra_ldg_strong(sq.tmem + p);
        e.b = s_base[lidx];
        e.l = cutoff ? s_lp[lidx] : 0.f;
        const int yy = p / cfg.W, xx = p - yy * cfg.W;
        e.small = ((yy | xx) & 1) == 0;
    };
    // What the events of a parked pixel at frame n need and what does NOT depend on the image-wide
    // max: computed while the grid barrier is in flight.
    struct RaPrep {
        PixelA a;    // polarity, threshold, count at frame n (v2e_model.py:375-404)
        float I;     // rescaled intensity at frame n (shot noise)
        float nw1;   // filtered value at frame n+1 (low-pass already advanced)
        float I1;    // rescaled intensity at frame n+1
    };
    // input side of frame m: low-pass (state advanced), returns the filtered value and I
    auto prep_step = [&](RaEntry& e, int m, float& nw, float& I) {
        float L;
        fetch_li(e, m, L, I);
        nw = L;
        if (cutoff) {
            e.l = lowpass(e.l, L, I, e.small ? s_fr[m].s_small : s_fr[m].s_large);
            nw = e.l;
        }
    };
    auto prep_entry = [&](RaEntry& e, int n, RaPrep& pr) {
        float L;
        fetch_li(e, n, L, pr.I);
        const float2 pn = s_pn[slot_of(e.lidx)];
        pr.a = diff_count(cutoff ? e.l : L, e.b, pn.x, pn.y);
        pr.nw1 = 0.f;
        pr.I1 = 0.f;
        if (n + 1 < T) prep_step(e, n + 1, pr.nw1, pr.I1);
    };
    // frame n: the events of a parked pixel (v2e_model.py:449-522)
    auto run_events = [&](RaEntry& e, int n, const RaPrep& pr, const FrameSc& fs, const FrameGlobals& fg) {
        const int p = tile_px0 + e.lidx;
        const PixelA a = pr.a;
        int mode = 0;
        float u = 0.f;
        if (shot && a.pol != 0.f && fg.mx > 0) {
            if (philox) {
                const uint4 w = noise_words<R>(sq.philox_ctr, n, (unsigned)(p >> 2), 0);
                const unsigned byte = (w.z >> (8 * (p & 3))) & 0xffu;
                const float thr256 = shot_thr256(a.pol > 0.f ? fs.kp : fs.kn, shot_m1, pr.I, a.C);
                if (ra_shot_candidate<R>(sq.philox_ctr, n, p, byte, thr256, u)) mode = 2;
            } else {
                mode = 1;
            }
        }
        // voxel accumulation: straight into the output grid (L2), in the reference's order per
        // address (one thread at a time owns a pixel; same-address atomics of a thread are
        // performed in program order)
        float* vp = g_vox + p;
        const int fin = iterate_pixel<R>(
            cfg, sq, fs, fg, n, p, a.pol, a.C, a.cnt, pr.I, e.tm, mode, u,
            [&](int bin, float val) { ra_red_add(vp + (size_t)bin * P, val); },
            [&](float, float) {});
        if (fin > 0) {
            ra_stg_strong(sq.tmem + p, e.tm);
            e.b = fadd(e.b, fmul(fmul(a.pol, (float)fin), a.C));  // memory update (:522)
            nev += fin;
            if (GENERIC && sq.dbg_final) sq.dbg_final[(size_t)(n - 1) * P + p] = fin;
        }
    };
    // state side of frame m (no events yet): leak, diff.  Returns true if the pixel fires (or is a
    // shot candidate) at m; cnt = its event count.
    auto finish_step = [&](RaEntry& e, int m, float nw, float I, int& cnt) -> bool {
        const int p = tile_px0 + e.lidx;
        const float2 pn = s_pn[slot_of(e.lidx)];
        unsigned shot_bytes = 0;
        if (leak || (shot && philox)) {
            float z = 0.f;
            if (philox) {
                const uint4 w = noise_words<R>(sq.philox_ctr, m, (unsigned)(p >> 2), 0);
                shot_bytes = w.z;
                if (leak) {
                    float z0, z1;
                    box_muller16((p & 2) ? w.y : w.x, z0, z1);
                    z = (p & 1) ? z1 : z0;
                }
            } else if (leak) {
                z = __ldg(sq.leak_randn + (size_t)(m - 1) * P + p);
            }
            if (leak) e.b = leak_sub(e.b, s_lkn[slot_of(e.lidx)], jit, z, s_fr[m].sdt, pn.x);
        }
        // diff / polarity / threshold as diff_count(); the count (a division) only if it fires
        const float diff = fsub(nw, e.b);
        const float ad = fabsf(diff);
        const bool live = ad > 1e-6f;
        const bool up = diff > 0.f;
        const float C = live ? (up ? pn.x : pn.y) : 0.f;
        const float den = fadd(C, 1e-9f);
        const bool cross = live && ad >= den;
        bool cand = false;
        if (shot && live) {
            if (philox) {
                const unsigned byte = (shot_bytes >> (8 * (p & 3))) & 0xffu;
                const float thr256 = shot_thr256(up ? s_fr[m].kp : s_fr[m].kn, shot_m1, I, C);
                float uu;
                cand = ra_shot_candidate<R>(sq.philox_ctr, m, p, byte, thr256, uu);
            } else {
                cand = true;
            }
        }
        cnt = 0;
        if (cross) cnt = (int)floorf(fdiv(ad, den));
        if (GENERIC) {
            if (sq.dbg_pol) sq.dbg_pol[(size_t)(m - 1) * P + p] = live ? (up ? 1 : -1) : 0;
            if (sq.dbg_cnt) sq.dbg_cnt[(size_t)(m - 1) * P + p] = cnt;
            if (sq.dbg_final) sq.dbg_final[(size_t)(m - 1) * P + p] = 0;
        }
        return cross || cand;
    };
    auto do_step = [&](RaEntry& e, int m, int& cnt) -> bool {
        float nw, I;
        prep_step(e, m, nw, I);
        return finish_step(e, m, nw, I, cnt);
    };
    // hand the pixel back: state -> shared memory; park byte = frame m it fires at next, or "done"
    auto put_back = [&](const RaEntry& e, int m) {
        s_base[e.lidx] = e.b;
        if (cutoff) s_lp[e.lidx] = e.l;
        s_park[e.lidx] = (unsigned char)(m < T ? m : kRaParkDone);
    };
    // the tile's parks at frame n are final: tell the other tiles its max (no flag, no fence:
    // the slot value itself says "arrived")
    auto arrive = [&](int n) {
        while (ra_lds_volatile(&ctl->stream_cnt[n]) < kRaStreamWarps) {}
        ra_fence_cta();
        ra_st_relaxed(&slots[n * kRaMaxTiles + tile], (unsigned)ra_lds_volatile(&ctl->lmax[n]) + 1u);
    };
    auto build_list = [&](int n, int rt, int nrt) {
        // every park byte of frame n written by the stream warps must be visible first.  Thread 0
        // has waited for that in arrive(n); the others only get here early in phase 1, where a
        // spinning resolve warp would steal issue slots from the stream warps: back off.
        while (ra_lds_volatile(&ctl->stream_cnt[n]) < kRaStreamWarps) __nanosleep(200);
        ra_fence_cta();
        const unsigned pat = (unsigned)n * 0x01010101u;
        const int chunks = tpa >> 4;
        for (int c = rt; c < chunks; c += nrt) {
            const uint4 w = *(const uint4*)(s_park + c * 16);
            const unsigned ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned m = __vcmpeq4(ws[k], pat);
                while (m) {
                    const int b = (__ffs(m) - 1) >> 3;
                    m &= ~(0xffu << (8 * b));
                    const int slot = atomicAdd(&ctl->listn[n & 1], 1);
                    s_list[slot] = (unsigned short)(c * 16 + k * 4 + b);
                }
            }
        }
    };


    if (is_stream) {
        // =========================================================================================
        // STREAM role
        // =========================================================================================
        float base[kRaPPT], lpv[kRaPPT];
        unsigned small_mask = 0;  // bit j: pixel j of this thread is a "small" (ps, qs) pixel
#pragma unroll
        for (int g = 0; g < kRaGroups; ++g) {
            const int gl = g * kRaStreamThreads + ts;
            float4 b4 = ld_b[g], l4 = ld_l[g], t4 = ld_t[g];
            if (valid[g]) {
                const int p0 = tile_px0 + gl * 4;
                if (sq.first_pack) {
                    // _init (v2e_model.py:170-174,239-240): base = lp = lin_log(frame 0), tmem = -Tr
                    const float4 x4 = b4;
                    float I;
                    log_inten(s_lut, x4.x, b4.x, I);
                    log_inten(s_lut, x4.y, b4.y, I);
                    log_inten(s_lut, x4.z, b4.z, I);
                    log_inten(s_lut, x4.w, b4.w, I);
                    l4 = b4;
                    const float t0 = fadd(0.f, -sq.Tr);
                    t4 = make_float4(t0, t0, t0, t0);
                } else {
                    // rollover (v2e_model.py:329-330)
                    float* tp = &t4.x;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float tm = tp[e];
                        if (tm > 0.f) tm = fsub(tm, sq.bins_m1);
                        if (tm < 0.f) tm = -sq.Tr;
                        tp[e] = tm;
                    }
                }
                // timestamp_mem lives in its state plane for the whole pack: the resolve warps
                // update the entries of the pixels that fire
                *(float4*)(sq.tmem + p0) = t4;
                // the output grid doubles as the accumulator: zero it before any event can land
                for (int b = 0; b < nb; ++b)
                    *(float4*)(g_vox + (size_t)b * P + p0) = make_float4(0, 0, 0, 0);
                // the constant planes stay in shared memory for both roles
                // layout [slot e][group]: the accesses of a warp are conflict-free
                {
                    const float* pp = &ld_pos[g].x; const float* np_ = &ld_neg[g].x;
                    const float* rp = &ld_nra[g].x;
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s_pn[e * kRaSlotStride + gl] = make_float2(pp[e], np_[e]);
                        if (leak) s_lkn[e * kRaSlotStride + gl] = fmul(cfg.leak_rate, rp[e]);
                    }
                }
                const int y = p0 / cfg.W, x = p0 - y * cfg.W;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int xe = x + e, ye = y;
                    if (xe >= cfg.W) { xe -= cfg.W; ye += 1; }  // a group may straddle a row end
                    if (xe >= cfg.W) { xe -= cfg.W; ye += 1; }
                    if (xe >= cfg.W) { xe -= cfg.W; ye += 1; }
                    if (((xe | ye) & 1) == 0) small_mask |= 1u << (g * 4 + e);
                }
            }
            const float* bp = &b4.x; const float* lp_ = &l4.x;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = g * 4 + e;
                base[j] = bp[e]; lpv[j] = lp_[e];
            }
        }

        // shared-window addresses kept opaque so that they stay in registers (the compiler
        // otherwise rebuilds the generic->shared conversion at every use)
        unsigned a_lut_r = a_lut, a_pn = ra_smem_u32(s_pn), a_lk = ra_smem_u32(s_lkn),
                 a_park = ra_smem_u32(s_park), a_byte = ra_smem_u32(s_byte);
        asm volatile("" : "+r"(a_lut_r), "+r"(a_pn), "+r"(a_lk), "+r"(a_park), "+r"(a_byte));
        unsigned dead = 0;     // bit j: pixel j is parked (its state belongs to the resolve warps)
        unsigned err_hi = 0;   // INTF: OR of (x + 1.5*2^23) ^ 0x4B400000: bits >= 8 set <=> out of 0..255
        unsigned err_lo = 0;   // INTF: OR of (rounded x) ^ x: non-zero <=> not an integer
        bool nonint = false;
        // running frame pointers of my groups (next frame = + P elements)
        const unsigned char* fptr[kRaGroups];
        RaVec<FDT> xcur[kRaGroups];
        const size_t fstride = (size_t)P * felem;
#pragma unroll
        for (int g = 0; g < kRaGroups; ++g) {
            const int gl = g * kRaStreamThreads + ts;
            fptr[g] = (const unsigned char*)sq.frames + ((size_t)P + tile_px0 + gl * 4) * felem;
            xcur[g] = RaVec<FDT>{};
            if (valid[g] && T > 1) xcur[g] = ra_load4<FDT>(fptr[g], 0);
        }
        for (int n = 1; n < T; ++n) {
            const float s_small = s_fr[n].s_small, s_large = s_fr[n].s_large;
            const float sdt = s_fr[n].sdt, kp = s_fr[n].kp, kn = s_fr[n].kn;
            int lm_frame = 0;  // max event count among the pixels this thread parks at frame n
#pragma unroll
            for (int g = 0; g < kRaGroups; ++g) {
                if (!valid[g]) continue;
                const int gl = g * kRaStreamThreads + ts;
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                unsigned shot_bytes = 0;
                // thresholds {pos, neg} / leak rate of pixel e of this group: plane[e][gl]
                const unsigned a_g = a_pn + (unsigned)gl * 8u;
                const unsigned a_gl = a_lk + (unsigned)gl * 4u;
                if (philox) {
                    const uint4 w = noise_words<R>(sq.philox_ctr, n, (unsigned)(g0 + gl), 0);
                    shot_bytes = w.z;
                    if (leak) {
                        box_muller16(w.x, z[0], z[1]);
                        box_muller16(w.y, z[2], z[3]);
                        if (GENERIC && sq.dbg_leak_randn)
                            *(float4*)(sq.dbg_leak_randn + (size_t)(n - 1) * P + tile_px0 + gl * 4) =
                                make_float4(z[0], z[1], z[2], z[3]);
                    }
                } else if (leak) {
                    const float4 zz = __ldg(
                        (const float4*)(sq.leak_randn + (size_t)(n - 1) * P + tile_px0 + gl * 4));
                    z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
                }
                unsigned flag = 0;  // bit e: pixel e may have to be parked at this frame
                unsigned bytes4 = 0;  // the four frame values as bytes (for the resolve warps)
                float Lk[4];        // lin-log value of the four pixels (the rare path needs it)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    // ---- (L, I) of two pixels: LUT for integer-valued 0..255 -------------------
                    float2 li[2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = 2 * h + q;
                        if constexpr (FDT == V2E_DT_U8) {
                            li[q] = ra_lds_f2(a_lut_r + 8u * ((xcur[g].v >> (8 * e)) & 0xffu));
                            bytes4 = xcur[g].v;
                        } else {
                            float xv;
                            bool okdummy;
                            (void)ra_elem<FDT>(xcur[g], e, okdummy, xv);
                            const float y = __fadd_rn(xv, 12582912.0f);
                            const unsigned yb = __float_as_uint(y);
                            li[q] = ra_lds_f2(a_lut_r + 8u * (yb & 0xffu));
                            bytes4 |= (yb & 0xffu) << (8 * e);
                            if (INTF) {
                                err_hi |= yb ^ 0x4B400000u;
                                err_lo |= __float_as_uint(__fsub_rn(y, 12582912.0f)) ^ __float_as_uint(xv);
                            } else if (__builtin_expect(
                                           (yb & 0xffffff00u) != 0x4B400000u ||
                                               __fsub_rn(y, 12582912.0f) != xv, 0)) {
                                li[q] = log_inten_slow(xv);  // fp64 lin_log, emulator_utils.py:13-37
                                nonint = true;               // (its byte is meaningless)
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int e = 2 * h + q;
                        const int j = g * 4 + e;
                        const float L = li[q].x, I = li[q].y;
                        Lk[e] = L;
                        float nw = L;
                        if (cutoff) {
                            const float s = (small_mask >> j) & 1u ? s_small : s_large;
                            lpv[j] = lowpass(lpv[j], L, I, s);
                            nw = lpv[j];
                        }
                        const float2 pn = ra_lds_f2(a_g + (unsigned)(e * kRaSlotStride * 8));
                        if (leak)
                            base[j] = leak_sub(base[j], ra_lds_f1(a_gl + (unsigned)(e * kRaSlotStride * 4)),
                                               jit, z[e], sdt, pn.x);
                        // diff / polarity / threshold (v2e_model.py:375,394-404), see diff_count()
                        const float diff = fsub(nw, base[j]);
                        const float ad = fabsf(diff);
                        const bool live = ad > 1e-6f;
                        const bool up = diff > 0.f;
                        const float C = up ? pn.x : pn.y;
                        bool maybe = live && ad >= fadd(C, 1e-9f);
                        if (shot) {
                            if (philox) {
                                // first stage of the candidate test (0 threshold if !live)
                                const float thr256 = shot_thr256(up ? kp : kn, shot_m1, I, live ? C : 0.f);
                                maybe = maybe ||
                                        ra_byte_to_float((shot_bytes >> (8 * e)) & 0xffu) < thr256;
                            } else {
                                maybe = maybe || live;  // injected draws: every pol != 0 pixel iterates
                            }
                        }
                        if (GENERIC && !((dead >> j) & 1u)) {
                            const size_t o = (size_t)(n - 1) * P + tile_px0 + gl * 4 + e;
                            if (sq.dbg_pol) sq.dbg_pol[o] = live ? (up ? 1 : -1) : 0;
                            if (sq.dbg_cnt) sq.dbg_cnt[o] = 0;
                            if (sq.dbg_final) sq.dbg_final[o] = 0;
                        }
                        if (maybe) flag |= 1u << e;
                    }
                }
                flag &= ~(dead >> (g * 4));
                if (use_bytes)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a_byte + (unsigned)((n - 1) * tpa + gl * 4)),
                                 "r"(bytes4) : "memory");
                const RaVec<FDT> xprev = xcur[g];  // (dead unless a shot candidate needs I again)
                // this group's pixels of the next frame: in flight while the other group and the
                // next frame's first group are computed
                fptr[g] += fstride;
                if (n + 1 < T) xcur[g] = ra_load4<FDT>(fptr[g], 0);
                if (flag) {
                    // rare per thread (~3 % of pixel-frames): exact count, second stage of the shot
                    // test, park.  Everything is recomputed from the registers of the pixel.
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (!((flag >> e) & 1u)) continue;
                        const int j = g * 4 + e;
                        const float nw = cutoff ? lpv[j] : Lk[e];
                        const float diff = fsub(nw, base[j]);
                        const float ad = fabsf(diff);
                        const bool up = diff > 0.f;
                        const float2 pn = ra_lds_f2(a_g + (unsigned)(e * kRaSlotStride * 8));
                        const float C = up ? pn.x : pn.y;
                        const float den = fadd(C, 1e-9f);
                        // floor(ad/den) == 1 <=> den <= ad < 2*den (same argument as diff_count)
                        int cnt = 0;
                        if (ad >= den) cnt = ad < fmul(2.0f, den) ? 1 : (int)floorf(fdiv(ad, den));
                        bool cand = false;
                        if (shot) {
                            if (philox) {
                                // I = (v+20)/275 from the LUT again (not kept in a register)
                                bool ok;
                                float xv;
                                const int v = ra_elem<FDT>(xprev, e, ok, xv) & 0xff;
                                const float I = ok ? ra_lds_f2(a_lut_r + 8u * (unsigned)v).y
                                                   : log_inten_slow(xv).y;
                                const float thr256 = shot_thr256(up ? kp : kn, shot_m1, I, C);
                                float u;
                                cand = ra_shot_candidate<R>(sq.philox_ctr, n, tile_px0 + gl * 4 + e,
                                                            (shot_bytes >> (8 * e)) & 0xffu, thr256, u);
                            } else {
                                cand = true;
                            }
                        }
                        if (cnt > 0 || cand) {
                            // state as of frame n (before its events) -> shared memory; ordered
                            // before the stream_cnt[n] increment by the fence below
                            s_base[gl * 4 + e] = base[j];
                            if (cutoff) s_lp[gl * 4 + e] = lpv[j];
                            if (GENERIC && sq.dbg_cnt)
                                sq.dbg_cnt[(size_t)(n - 1) * P + tile_px0 + gl * 4 + e] = cnt;
                            asm volatile("st.shared.u8 [%0], %1;" ::"r"(a_park + (unsigned)(gl * 4 + e)),
                                         "r"(n) : "memory");
                            lm_frame = max(lm_frame, cnt);
                            dead |= 1u << j;
                        }
                    }
                }
            }
            if (!INTF && nonint) ctl->nonint = 1;  // before the release below
            // this warp has streamed frame n: publish frame bytes, park bytes, parked state and max
            lm_frame = __reduce_max_sync(0xffffffffu, lm_frame);
            if (lane == 0) {
                if (lm_frame > 0) atomicMax(&ctl->lmax[n], lm_frame);
                ra_fence_cta();
                atomicAdd(&ctl->stream_cnt[n], 1);
                if (tl) atomicMax((unsigned long long*)&tl[n * 8 + 0], (unsigned long long)clock64());
            }
        }
        if (INTF && ((err_hi & 0xffffff00u) != 0u || err_lo != 0u)) nonint = true;

        if (INTF && nonint) ctl->nonint = 1;
        // ---- write back the state of the pixels that never left the streaming registers ---------
#pragma unroll
        for (int g = 0; g < kRaGroups; ++g) {
            if (!valid[g]) continue;
            const int gl = g * kRaStreamThreads + ts;
            const int p0 = tile_px0 + gl * 4;
            if (((dead >> (g * 4)) & 0xfu) == 0) {
                *(float4*)(sq.base + p0) =
                    make_float4(base[g * 4], base[g * 4 + 1], base[g * 4 + 2], base[g * 4 + 3]);
                if (cutoff)
                    *(float4*)(sq.lp + p0) =
                        make_float4(lpv[g * 4], lpv[g * 4 + 1], lpv[g * 4 + 2], lpv[g * 4 + 3]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    if (!((dead >> j) & 1u)) {
                        sq.base[p0 + e] = base[j];
                        if (cutoff) sq.lp[p0 + e] = lpv[j];
                    }
                }
            }
        }
    }

    // =============================================================================================
    // The dependent chain: rounds n = 1..T-1.  Phase 1: the resolve warps alone (named barrier 1)
    // while the stream warps are busy.  Phase 2: once the stream warps have nothing left to
    // stream they join (named barrier 2, all warps), so that every parked pixel of a round gets
    // its own thread.  The hand-over happens at a round boundary.
    //
    // Round n, per thread:   wait for the image-wide max of frame n
    //                        events of my pixel at n, then its frame n+1        <- critical path
    //                        publish the tile max of frame n+1 (thread 0)
    //                        list the pixels parked at n+1, fetch mine           } while the grid
    //                        carry my pixel of round n on through n+2..          } barrier flies
    // =============================================================================================
    {
        int rt = tid - kRaStreamThreads;  // resolve-thread index (stream threads: set below)
        int nrt = kRaResolveThreads, bar = 1;
        bool phase1 = !is_stream;
        int n_first = 1;
        int ln = 0;
        // Phase 2 has four times more threads than a round has parked pixels: consecutive rounds
        // use disjoint quarters of the threads, so that the thread that carries a pixel of round
        // n on through n+2.. (the "tail") is not the one that serves a pixel of round n+1.
        auto slot_in_round = [&](int n) {
            // phase 2: consecutive rounds use disjoint blocks of 8 warps (256 list slots), so
            // that the warps carrying last round's quiet pixels on are not the ones serving
            // this round's parked pixels
            return phase1 ? rt : (rt + kRaThreads - (256 * n) % kRaThreads) % kRaThreads;
        };
        const int nslots2 = kRaThreads;
        RaEntry cur, tle;   // my pixel of this round; my pixel of the last round still to carry on
        RaPrep prep;        // what was computed for `cur` ahead of the barrier release
        int tle_m = 0;      // ... from this frame on
        cur.lidx = tle.lidx = -1;
        if (is_stream) {
            if (tl && tid == 0) tl[2] = clock64();
            ra_named_sync(2, kRaThreads);  // meets the resolve warps at their hand-over
            n_first = ra_lds_volatile(&ctl->phase2);
            rt = tid + kRaResolveThreads;
            nrt = kRaThreads;
            bar = 2;
            if (n_first < T) {
                ln = ra_lds_volatile(&ctl->listn[n_first & 1]);
                const int q = slot_in_round(n_first);
                if (q < ln) { load_entry(cur, s_list[q], n_first); prep_entry(cur, n_first, prep); }
            }
        } else if (T > 1) {
            if (rt == 0) arrive(1);
            build_list(1, rt, nrt);
            ra_named_sync(1, kRaResolveThreads);
            ln = ra_lds_volatile(&ctl->listn[1]);
            if (rt < ln) { load_entry(cur, s_list[rt], 1); prep_entry(cur, 1, prep); }
        }
        const bool poller = warp == kRaStreamWarps;  // first resolve warp talks to the other tiles
        for (int n = n_first; n < T; ++n) {
            if (tl && rt == 0) { tl[n * 8 + 1] = clock64(); tl[n * 8 + 7] = ln; }
            // ---- wait for the other tiles: image-wide max of frame n (v2e_model.py:405-406) -------
            if (poller) {
                int mx;
                bool all;
                do {
                    mx = 0;
                    all = true;
                    for (int i = lane; i < tiles; i += 32) {
                        const unsigned v = ra_ld_relaxed(&slots[n * kRaMaxTiles + i]);
                        all = all && v != 0u;
                        mx = max(mx, (int)v - 1);
                    }
                    all = __all_sync(0xffffffffu, all);
                } while (!all);
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (lane == 0) {
                    ctl->mx = mx;
                    ctl->listn[(n + 1) & 1] = 0;  // list of the next round: unused since round n-1
                    if (tile == 0 && sq.max_iters) sq.max_iters[n] = mx;
                    if (tl) tl[n * 8 + 4] = clock64();
                }
            }
            ra_named_sync(bar, nrt);
            const FrameGlobals fg = frame_globals(ra_lds_volatile(&ctl->mx), sq.duration, sq.Tr);
            const FrameSc fs = s_fr[n];
            const long long t_sec = tl ? clock64() : 0;

            // ---- my parked pixel: events at n, then frame n+1 (what the next barrier needs) --------
            int lm_next = 0;
            if (cur.lidx >= 0) {
                run_events(cur, n, prep, fs, fg);
                int c = 0;
                const bool quiet = n + 1 < T && !finish_step(cur, n + 1, prep.nw1, prep.I1, c);
                if (quiet) {
                    tle = cur;  // quiet at n+1: carried on during the next round
                    tle_m = n + 2;
                } else {
                    put_back(cur, n + 1);
                    lm_next = max(lm_next, c);
                }
                cur.lidx = -1;
            }
            // more parked pixels than threads (dense events, phase 1): carried to their next park here
            for (int q = slot_in_round(n) + (phase1 ? nrt : nslots2); q < ln; q += (phase1 ? nrt : nslots2)) {
                RaEntry ex;
                RaPrep px;
                load_entry(ex, s_list[q], n);
                prep_entry(ex, n, px);
                run_events(ex, n, px, fs, fg);
                int m = n + 1, c = 0;
                if (m < T && !finish_step(ex, m, px.nw1, px.I1, c)) ++m;
                else if (m < T) { put_back(ex, m); lm_next = max(lm_next, c); continue; }
                for (; m < T; ++m)
                    if (do_step(ex, m, c)) break;
                put_back(ex, m);
                if (m == n + 1) lm_next = max(lm_next, c);
                else if (m < T && c > 0) atomicMax(&ctl->lmax[m], c);
            }
            lm_next = __reduce_max_sync(0xffffffffu, lm_next);
            if (lane == 0 && lm_next > 0 && n + 1 < T) atomicMax(&ctl->lmax[n + 1], lm_next);
            if (tl && rt == 0) tl[n * 8 + 5] = clock64();
            if (tl && lane == 0 && (n == 5 || n == 6))   // per-warp time of the critical section
                tl[(16 + (n - 5) * 4) * 8 + warp] = clock64() - t_sec;
            ra_fence_cta();
            ra_named_sync(bar, nrt);  // parks at n+1 (and later) and lmax[n+1] are complete
            if (rt == 0) {
                if (n + 1 < T) arrive(n + 1);
                if (phase1) ctl->handover = ra_lds_volatile(&ctl->stream_cnt[T - 1]) >= kRaStreamWarps;
                if (tl) tl[n * 8 + 3] = clock64();
            }

            // ---- while the barrier of frame n+1 is in flight: list its pixels, fetch mine ------------
            if (n + 1 < T) {
                build_list(n + 1, rt, nrt);
                ra_named_sync(bar, nrt);
                ln = ra_lds_volatile(&ctl->listn[(n + 1) & 1]);
                if (phase1 && ra_lds_volatile(&ctl->handover)) {
                    // the stream warps are waiting at barrier 2: from round n+1 on everybody works
                    if (rt == 0) ctl->phase2 = n + 1;
                    ra_named_sync(2, kRaThreads);
                    phase1 = false;
                    nrt = kRaThreads;
                    bar = 2;
                }
                const int q = slot_in_round(n + 1);
                if (q < ln) { load_entry(cur, s_list[q], n + 1); prep_entry(cur, n + 1, prep); }
            }
            // ---- also in the shadow of that barrier: the pixel of this round that stayed quiet at
            //      n+1 is carried on through n+2.. until it fires again (re-park) or the pack ends.
            //      (Its park byte is in place before the list of frame n+2 is built: that happens
            //      after the next round's second barrier.)  In phase 2 the threads that serve round
            //      n+1 are other warps than these. ------------------------------------------------
            if (tle.lidx >= 0) {
                int m = tle_m, c = 0;
                for (; m < T; ++m)
                    if (do_step(tle, m, c)) break;
                put_back(tle, m);
                if (m < T && c > 0) atomicMax(&ctl->lmax[m], c);
                tle.lidx = -1;
            }
            if (tl && rt == 0) tl[n * 8 + 6] = clock64();
        }
        if (phase1) {  // the resolve warps finished the chain on their own
            if (rt == 0) ctl->phase2 = T;
            ra_named_sync(2, kRaThreads);
        }
    }

    // ---- both roles are done with the pack ----------------------------------------------------------
    __syncthreads();
    if (tl && tid == 0) tl[3] = clock64();

    // ---- the pixels the chain owned: final base / lp from shared memory to their planes ---------
    for (int c = tid; c < (tpa >> 2); c += kRaThreads) {
        const unsigned w = *(const unsigned*)(s_park + c * 4);
        if (w == 0u) continue;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (((w >> (8 * e)) & 0xffu) == 0u) continue;
            const int lidx = c * 4 + e;
            sq.base[tile_px0 + lidx] = s_base[lidx];
            if (cutoff) sq.lp[tile_px0 + lidx] = s_lp[lidx];
        }
    }

    if (tl && tid == 0) tl[31 * 8 + 0] = clock64();
    // ---- statistics of the un-normalised grid (utils/event_process.py:166-174) ---------------------
    // The accumulators sit in the output grid (L2): every red of this CTA was issued before the
    // barrier above and is ordered by the fences of the chain; read them back with strong loads.
    const bool normalise = cfg.output_mode == V2E_OUT_VOXEL;
    const int nchunks = nb * ngroups;  // float4 chunks of this tile's part of the grid
    auto chunk_ptr = [&](int i) {
        const int b = i / ngroups, gl = i - b * ngroups;
        return g_vox + (size_t)b * P + tile_px0 + gl * 4;
    };
    double s1 = 0., s2 = 0., nz = 0.;
    if (normalise) {
        for (int i = tid; i < nchunks; i += kRaThreads) {
            const float4 v4 = ra_ldg_strong4(chunk_ptr(i));
            const float* vp = &v4.x;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (vp[e] != 0.f) {
                    s1 += (double)vp[e];
                    s2 += (double)vp[e] * (double)vp[e];
                    nz += 1.;
                }
        }
    }
    if (tl && tid == 0) tl[31 * 8 + 1] = clock64();
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nz += __shfl_down_sync(0xffffffffu, nz, o);
        nev += __shfl_down_sync(0xffffffffu, nev, o);
    }
    if (lane == 0) {
        ctl->dred[warp * 4 + 0] = s1;
        ctl->dred[warp * 4 + 1] = s2;
        ctl->dred[warp * 4 + 2] = nz;
        ctl->dred[warp * 4 + 3] = (double)nev;
    }
    __syncthreads();
    if (tid < 32) {
        // warp 0: publish this tile's partial sums (bit-inverted: 0 = not there yet; no flag, no
        // fence), wait for the other tiles, combine
        if (tid < 4) {
            double a = 0.;
            for (int i = 0; i < kRaWarps; ++i) a += ctl->dred[i * 4 + tid];
            ra_st_relaxed64(&wk.ra_stat[4 * tile + tid], ~(unsigned long long)__double_as_longlong(a));
        }
        if (tl && tid == 0) tl[31 * 8 + 2] = clock64();
        // fixed assignment tile -> lane and a fixed shuffle tree: deterministic, and identical in
        // every CTA of the sequence
        double a = 0., b = 0., c = 0., d = 0.;
        for (int i = tid; i < tiles; i += 32) {
            unsigned long long v[4];
            bool all;
            do {  // the four loads are in flight together
                all = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = ra_ld_relaxed64(&wk.ra_stat[4 * i + k]);
#pragma unroll
                for (int k = 0; k < 4; ++k) all = all && v[k] != 0ull;
            } while (!all);
            a += __longlong_as_double((long long)~v[0]);
            b += __longlong_as_double((long long)~v[1]);
            c += __longlong_as_double((long long)~v[2]);
            d += __longlong_as_double((long long)~v[3]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
            d += __shfl_xor_sync(0xffffffffu, d, o);
        }
        if (tid == 0) {
            float mean = 0.f, den = 1.f, on = 0.f;
            if (c > 0. && normalise) {
                mean = fdiv((float)a, (float)c);
                den = fadd(sqrtf(fsub(fdiv((float)b, (float)c), fmul(mean, mean))), 1e-8f);
                on = 1.f;
            }
            if (tl) tl[4] = clock64();
            ctl->stat[0] = mean; ctl->stat[1] = den; ctl->stat[2] = on;
            if (tile == 0) *sq.num_events = (unsigned long long)d;
            if (INTF && sq.status && ctl->nonint) *sq.status = V2E_STATUS_NONINTEGER_FRAMES;
        }
    }
    __syncthreads();

    // ---- normalise in place (utils/event_process.py:175): only chunks that hold events change ----
    if (ctl->stat[2] != 0.f) {
        const float mean = ctl->stat[0], den = ctl->stat[1];
        for (int i = tid; i < nchunks; i += kRaThreads) {
            float* ptr = chunk_ptr(i);
            float4 v = ra_ldg_strong4(ptr);
            if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
            float* vp = &v.x;
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (vp[e] != 0.f) vp[e] = fdiv(fsub(vp[e], mean), den);
            __stcs((float4*)ptr, v);
        }
    }

    // ---- leave the work area zeroed: the last CTA of the sequence to get here cleans up -----------
    if (tid == 0) {
        if (tl) tl[5] = clock64();
        const unsigned prev = atomicAdd(&wk.exit_count, 1u);
        ctl->last = prev == (unsigned)tiles - 1;
    }
    __syncthreads();
    if (ctl->last) {
        // every other CTA of the sequence has finished all its polls
        for (int i = tid; i < T * kRaMaxTiles; i += kRaThreads) wk.ra_slot[i] = 0u;
        for (int i = tid; i < 4 * kRaMaxTiles; i += kRaThreads) wk.ra_stat[i] = 0ull;
        if (tid == 0) wk.exit_count = 0;
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct RaPlan {
    int ok;
    int tiles_per_seq, tile_groups, use_bytes;
    size_t smem;
};

static int g_ra_sms = -1, g_ra_smem_optin = 0, g_ra_coop = 0;

static int ra_device_props() {
    if (g_ra_sms >= 0) return V2E_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return V2E_ERR_NO_DEVICE;
    int major = 0, sms = 0, optin = 0, coop = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (major < 10) return V2E_ERR_NO_DEVICE;
    g_ra_smem_optin = optin;
    g_ra_coop = coop;
    g_ra_sms = sms;
    return V2E_OK;
}

// T <= 0: plan without the frame-byte store (the "does it fit at all" question)
static RaPlan ra_make_plan(const CfgDev& cfg, int n_seq, int T) {
    RaPlan pl{};
    if (ra_device_props() != V2E_OK || !g_ra_coop) return pl;
    if (cfg.P % 4 != 0 || cfg.output_mode == V2E_OUT_RAW) return pl;
    const int total_groups = cfg.P / 4;
    int tiles = g_ra_sms / n_seq;  // CTAs (= SMs) available per sequence
    if (tiles < 1) return pl;
    // small images: at least 1024 pixels per tile (fewer tiles = cheaper grid barriers)
    const int min_groups = 256;
    int tg = (total_groups + tiles - 1) / tiles;
    if (tg < min_groups) tg = min_groups;
    tg = (tg + 3) & ~3;  // tile boundaries are multiples of 16 pixels
    if (tg * 4 > kRaTileMaxPx) return pl;
    tiles = (total_groups + tg - 1) / tg;
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const size_t tpa = (size_t)tg * 4;
    if (cfg.voxel_bf16) return pl;  // the f32 output grid doubles as the accumulator
    // header + threshold planes (+ leak plane) + parked base (+ lp) + park bytes + park list
    const size_t fixed = kRaHeaderBytes + (size_t)kRaSlotStride * 4 * 8 +
                         (leak ? (size_t)kRaSlotStride * 4 * 4 : 0) + tpa * 4 * (cfg.cutoff_on ? 2 : 1) +
                         tpa + tpa * 2;
    if (fixed > (size_t)g_ra_smem_optin) return pl;
    size_t smem = fixed;
    int use_bytes = 0;
    if (T > 1 && fixed + tpa * (size_t)(T - 1) <= (size_t)g_ra_smem_optin) {
        use_bytes = 1;  // one byte per pixel and frame for the resolve warps
        smem = fixed + tpa * (size_t)(T - 1);
    }
    pl.ok = 1;
    pl.tiles_per_seq = tiles;
    pl.tile_groups = tg;
    pl.use_bytes = use_bytes;
    pl.smem = smem;
    return pl;
}

int runahead_fits(const CfgDev& cfg, int frame_dtype, int n_seq) {
    (void)frame_dtype;
    const int rc = ra_device_props();
    if (rc != V2E_OK) return rc;
    return ra_make_plan(cfg, n_seq, 0).ok;
}

template <int FDT, int R, int NOISE, bool CUTOFF, bool INTF>
static int launch_runahead(const RaParams& prm, const RaPlan& pl, cudaStream_t st) {
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_runahead<FDT, R, NOISE, CUTOFF, INTF>;
    int rc = V2E_OK;
    if (!attr_set) {
        rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          g_ra_smem_optin));
        if (rc != V2E_OK) return rc;
        attr_set = true;
    }
    void* args[] = {(void*)&prm};
    const dim3 grid((unsigned)(pl.tiles_per_seq * prm.n_seq));
    rc = cuda_ok(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(kRaThreads), args,
                                             pl.smem, st));
    count_launch();
    return rc;
}

template <int FDT>
static int run_runahead_t(const CfgDev& cfg, int rounds, const v2e_pack* packs, int n_seq,
                          void* workspace, cudaStream_t st) {
    const int T = packs[0].sc.num_frames;
    const RaPlan pl = ra_make_plan(cfg, n_seq, T);
    if (!pl.ok) return V2E_ERR_UNSUPPORTED;
    static thread_local RaParams prm;  // 16 KB: keep it off the stack
    prm.cfg = cfg;
    prm.n_seq = n_seq;
    prm.tiles_per_seq = pl.tiles_per_seq;
    prm.tile_groups = pl.tile_groups;
    prm.use_bytes = pl.use_bytes;
    prm.work = (SeqWork*)workspace;
    static const bool want_timeline = getenv("V2E_TIMELINE") != nullptr;
    // the stepwise engine's accumulator area of the workspace is unused here: borrow it
    prm.timeline = want_timeline ? (long long*)((char*)workspace + step_layout(cfg, n_seq).vox_off)
                                 : nullptr;
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
    // fast variants: no debug outputs, and either no noise or PHILOX leak+shot
    int noise = 2;
    if (!dbg) {
        if (!leak && !shot) noise = 0;
        else if (philox && leak && shot) noise = 1;
    }
    const bool r7 = rounds == 7;
    // INTF (no fp64 lin_log path in the streaming loop): uint8 frames, or the caller's guarantee
    constexpr bool kU8 = FDT == V2E_DT_U8;
    const bool intf = kU8 || cfg.frames_integer;
#define V2E_RA_LAUNCH(RR, NN, CC)                                                   \
    (intf ? launch_runahead<FDT, RR, NN, CC, true>(prm, pl, st)                     \
          : launch_runahead<FDT, RR, NN, CC, kU8>(prm, pl, st))
    if (noise == 0) return cfg.cutoff_on ? V2E_RA_LAUNCH(10, 0, true) : V2E_RA_LAUNCH(10, 0, false);
    if (noise == 1) {
        if (cfg.cutoff_on) return r7 ? V2E_RA_LAUNCH(7, 1, true) : V2E_RA_LAUNCH(10, 1, true);
        return r7 ? V2E_RA_LAUNCH(7, 1, false) : V2E_RA_LAUNCH(10, 1, false);
    }
    return r7 ? V2E_RA_LAUNCH(7, 2, false) : V2E_RA_LAUNCH(10, 2, false);
#undef V2E_RA_LAUNCH
}

int run_runahead(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st) {
    switch (frame_dtype) {
        case V2E_DT_U8: return run_runahead_t<V2E_DT_U8>(cfg, rounds, packs, n_seq, workspace, st);
        case V2E_DT_BF16: return run_runahead_t<V2E_DT_BF16>(cfg, rounds, packs, n_seq, workspace, st);
        default: return run_runahead_t<V2E_DT_F32>(cfg, rounds, packs, n_seq, workspace, st);
    }
}

}  // namespace v2e
