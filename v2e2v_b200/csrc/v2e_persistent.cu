// v2e_persistent.cu — the "persistent" engine: ONE cooperative launch per pack.
//
// Each CTA owns a contiguous tile of pixels of one sequence for the whole pack:
//   registers      base_log_frame, lp_log_frame0, pos/neg thresholds, leak_rate*noise_rate
//   shared memory  timestamp_mem and the num_bins voxel accumulators of the tile (only touched
//                  by the ~3 % of pixel-frames that emit events), the lin-log LUT, and a ring of
//                  frame tiles filled by TMA bulk copies (cp.async.bulk + mbarrier) that run
//                  ahead of the computation
//   HBM            each frame pixel is read once, each state plane once in / once out, the
//                  normalised voxel grid is written once: 92 B/pixel/pack
// The only cross-CTA communication is the per-frame image-wide max event count
// (v2e/v2e_model.py:405-406), which decides the event time stamps of that frame: one atomicMax +
// one arrival counter per frame per sequence (release/acquire at gpu scope), and one more
// barrier for the normalisation statistics (utils/event_process.py:166-175).  CTAs of different
// sequences never wait for each other.  Cooperative launch guarantees co-residency.
#include <cstdlib>

#include "v2e_engine.cuh"

namespace v2e {

constexpr int kPThreads = 896;  // 28 warps = 7 per SMSP, one CTA per SM, <= 72 registers per thread
constexpr int kPPT = 8;         // pixels per thread (2 groups of 4)
constexpr int kGroups = kPPT / 4;
constexpr int kTileMaxPx = kPThreads * kPPT;
constexpr int kHeaderBytes = 6144;  // LUT, mbarrier, reduction scratch, TileCtx, candidates, frame scalars
constexpr int kCandCap = 64;        // shot candidates per CTA-frame kept with their uniform
constexpr int kWarps = kPThreads / 32;
static_assert(kTileMaxPx <= 8192, "queue entries hold a 13-bit pixel index");

struct PersistParams {
    CfgDev cfg;
    int n_seq;
    int tiles_per_seq;
    int tile_groups;   // 4-pixel groups per tile (last tile of a sequence may hold fewer)
    int stages;        // frame ring depth (0: frames are read straight from global memory)
    int frame_elem;    // bytes per frame element
    SeqWork* work;
    long long* timeline;  // development aid (V2E_TIMELINE=1): [cta][frame][8] clock64 stamps
    SeqDev seq[V2E_MAX_SEQ_PER_LAUNCH];
    FrameSc fr[V2E_MAX_SEQ_PER_LAUNCH][V2E_MAX_FRAMES];
};

// ---- PTX wrappers: mbarrier + 1-D TMA bulk copy ------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes,
                                            unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Per-frame barrier word: low 32 bits = arrivals, high 32 bits = running max of the event counts.
// A CTA publishes its max (relaxed atomicMax on the high half) and then arrives with release
// semantics on the low half; both halves live in one 8-byte word served by one L2 slice, so a
// single relaxed 64-bit load that observes all arrivals also observes every max.  Polling uses
// relaxed (L2) loads: an acquire load per poll would invalidate the SM's L1 every iteration.
__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int frame_barrier(unsigned long long* bar, int my_max, unsigned count) {
    if (my_max > 0) atomicMax((int*)bar + 1, my_max);
    red_release_add((unsigned*)bar, 1u);
    unsigned long long v;
    do { v = ld_relaxed64(bar); } while ((unsigned)v < count);
    return (int)(v >> 32);
}
// Barrier that also orders ordinary data written before it (used once, for the statistics).
__device__ __forceinline__ void wait_arrivals(const unsigned* p, unsigned count) {
    while (ld_relaxed(p) < count) {}
    (void)ld_acquire(p);
}

// per-CTA constants kept in shared memory for the event path
struct TileCtx {
    CfgDev cfg;
    SeqDev sq;
};
static_assert(sizeof(TileCtx) <= 704, "TileCtx must fit its shared-memory slot");

struct CandEntry {  // shot candidate found in phase A: pixel (tile-local), its uniform, I and C
    int lidx;
    float u, I, C;
};

// Event-queue entry (one 32-bit word per flagged pixel-frame):
//   [12:0] pixel index inside the tile, [13] polarity is ON, [14] shot candidate, [31:15] count.
// After processing, the slot holds pol*final_evts_frame as a signed int; the pixel's owner turns it
// into the memory update (pol*final)*C (:522) with its own copy of the thresholds.
__device__ __forceinline__ unsigned q_pack(int lidx, bool up, bool cand, int cnt) {
    return (unsigned)lidx | ((unsigned)up << 13) | ((unsigned)cand << 14) |
           ((unsigned)min(cnt, 0x1ffff) << 15);
}

template <int FDT>
__device__ __forceinline__ float ring_px(const unsigned char* s_frame, int lidx) {
    if (FDT == V2E_DT_F32) return *(const float*)(s_frame + lidx * 4);
    if (FDT == V2E_DT_U8) return (float)s_frame[lidx];
    return __uint_as_float(((unsigned)*(const unsigned short*)(s_frame + lidx * 2)) << 16);
}

// shared-memory loads through a 32-bit shared-window address (saves the generic->shared address
// arithmetic the compiler otherwise redoes at every use)
__device__ __forceinline__ float2 lds_f2(unsigned addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 lds_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds_u32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2 lds_u2(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}

// Frame value -> (L, I) with the LUT addressed through its shared-window address.
__device__ __forceinline__ void log_inten_s(unsigned lut_addr, float x, float& L, float& I) {
    const int v = __float2int_rn(x);
    float2 e;
    if (__builtin_expect((float)v == x && (unsigned)v < 256u, 1)) e = lds_f2(lut_addr + 8u * (unsigned)v);
    else e = log_inten_slow(x);
    L = e.x;
    I = e.y;
}


// NOISE: 0 = no noise, 1 = PHILOX leak + shot (the V2E2VNet wiring), 2 = generic: every feature
// flag is read at run time, debug outputs and injected noise are honoured (tests, odd configs).
template <int FDT, int R, int NOISE, bool CUTOFF>
__global__ void __launch_bounds__(kPThreads, 1) k_persistent(const __grid_constant__ PersistParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr bool GENERIC = NOISE == 2;
    const CfgDev& cfg = prm.cfg;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    const SeqDev& sq = prm.seq[seq];
    SeqWork& wk = prm.work[seq];
    const int P = cfg.P;
    const int nb = cfg.nb;
    const int T = sq.T;
    const int tid = threadIdx.x;

    const bool leak = GENERIC ? (cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool shot = GENERIC ? (cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF) : (NOISE == 1);
    const bool philox = GENERIC ? (cfg.noise_mode == V2E_NOISE_PHILOX) : (NOISE == 1);
    const bool cutoff = GENERIC ? (cfg.cutoff_on != 0) : CUTOFF;
    long long* tl = GENERIC ? prm.timeline : nullptr;  // thread-0 time stamps (development aid)

    // ---- tile geometry -------------------------------------------------------------------
    const int total_groups = P >> 2;  // P % 4 == 0 is a precondition of this engine
    const int g0 = tile * prm.tile_groups;
    int ngroups = total_groups - g0;
    ngroups = ngroups > prm.tile_groups ? prm.tile_groups : (ngroups < 0 ? 0 : ngroups);
    const int tile_px0 = g0 * 4;
    const int tile_px = ngroups * 4;
    const int tile_px_alloc = prm.tile_groups * 4;

    // ---- shared memory carve-up ------------------------------------------------------------
    float2* s_lut = (float2*)smem_raw;                                   // 2 KB
    unsigned long long* s_bar = (unsigned long long*)(smem_raw + 2048);  // frame mbarrier
    int* s_red = (int*)(smem_raw + 2112);                                // one int per warp
    float* s_stat = (float*)(smem_raw + 2240);                           // mean, den, on
    int* s_qcount = (int*)(smem_raw + 2272);
    int* s_ncand = (int*)(smem_raw + 2276);
    int* s_mx = (int*)(smem_raw + 2280);
    TileCtx* s_ctx = (TileCtx*)(smem_raw + 2304);                        // <= 704 B
    double* s_dred = (double*)(smem_raw + 3072);                         // kWarps x 4
    CandEntry* s_cand = (CandEntry*)(smem_raw + 4096);                   // kCandCap x 16 B
    FrameSc* s_fr = (FrameSc*)(smem_raw + 5120);                         // V2E_MAX_FRAMES x 32 B
    unsigned char* s_ring = smem_raw + kHeaderBytes;
    const int stage_bytes = (tile_px_alloc * prm.frame_elem + 127) & ~127;
    float* s_tmem = (float*)(s_ring + (size_t)prm.stages * stage_bytes);
    float* s_lkn = s_tmem + tile_px_alloc;                               // leak_rate * noise_rate
    float* s_vox = s_lkn + tile_px_alloc;
    unsigned* s_q = (unsigned*)(s_vox + (size_t)nb * tile_px_alloc);     // tile_px entries
    const unsigned a_lut = smem_u32(s_lut), a_ring = smem_u32(s_ring), a_lkn = smem_u32(s_lkn);

    const bool use_ring = prm.stages > 0;
    const unsigned frame_bytes = (unsigned)(tile_px * prm.frame_elem);

    load_lut(s_lut, cfg.lut);
    for (int i = tid; i < T; i += kPThreads) s_fr[i] = prm.fr[seq][i];
    if (tid == 0) {
        mbar_init(&s_bar[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_ctx->cfg = cfg;
        s_ctx->sq = sq;
        *s_qcount = 0;
        *s_ncand = 0;
    }
    __syncthreads();
    // The frame tile of step n+1 is fetched by one TMA bulk copy while step n waits for the other
    // tiles (it does not depend on any state); the whole pack is first pulled into L2.
    if (use_ring && tid == 32 && ngroups > 0) {
        const unsigned char* src = (const unsigned char*)sq.frames + (size_t)tile_px0 * prm.frame_elem;
        for (int n = 2; n < T; ++n)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src + (size_t)n * P * prm.frame_elem),
                         "r"(frame_bytes) : "memory");
        mbar_expect_tx(&s_bar[0], frame_bytes);
        tma_load_1d(s_ring, src + (size_t)P * prm.frame_elem, frame_bytes, &s_bar[0]);
    }

    // ---- load per-pixel state into registers / shared memory --------------------------------
    float base[kPPT], lpv[kPPT], pos[kPPT], neg[kPPT];
    unsigned small_mask = 0;  // bit j: pixel j of this thread is a "small" (ps, qs) pixel
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int gl = g * kPThreads + tid;  // group index inside the tile
        const bool valid = gl < ngroups;
        const int p0 = tile_px0 + gl * 4;
        float4 b4 = make_float4(0, 0, 0, 0), l4 = b4, t4 = b4, ps4 = make_float4(1, 1, 1, 1),
               ng4 = ps4, nr4 = b4;
        if (valid) {
            ps4 = __ldg((const float4*)(sq.pos + p0));
            ng4 = __ldg((const float4*)(sq.neg + p0));
            if (leak) nr4 = __ldg((const float4*)(sq.nra + p0));
            if (sq.first_pack) {
                // _init (v2e_model.py:170-174,239-240): base = lp = lin_log(frame 0), tmem = -Tr
                const float4 x4 = load_frame4<FDT>(sq.frames, p0);
                float I;
                log_inten(s_lut, x4.x, b4.x, I);
                log_inten(s_lut, x4.y, b4.y, I);
                log_inten(s_lut, x4.z, b4.z, I);
                log_inten(s_lut, x4.w, b4.w, I);
                l4 = b4;
                const float t0 = fadd(0.f, -sq.Tr);
                t4 = make_float4(t0, t0, t0, t0);
            } else {
                b4 = *(const float4*)(sq.base + p0);
                if (cutoff) l4 = *(const float4*)(sq.lp + p0);
                t4 = *(const float4*)(sq.tmem + p0);
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
            *(float4*)(s_tmem + gl * 4) = t4;
            *(float4*)(s_lkn + gl * 4) =
                make_float4(fmul(cfg.leak_rate, nr4.x), fmul(cfg.leak_rate, nr4.y),
                            fmul(cfg.leak_rate, nr4.z), fmul(cfg.leak_rate, nr4.w));
            for (int b = 0; b < nb; ++b)
                *(float4*)(s_vox + b * tile_px + gl * 4) = make_float4(0, 0, 0, 0);
            const int y = p0 / cfg.W, x = p0 - y * cfg.W;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int xe = x + e, ye = y;
                if (xe >= cfg.W) { xe -= cfg.W; ye += 1; }   // a group may straddle a row end
                if (xe >= cfg.W) { xe -= cfg.W; ye += 1; }
                if (((xe | ye) & 1) == 0) small_mask |= 1u << (g * 4 + e);
            }
        }
        const float* bp = &b4.x; const float* lp_ = &l4.x; const float* pp = &ps4.x;
        const float* np_ = &ng4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = g * 4 + e;
            base[j] = bp[e]; lpv[j] = lp_[e]; pos[j] = pp[e]; neg[j] = np_[e];
        }
    }
    __syncthreads();

    long long nev = 0;
    const float jit = cfg.leak_jitter;
    const float shot_m1 = cfg.shot_inten_m1;
    // ---- time-step over the frames of the pack ---------------------------------------------
    for (int n = 1; n < T; ++n) {
        if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 0] = clock64();
        if (use_ring && ngroups > 0) mbar_wait(&s_bar[0], (n - 1) & 1);
        if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 1] = clock64();
        const FrameSc fs = s_fr[n];

        // ---- phase A: low-pass, leak, diff, count; flagged pixels go to the event queue ------
        int lmax = 0;
        unsigned flags = 0;      // bit j: pixel j was queued this frame
        int slot[kGroups];       // first queue slot of each group's entries
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const int gl = g * kPThreads + tid;
            slot[g] = 0;
            if (gl < ngroups) {
                const int p0 = tile_px0 + gl * 4;
                float4 x4;
                if (use_ring) {
                    if (FDT == V2E_DT_F32) x4 = lds_f4(a_ring + gl * 16);
                    else if (FDT == V2E_DT_U8) {
                        const unsigned v = lds_u32(a_ring + gl * 4);
                        x4 = make_float4((float)(v & 0xffu), (float)((v >> 8) & 0xffu),
                                         (float)((v >> 16) & 0xffu), (float)(v >> 24));
                    } else {
                        const uint2 v = lds_u2(a_ring + gl * 8);
                        x4 = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                                         __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
                    }
                } else {
                    x4 = load_frame4<FDT>(sq.frames, (long long)n * P + p0);
                }
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                float4 lk4 = make_float4(0, 0, 0, 0);
                unsigned shot_bytes = 0;
                if (leak) lk4 = lds_f4(a_lkn + gl * 16);
                if (philox) {
                    const uint4 w = noise_words<R>(sq.philox_ctr, n, (unsigned)(g0 + gl), 0);
                    shot_bytes = w.z;
                    if (leak) {
                        box_muller16(w.x, z[0], z[1]);
                        box_muller16(w.y, z[2], z[3]);
                        if (GENERIC && sq.dbg_leak_randn)
                            *(float4*)(sq.dbg_leak_randn + (size_t)(n - 1) * P + p0) =
                                make_float4(z[0], z[1], z[2], z[3]);
                    }
                } else if (leak) {
                    const float4 zz = __ldg((const float4*)(sq.leak_randn + (size_t)(n - 1) * P + p0));
                    z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
                }
                const float* xp = &x4.x;
                const float* lkp = &lk4.x;
                unsigned ent[4];
                unsigned gflags = 0;
                signed char polb[4];
                int cntb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    float L, I;
                    log_inten_s(a_lut, xp[e], L, I);
                    float nw = L;
                    if (cutoff) {
                        const float s = (small_mask >> j) & 1u ? fs.s_small : fs.s_large;
                        lpv[j] = lowpass(lpv[j], L, I, s);
                        nw = lpv[j];
                    }
                    if (leak) base[j] = leak_sub(base[j], lkp[e], jit, z[e], fs.sdt, pos[j]);
                    // diff / polarity / threshold (v2e_model.py:375,394-404), see diff_count()
                    const float diff = fsub(nw, base[j]);
                    const float ad = fabsf(diff);
                    const bool live = ad > 1e-6f;
                    const bool up = diff > 0.f;
                    const float C = live ? (up ? pos[j] : neg[j]) : 0.f;
                    const bool cross = live && ad >= fadd(C, 1e-9f);
                    bool maybe = cross;
                    float thr256 = 0.f;
                    unsigned byte = 0;
                    if (shot) {
                        if (philox) {
                            thr256 = shot_thr256(up ? fs.kp : fs.kn, shot_m1, I, C);   // 0 if !live
                            byte = (shot_bytes >> (8 * e)) & 0xffu;
                            maybe = maybe || (float)byte < thr256;
                        } else {
                            maybe = maybe || live;  // injected draws: every pol != 0 pixel iterates
                        }
                    }
                    ent[e] = 0;
                    if (GENERIC) { polb[e] = live ? (up ? 1 : -1) : 0; cntb[e] = 0; }
                    if (__builtin_expect(maybe, 0)) {
                        // rare path (~4 % of pixel-frames): exact count, shot second stage, queue
                        const int cnt = cross ? count_div(ad, fadd(C, 1e-9f)) : 0;
                        bool cand = false;
                        if (shot && live) {
                            if (philox) {
                                if ((float)byte < thr256) {
                                    const float u = shot_uniform<R>(sq.philox_ctr, n, p0 + e, byte);
                                    cand = u * 256.0f < thr256;
                                    if (cand) {
                                        const int k = atomicAdd(s_ncand, 1);
                                        if (k < kCandCap) s_cand[k] = CandEntry{gl * 4 + e, u, I, C};
                                    }
                                }
                            } else {
                                cand = true;
                            }
                        }
                        if (cnt > 0 || cand) {
                            gflags |= 1u << e;
                            lmax = max(lmax, cnt);
                            ent[e] = q_pack(gl * 4 + e, up, cand, cnt);
                        }
                        if (GENERIC) cntb[e] = cnt;
                    }
                }
                if (gflags) {
                    const int s0 = atomicAdd(s_qcount, __popc(gflags));
                    slot[g] = s0;
                    int k = s0;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((gflags >> e) & 1u) s_q[k++] = ent[e];
                    flags |= gflags << (4 * g);
                }
                if (GENERIC) {
                    if (sq.dbg_pol)
                        *(char4*)(sq.dbg_pol + (size_t)(n - 1) * P + p0) =
                            make_char4(polb[0], polb[1], polb[2], polb[3]);
                    if (sq.dbg_cnt)
                        *(int4*)(sq.dbg_cnt + (size_t)(n - 1) * P + p0) =
                            make_int4(cntb[0], cntb[1], cntb[2], cntb[3]);
                    if (sq.dbg_final)
                        *(int4*)(sq.dbg_final + (size_t)(n - 1) * P + p0) = make_int4(0, 0, 0, 0);
                }
            }
        }
        if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 2] = clock64();

        // ---- per-sequence max over all tiles (v2e_model.py:405-406) -----------------------
        lmax = __reduce_max_sync(0xffffffffu, lmax);
        if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
        __syncthreads();
        if (tid < 32) {
            if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 3] = clock64();
            int m = tid < kWarps ? s_red[tid] : 0;
            m = __reduce_max_sync(0xffffffffu, m);
            if (tid == 0) {
                const int mx = frame_barrier(&wk.bar[n], m, (unsigned)prm.tiles_per_seq);
                if (tl) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 5] = clock64();
                *s_mx = mx;
                if (tile == 0 && sq.max_iters) sq.max_iters[n] = mx;
            }
        } else if (tid == 32) {
            // every thread has consumed this frame's tile: fetch the next one (L2 -> smem) while
            // the barrier is in flight
            if (use_ring && ngroups > 0 && n + 1 < T) {
                mbar_expect_tx(&s_bar[0], frame_bytes);
                tma_load_1d(s_ring, (const unsigned char*)sq.frames +
                                        ((size_t)(n + 1) * P + tile_px0) * prm.frame_elem,
                            frame_bytes, &s_bar[0]);
            }
        }
        __syncthreads();
        if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 6] = clock64();

        // ---- phase B: the CTA drains the event queue, one entry per thread -------------------
        {
            const int qn = *s_qcount;
            // entry q goes to lane q / kWarps of warp q % kWarps: every warp gets ~qn/kWarps
            // entries, so no warp walks a long divergent iteration loop on the critical path
            const int q0 = (tid & 31) * kWarps + (tid >> 5);
            if (q0 < qn) {
                const FrameGlobals g = frame_globals(*s_mx, sq.duration, sq.Tr);
                const int ncand = min(*s_ncand, kCandCap);
                for (int q = q0; q < qn; q += kPThreads) {
                    const unsigned e = s_q[q];
                    const int lidx = (int)(e & 0x1fffu);
                    const bool up = (e >> 13) & 1u;
                    const bool cand = (e >> 14) & 1u;
                    const int cnt = (int)(e >> 15);
                    const int p = tile_px0 + lidx;
                    float I = 0.f, u = 0.f, C = 0.f;
                    int mode = 0;
                    if (cand && g.mx > 0) {   // rare: shot candidate (or injected shot draws)
                        mode = philox ? 2 : 1;
                        bool found = false;
                        if (philox)
                            for (int k = 0; k < ncand; ++k)
                                if (s_cand[k].lidx == lidx) {
                                    u = s_cand[k].u; I = s_cand[k].I; C = s_cand[k].C; found = true;
                                }
                        if (!found) {   // injected draws, or the candidate table overflowed
                            C = up ? __ldg(sq.pos + p) : __ldg(sq.neg + p);
                            float L;
                            log_inten(s_lut, load_frame1<FDT>(sq.frames, (long long)n * P + p), L, I);
                            if (philox) {   // redraw (same value as in phase A)
                                const uint4 w = noise_words<R>(sq.philox_ctr, n, (unsigned)(p >> 2), 0);
                                u = shot_uniform<R>(sq.philox_ctr, n, p, (w.z >> (8 * (p & 3))) & 0xffu);
                            }
                        }
                    }
                    float tm = s_tmem[lidx];
                    float* vp = s_vox + lidx;
                    const float polf = up ? 1.f : -1.f;
                    const int fin = iterate_pixel<R>(
                        s_ctx->cfg, s_ctx->sq, fs, g, n, p, polf, C, cnt, I, tm, mode, u,
                        [&](int b, float v) { vp[b * tile_px] = fadd(vp[b * tile_px], v); },
                        [&](float ts, float pl) {
                            const unsigned long long sl = atomicAdd(&wk.raw_count, 1ull);
                            if ((long long)sl < sq.raw_capacity) {
                                float* row = sq.raw_events + 5 * sl;
                                const int y = p / cfg.W;
                                row[0] = ts; row[1] = (float)(p - y * cfg.W); row[2] = (float)y;
                                row[3] = pl; row[4] = 0.f;
                            }
                        });
                    s_tmem[lidx] = tm;
                    s_q[q] = (unsigned)(up ? fin : -fin);   // pol*final, applied by the owner below
                    nev += fin;
                    if (GENERIC && fin > 0 && sq.dbg_final) sq.dbg_final[(size_t)(n - 1) * P + p] = fin;
                }
            }
        }
        if (tl && tid == 0) tl[((size_t)blockIdx.x * V2E_MAX_FRAMES + n) * 8 + 7] = clock64();
        __syncthreads();
        if (tid == 0) { *s_qcount = 0; *s_ncand = 0; }   // nobody reads them until the next frame
        if (flags) {
            // memory update base += (pol*final)*C (v2e_model.py:522)
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                int k = slot[g];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    if ((flags >> j) & 1u) {
                        const int pf = (int)s_q[k++];
                        if (pf != 0)
                            base[j] = fadd(base[j], fmul((float)pf, pf > 0 ? pos[j] : neg[j]));
                    }
                }
            }
        }
        __syncthreads();  // queue slots are reused by the next frame
    }

    // ---- write back the dynamic state ----------------------------------------------------------
    __syncthreads();
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int gl = g * kPThreads + tid;
        if (gl < ngroups) {
            const int p0 = tile_px0 + gl * 4;
            *(float4*)(sq.base + p0) = make_float4(base[g * 4], base[g * 4 + 1], base[g * 4 + 2], base[g * 4 + 3]);
            if (cutoff)
                *(float4*)(sq.lp + p0) = make_float4(lpv[g * 4], lpv[g * 4 + 1], lpv[g * 4 + 2], lpv[g * 4 + 3]);
            *(float4*)(sq.tmem + p0) = *(const float4*)(s_tmem + gl * 4);
        }
    }

    // ---- statistics of the un-normalised grid (utils/event_process.py:166-174) -----------------
    const bool voxel_out = cfg.output_mode != V2E_OUT_RAW;
    double s1 = 0., s2 = 0., nz = 0.;
    if (voxel_out && cfg.output_mode == V2E_OUT_VOXEL) {
        for (int b = 0; b < nb; ++b)
            for (int i = tid; i < tile_px; i += kPThreads) {
                const float v = s_vox[b * tile_px + i];
                s1 += (double)v;
                s2 += (double)v * (double)v;
                nz += v != 0.f ? 1. : 0.;
            }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nz += __shfl_down_sync(0xffffffffu, nz, o);
        nev += __shfl_down_sync(0xffffffffu, nev, o);
    }
    if ((tid & 31) == 0) {
        s_dred[(tid >> 5) * 4 + 0] = s1;
        s_dred[(tid >> 5) * 4 + 1] = s2;
        s_dred[(tid >> 5) * 4 + 2] = nz;
        s_dred[(tid >> 5) * 4 + 3] = (double)nev;
    }
    __syncthreads();
    const int fin_slot = V2E_MAX_FRAMES;  // barrier slot of the final (statistics) barrier
    if (tid < 32) {
        // warp 0: publish this tile's partial sums, wait for the other tiles, combine
        if (tid == 0) {
            double a = 0., b = 0., c = 0., d = 0.;
            for (int i = 0; i < kWarps; ++i) {
                a += s_dred[i * 4]; b += s_dred[i * 4 + 1]; c += s_dred[i * 4 + 2]; d += s_dred[i * 4 + 3];
            }
            wk.partials[3 * tile] = a; wk.partials[3 * tile + 1] = b; wk.partials[3 * tile + 2] = c;
            if (d > 0.) atomicAdd(&wk.nev, (unsigned long long)d);
            __threadfence();
            red_release_add((unsigned*)&wk.bar[fin_slot], 1u);
            wait_arrivals((const unsigned*)&wk.bar[fin_slot], (unsigned)prm.tiles_per_seq);
        }
        __syncwarp();
        // fixed assignment tile -> lane and a fixed shuffle tree: deterministic, and identical in
        // every CTA of the sequence
        double a = 0., b = 0., c = 0.;
        for (int i = tid; i < prm.tiles_per_seq; i += 32) {
            a += __ldcg(&wk.partials[3 * i]); b += __ldcg(&wk.partials[3 * i + 1]); c += __ldcg(&wk.partials[3 * i + 2]);
        }
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
            c += __shfl_xor_sync(0xffffffffu, c, o);
        }
        if (tid == 0) {
            float mean = 0.f, den = 1.f, on = 0.f;
            if (c > 0. && cfg.output_mode == V2E_OUT_VOXEL) {
                mean = fdiv((float)a, (float)c);
                den = fadd(sqrtf(fsub(fdiv((float)b, (float)c), fmul(mean, mean))), 1e-8f);
                on = 1.f;
            }
            s_stat[0] = mean; s_stat[1] = den; s_stat[2] = on;
            if (tile == 0) *sq.num_events = __ldcg(&wk.nev);
        }
    }
    __syncthreads();

    // ---- normalise and store the voxel tile (utils/event_process.py:175) ------------------------
    if (voxel_out) {
        const float mean = s_stat[0], den = s_stat[1];
        const bool norm = s_stat[2] != 0.f;
        for (int b = 0; b < nb; ++b) {
            for (int gl = tid; gl < ngroups; gl += kPThreads) {
                float4 v = *(const float4*)(s_vox + b * tile_px + gl * 4);
                if (norm) {
                    float* vp = &v.x;
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        vp[e] = fdiv(fmul(vp[e] != 0.f ? 1.f : 0.f, fsub(vp[e], mean)), den);
                }
                const size_t o = (size_t)b * P + tile_px0 + gl * 4;
                if (cfg.voxel_bf16) {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                    uint2 pk;
                    pk.x = *(unsigned*)&lo; pk.y = *(unsigned*)&hi;
                    *(uint2*)((__nv_bfloat16*)sq.voxel + o) = pk;
                } else {
                    __stcs((float4*)((float*)sq.voxel + o), v);
                }
            }
        }
    }

    // ---- leave the work area zeroed: the last CTA of the sequence to get here cleans up -------
    if (tid == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(&wk.exit_count, 1u);
        if (prev == (unsigned)prm.tiles_per_seq - 1) {
            for (int i = 0; i < V2E_MAX_FRAMES + 4; ++i) wk.bar[i] = 0ull;
            wk.nev = 0ull;
            wk.raw_count = 0ull;
            __threadfence();
            wk.exit_count = 0;
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct PersistPlan {
    int ok;
    int tiles_per_seq, tile_groups, stages;
    size_t smem;
};

static int g_sm_count = -1, g_smem_optin = 0, g_coop = 0;

static int device_props() {
    if (g_sm_count >= 0) return V2E_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return V2E_ERR_NO_DEVICE;
    int major = 0, sms = 0, optin = 0, coop = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (major < 10) return V2E_ERR_NO_DEVICE;
    g_smem_optin = optin;
    g_coop = coop;
    g_sm_count = sms;
    return V2E_OK;
}

static int frame_elem_bytes(int fdt) { return fdt == V2E_DT_U8 ? 1 : (fdt == V2E_DT_BF16 ? 2 : 4); }

static PersistPlan make_plan(const CfgDev& cfg, int frame_dtype, int n_seq) {
    PersistPlan pl{};
    if (device_props() != V2E_OK || !g_coop) return pl;
    if (cfg.P % 4 != 0) return pl;
    const int fe = frame_elem_bytes(frame_dtype);
    const int total_groups = cfg.P / 4;
    int tiles = g_sm_count / n_seq;  // CTAs (= SMs) available per sequence
    if (tiles < 1) return pl;
    // tile boundaries are multiples of 4 groups (16 px) so that every TMA source address is
    // 16 B aligned for any frame dtype whose plane size allows it
    int tg = (total_groups + tiles - 1) / tiles;
    tg = (tg + 3) & ~3;
    if (tg * 4 > kTileMaxPx) return pl;
    tiles = (total_groups + tg - 1) / tg;
    const bool tma_ok = ((size_t)cfg.P * fe) % 16 == 0;
    // header + timestamp_mem + leak plane + num_bins voxel planes + event queue (1 word / pixel)
    const size_t fixed = kHeaderBytes + (size_t)tg * 4 * sizeof(float) * (3 + cfg.nb);
    const size_t sb = ((size_t)tg * 4 * fe + 127) & ~(size_t)127;
    // one frame tile in flight: it is refilled (from L2) while the CTA sits in the frame barrier
    const int stages = tma_ok ? 1 : 0;
    const size_t smem = fixed + (size_t)stages * sb;
    if (smem > (size_t)g_smem_optin) return pl;
    pl.ok = 1;
    pl.tiles_per_seq = tiles;
    pl.tile_groups = tg;
    pl.stages = stages;
    pl.smem = smem;
    return pl;
}

int persistent_fits(const CfgDev& cfg, int frame_dtype, int n_seq) {
    const int rc = device_props();
    if (rc != V2E_OK) return rc;
    return make_plan(cfg, frame_dtype, n_seq).ok;
}

template <int FDT, int R, int NOISE, bool CUTOFF>
static int launch_persistent(const PersistParams& prm, const PersistPlan& pl, cudaStream_t st) {
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_persistent<FDT, R, NOISE, CUTOFF>;
    int rc = V2E_OK;
    if (!attr_set) {
        rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          g_smem_optin));
        if (rc != V2E_OK) return rc;
        attr_set = true;
    }
    void* args[] = {(void*)&prm};
    const dim3 grid((unsigned)(pl.tiles_per_seq * prm.n_seq));
    rc = cuda_ok(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(kPThreads), args,
                                             pl.smem, st));
    count_launch();
    return rc;
}

template <int FDT>
static int run_persistent_t(const CfgDev& cfg, int rounds, const v2e_pack* packs, int n_seq,
                            void* workspace, cudaStream_t st) {
    const PersistPlan pl = make_plan(cfg, FDT, n_seq);
    if (!pl.ok) return V2E_ERR_UNSUPPORTED;
    static thread_local PersistParams prm;  // 16 KB: keep it off the stack
    prm.cfg = cfg;
    prm.n_seq = n_seq;
    prm.tiles_per_seq = pl.tiles_per_seq;
    prm.tile_groups = pl.tile_groups;
    prm.stages = pl.stages;
    prm.frame_elem = frame_elem_bytes(FDT);
    prm.work = (SeqWork*)workspace;
    static const bool want_timeline = getenv("V2E_TIMELINE") != nullptr;
    // the stepwise engine's accumulator area of the workspace is unused here: borrow it
    prm.timeline = want_timeline ? (long long*)((char*)workspace + step_layout(cfg, n_seq).vox_off)
                                 : nullptr;
    bool dbg = want_timeline;
    for (int s = 0; s < n_seq; ++s) {
        fill_seq(prm.seq[s], packs[s]);
        for (int n = 1; n < packs[s].sc.num_frames; ++n) {
            const v2e_frame_scalars& f = packs[s].sc.fr[n];
            prm.fr[s][n] = make_frame_sc(cfg, f);
        }
        const v2e_pack& p = packs[s];
        dbg = dbg || p.dbg_pol || p.dbg_cnt || p.dbg_final || p.dbg_leak_randn || p.dbg_shot;
    }
    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool philox = cfg.noise_mode == V2E_NOISE_PHILOX;
    // fast variants: no debug outputs, voxel output, and either no noise or PHILOX leak+shot
    int noise = 2;
    if (!dbg && cfg.output_mode != V2E_OUT_RAW) {
        if (!leak && !shot) noise = 0;
        else if (philox && leak && shot) noise = 1;
    }
    const bool r7 = rounds == 7;
    if (noise == 0)
        return cfg.cutoff_on ? launch_persistent<FDT, 10, 0, true>(prm, pl, st)
                             : launch_persistent<FDT, 10, 0, false>(prm, pl, st);
    if (noise == 1) {
        if (cfg.cutoff_on)
            return r7 ? launch_persistent<FDT, 7, 1, true>(prm, pl, st)
                      : launch_persistent<FDT, 10, 1, true>(prm, pl, st);
        return r7 ? launch_persistent<FDT, 7, 1, false>(prm, pl, st)
                  : launch_persistent<FDT, 10, 1, false>(prm, pl, st);
    }
    return r7 ? launch_persistent<FDT, 7, 2, false>(prm, pl, st)
              : launch_persistent<FDT, 10, 2, false>(prm, pl, st);
}

int run_persistent(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs,
                   int n_seq, void* workspace, cudaStream_t st) {
    switch (frame_dtype) {
        case V2E_DT_U8: return run_persistent_t<V2E_DT_U8>(cfg, rounds, packs, n_seq, workspace, st);
        case V2E_DT_BF16: return run_persistent_t<V2E_DT_BF16>(cfg, rounds, packs, n_seq, workspace, st);
        default: return run_persistent_t<V2E_DT_F32>(cfg, rounds, packs, n_seq, workspace, st);
    }
}

}  // namespace v2e
