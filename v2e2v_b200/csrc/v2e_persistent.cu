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
#include <cooperative_groups.h>

#include "v2e_engine.cuh"

namespace v2e {

constexpr int kPThreads = 512;  // one CTA per SM
constexpr int kPPT = 16;        // pixels per thread (4 groups of 4)
constexpr int kGroups = kPPT / 4;
constexpr int kTileMaxPx = kPThreads * kPPT;
constexpr int kStagesMax = 3;
constexpr int kHeaderBytes = 4096;  // LUT, mbarriers, reduction scratch, TileCtx

struct PersistParams {
    CfgDev cfg;
    int n_seq;
    int tiles_per_seq;
    int tile_groups;   // 4-pixel groups per tile (last tile of a sequence may hold fewer)
    int stages;        // frame ring depth (0: frames are read straight from global memory)
    int frame_elem;    // bytes per frame element
    SeqWork* work;
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
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// per-CTA constants kept in shared memory for the out-of-line event path
struct TileCtx {
    CfgDev cfg;
    SeqDev sq;
    FrameSc fs;
    FrameGlobals g;
    int n;
    int tile_px0;      // first pixel of the tile (global pixel index)
    int tile_px;       // pixels in this tile
    float* s_tmem;
    float* s_vox;
    unsigned long long* raw_count;
};

static_assert(sizeof(TileCtx) <= 704, "TileCtx must fit its shared-memory slot");

// Out-of-line event path for one pixel (3 % of pixel-frames): iteration loop on smem state.
template <int R>
__device__ __noinline__ int tile_iterate(const TileCtx* c, int lp_idx, float diff, int cnt, float I,
                                         float pos, float neg, int shot_mode, unsigned word) {
    const int p = c->tile_px0 + lp_idx;
    float tm = c->s_tmem[lp_idx];
    float* vox = c->s_vox;
    const int tile_px = c->tile_px;
    const int fin = iterate_pixel<R>(
        c->cfg, c->sq, c->fs, c->g, c->n, p, diff, cnt, I, pos, neg, tm, shot_mode, word,
        [&](int b, float v) { vox[b * tile_px + lp_idx] = fadd(vox[b * tile_px + lp_idx], v); },
        [&](float ts, float polf) {
            const unsigned long long slot = atomicAdd(c->raw_count, 1ull);
            if ((long long)slot < c->sq.raw_capacity) {
                float* row = c->sq.raw_events + 5 * slot;
                const int y = p / c->cfg.W;
                row[0] = ts; row[1] = (float)(p - y * c->cfg.W); row[2] = (float)y;
                row[3] = polf; row[4] = 0.f;
            }
        });
    c->s_tmem[lp_idx] = tm;
    return fin;
}

template <int FDT, int R>
__global__ void __launch_bounds__(kPThreads, 1) k_persistent(const __grid_constant__ PersistParams prm) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const CfgDev& cfg = prm.cfg;
    const int seq = blockIdx.x / prm.tiles_per_seq;
    const int tile = blockIdx.x - seq * prm.tiles_per_seq;
    const SeqDev& sq = prm.seq[seq];
    SeqWork& wk = prm.work[seq];
    const int P = cfg.P;
    const int nb = cfg.nb;
    const int T = sq.T;
    const int tid = threadIdx.x;

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
    unsigned long long* s_bar = (unsigned long long*)(smem_raw + 2048);  // kStagesMax mbarriers
    int* s_red = (int*)(smem_raw + 2048 + 64);                           // 16 warps
    int* s_bcast = s_red + 32;
    TileCtx* s_ctx = (TileCtx*)(smem_raw + 2048 + 64 + 256);             // <= 704 B
    double* s_dred = (double*)(smem_raw + 3072);                         // 16 warps x 4
    unsigned char* s_ring = smem_raw + kHeaderBytes;
    const int stage_bytes = (tile_px_alloc * prm.frame_elem + 127) & ~127;
    float* s_tmem = (float*)(s_ring + (size_t)prm.stages * stage_bytes);
    float* s_vox = s_tmem + tile_px_alloc;

    const bool leak = cfg.leak_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool shot = cfg.shot_on && cfg.noise_mode != V2E_NOISE_OFF;
    const bool philox = cfg.noise_mode == V2E_NOISE_PHILOX;
    const bool use_ring = prm.stages > 0;
    const unsigned frame_bytes = (unsigned)(tile_px * prm.frame_elem);

    load_lut(s_lut, cfg.lut);
    if (tid == 0) {
        for (int s = 0; s < kStagesMax; ++s) mbar_init(&s_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_ctx->cfg = cfg;
        s_ctx->sq = sq;
        s_ctx->tile_px0 = tile_px0;
        s_ctx->tile_px = tile_px;
        s_ctx->s_tmem = s_tmem;
        s_ctx->s_vox = s_vox;
        s_ctx->raw_count = &wk.raw_count;
    }
    __syncthreads();
    // kick off the first frames of the pack (they do not depend on any state)
    if (use_ring && tid == 0 && ngroups > 0) {
        for (int n = 1; n < T && n <= prm.stages; ++n) {
            const int st = (n - 1) % prm.stages;
            mbar_expect_tx(&s_bar[st], frame_bytes);
            tma_load_1d(s_ring + (size_t)st * stage_bytes,
                        (const unsigned char*)sq.frames + ((size_t)n * P + tile_px0) * prm.frame_elem,
                        frame_bytes, &s_bar[st]);
        }
    }

    // ---- load per-pixel state into registers / shared memory --------------------------------
    float base[kPPT], lpv[kPPT], pos[kPPT], neg[kPPT], lkn[kPPT];
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
                if (cfg.cutoff_on) l4 = *(const float4*)(sq.lp + p0);
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
            for (int b = 0; b < nb; ++b)
                *(float4*)(s_vox + b * tile_px + gl * 4) = make_float4(0, 0, 0, 0);
        }
        const float* bp = &b4.x; const float* lp_ = &l4.x; const float* pp = &ps4.x;
        const float* np_ = &ng4.x; const float* rp = &nr4.x;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = g * 4 + e;
            base[j] = bp[e]; lpv[j] = lp_[e]; pos[j] = pp[e]; neg[j] = np_[e];
            lkn[j] = fmul(cfg.leak_rate, rp[e]);
            if (valid && is_small_px(p0 + e, cfg.W)) small_mask |= 1u << j;
        }
    }
    __syncthreads();

    long long nev = 0;
    // ---- time-step over the frames of the pack ---------------------------------------------
    for (int n = 1; n < T; ++n) {
        const FrameSc& fs = prm.fr[seq][n];
        const int st = use_ring ? (n - 1) % prm.stages : 0;
        const unsigned char* s_frame = s_ring + (size_t)st * stage_bytes;
        if (use_ring && ngroups > 0) mbar_wait(&s_bar[st], ((n - 1) / prm.stages) & 1);

        // ---- phase A: low-pass, leak, diff, count; flags for the event path ---------------
        int lmax = 0;
        unsigned evmask = 0, candmask = 0;
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            const int gl = g * kPThreads + tid;
            if (gl < ngroups) {
                const int p0 = tile_px0 + gl * 4;
                float4 x4;
                if (use_ring) {
                    if (FDT == V2E_DT_F32) x4 = *(const float4*)(s_frame + gl * 16);
                    else if (FDT == V2E_DT_U8) {
                        const uchar4 v = *(const uchar4*)(s_frame + gl * 4);
                        x4 = make_float4((float)v.x, (float)v.y, (float)v.z, (float)v.w);
                    } else {
                        const uint2 v = *(const uint2*)(s_frame + gl * 8);
                        x4 = make_float4(__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u),
                                         __uint_as_float(v.y << 16), __uint_as_float(v.y & 0xffff0000u));
                    }
                } else {
                    x4 = load_frame4<FDT>(sq.frames, (long long)n * P + p0);
                }
                float z[4] = {0.f, 0.f, 0.f, 0.f};
                uint4 w_shot = make_uint4(0, 0, 0, 0);
                if (leak) {
                    if (philox) {
                        const uint4 w = noise_words<R>(sq.philox_seed, sq.philox_offset, n,
                                                       (unsigned)(g0 + gl), 0x80);
                        box_muller16(w.x, z[0], z[1]);
                        box_muller16(w.y, z[2], z[3]);
                        if (sq.dbg_leak_randn)
                            *(float4*)(sq.dbg_leak_randn + (size_t)(n - 1) * P + p0) =
                                make_float4(z[0], z[1], z[2], z[3]);
                    } else {
                        const float4 zz = __ldg((const float4*)(sq.leak_randn + (size_t)(n - 1) * P + p0));
                        z[0] = zz.x; z[1] = zz.y; z[2] = zz.z; z[3] = zz.w;
                    }
                }
                if (shot && philox)
                    w_shot = noise_words<R>(sq.philox_seed, sq.philox_offset, n, (unsigned)(g0 + gl), 0);
                const unsigned ws[4] = {w_shot.x, w_shot.y, w_shot.z, w_shot.w};
                const float* xp = &x4.x;
                signed char polb[4];
                int cntb[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = g * 4 + e;
                    float L, I;
                    log_inten(s_lut, xp[e], L, I);
                    float nw = L;
                    if (cfg.cutoff_on) {
                        const float s = (small_mask >> j) & 1u ? fs.s_small : fs.s_large;
                        lpv[j] = lowpass(lpv[j], L, I, s);
                        nw = lpv[j];
                    }
                    if (leak) base[j] = leak_sub(base[j], lkn[j], cfg.leak_jitter, z[e], fs.sdt, pos[j]);
                    const PixelA a = diff_count(nw, base[j], pos[j], neg[j]);
                    const int pol = a.diff > 0.f ? 1 : (a.diff < 0.f ? -1 : 0);
                    lmax = max(lmax, a.cnt);
                    if (a.cnt > 0) evmask |= 1u << j;
                    if (shot && pol != 0) {
                        if (philox) {
                            const float bound = shot_any_bound(cfg, (float)fs.shot_c, I,
                                                               pol > 0 ? pos[j] : neg[j],
                                                               pol > 0 ? cfg.rn_pos : cfg.rn_neg);
                            if (u32_to_unit(ws[e]) < bound) candmask |= 1u << j;
                        } else {
                            candmask |= 1u << j;  // injected draws: every pol != 0 pixel iterates
                        }
                    }
                    polb[e] = (signed char)pol;
                    cntb[e] = a.cnt;
                }
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

        // ---- per-sequence max over all tiles (v2e_model.py:405-406) -----------------------
        lmax = __reduce_max_sync(0xffffffffu, lmax);
        if ((tid & 31) == 0) s_red[tid >> 5] = lmax;
        __syncthreads();
        if (tid == 0) {
            int m = 0;
#pragma unroll
            for (int i = 0; i < kPThreads / 32; ++i) m = max(m, s_red[i]);
            if (m > 0) atomicMax(&wk.gmax[n], m);
            red_release_add(&wk.arrive[n], 1u);
            while (ld_acquire(&wk.arrive[n]) < (unsigned)prm.tiles_per_seq) {}
            const int mx = *(volatile int*)&wk.gmax[n];
            s_ctx->fs = fs;
            s_ctx->n = n;
            s_ctx->g = frame_globals(mx, sq.duration, sq.Tr, fs.shot_c, (sq.shot_f64_mask >> n) & 1u);
            if (tile == 0 && sq.max_iters) sq.max_iters[n] = mx;
            // every thread has consumed the previous frame's stage (it passed the barrier
            // above after its phase B): refill it with the frame `stages` ahead
            if (use_ring && ngroups > 0 && n >= 2) {
                const int nn = n - 1 + prm.stages;
                if (nn < T) {
                    const int st2 = (nn - 1) % prm.stages;
                    mbar_expect_tx(&s_bar[st2], frame_bytes);
                    tma_load_1d(s_ring + (size_t)st2 * stage_bytes,
                                (const unsigned char*)sq.frames + ((size_t)nn * P + tile_px0) * prm.frame_elem,
                                frame_bytes, &s_bar[st2]);
                }
            }
        }
        __syncthreads();

        // ---- phase B: event path for flagged pixels (v2e_model.py:449-522) -----------------
        const int mx = s_ctx->g.mx;
        unsigned todo = evmask | (mx > 0 ? candmask : 0u);
        if (todo) {
#pragma unroll
            for (int j = 0; j < kPPT; ++j) {
                if ((todo >> j) & 1u) {
                    const int g = j >> 2, e = j & 3;
                    const int gl = g * kPThreads + tid;
                    const int lidx = gl * 4 + e;
                    // recompute this pixel's diff / count (cheaper than keeping 16 more registers)
                    float x;
                    if (use_ring) {
                        if (FDT == V2E_DT_F32) x = *(const float*)(s_frame + lidx * 4);
                        else if (FDT == V2E_DT_U8) x = (float)s_frame[lidx];
                        else x = __uint_as_float(((unsigned)*(const unsigned short*)(s_frame + lidx * 2)) << 16);
                    } else {
                        x = load_frame1<FDT>(sq.frames, (long long)n * P + tile_px0 + lidx);
                    }
                    float L, I;
                    log_inten(s_lut, x, L, I);
                    const float nw = cfg.cutoff_on ? lpv[j] : L;
                    const PixelA a = diff_count(nw, base[j], pos[j], neg[j]);
                    const bool cand = (candmask >> j) & 1u;
                    int mode = 0;
                    unsigned word = 0;
                    if (cand && mx > 0) {
                        mode = philox ? 2 : 1;
                        if (philox) {
                            const uint4 w = noise_words<R>(sq.philox_seed, sq.philox_offset, n,
                                                           (unsigned)(g0 + gl), 0);
                            const unsigned ws[4] = {w.x, w.y, w.z, w.w};
                            word = ws[e];
                        }
                    }
                    const int fin = tile_iterate<R>(s_ctx, lidx, a.diff, a.cnt, I, pos[j], neg[j], mode, word);
                    if (fin > 0) {
                        const float C = a.diff > 0.f ? pos[j] : neg[j];
                        const float polf = a.diff > 0.f ? 1.f : -1.f;
                        base[j] = fadd(base[j], fmul(fmul(polf, (float)fin), C));  // :522
                        nev += fin;
                        if (sq.dbg_final) sq.dbg_final[(size_t)(n - 1) * P + tile_px0 + lidx] = fin;
                    }
                }
            }
        }
    }

    // ---- write back the dynamic state ----------------------------------------------------------
    __syncthreads();
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const int gl = g * kPThreads + tid;
        if (gl < ngroups) {
            const int p0 = tile_px0 + gl * 4;
            *(float4*)(sq.base + p0) = make_float4(base[g * 4], base[g * 4 + 1], base[g * 4 + 2], base[g * 4 + 3]);
            if (cfg.cutoff_on)
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
    if (tid == 0) {
        double a = 0., b = 0., c = 0., d = 0.;
        for (int i = 0; i < kPThreads / 32; ++i) {
            a += s_dred[i * 4]; b += s_dred[i * 4 + 1]; c += s_dred[i * 4 + 2]; d += s_dred[i * 4 + 3];
        }
        wk.partials[3 * tile] = a; wk.partials[3 * tile + 1] = b; wk.partials[3 * tile + 2] = c;
        if (d > 0.) atomicAdd(&wk.nev, (unsigned long long)d);
        __threadfence();
        red_release_add(&wk.arrive[fin_slot], 1u);
        while (ld_acquire(&wk.arrive[fin_slot]) < (unsigned)prm.tiles_per_seq) {}
        // fixed-order combine: deterministic, identical in every CTA of the sequence
        a = b = c = 0.;
        for (int i = 0; i < prm.tiles_per_seq; ++i) {
            a += __ldcg(&wk.partials[3 * i]); b += __ldcg(&wk.partials[3 * i + 1]); c += __ldcg(&wk.partials[3 * i + 2]);
        }
        float mean = 0.f, den = 1.f, on = 0.f;
        if (c > 0. && cfg.output_mode == V2E_OUT_VOXEL) {
            mean = fdiv((float)a, (float)c);
            den = fadd(sqrtf(fsub(fdiv((float)b, (float)c), fmul(mean, mean))), 1e-8f);
            on = 1.f;
        }
        float* s_stat = (float*)s_bcast;
        s_stat[0] = mean; s_stat[1] = den; s_stat[2] = on;
        if (tile == 0) {
            const unsigned long long total = *(volatile unsigned long long*)&wk.nev;
            *sq.num_events = total;
        }
    }
    __syncthreads();

    // ---- normalise and store the voxel tile (utils/event_process.py:175) ------------------------
    if (voxel_out) {
        const float* s_stat = (const float*)s_bcast;
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
            for (int i = 0; i < V2E_MAX_FRAMES; ++i) wk.gmax[i] = 0;
            for (int i = 0; i < V2E_MAX_FRAMES + 4; ++i) wk.arrive[i] = 0;
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
    const size_t fixed = kHeaderBytes + (size_t)tg * 4 * sizeof(float) * (1 + cfg.nb);
    int stages = 0;
    if (tma_ok) {
        const size_t sb = ((size_t)tg * 4 * fe + 127) & ~(size_t)127;
        for (int s = kStagesMax; s >= 1; --s)
            if (fixed + s * sb <= (size_t)g_smem_optin) { stages = s; break; }
    }
    if (stages == 1) stages = 0;  // the ring needs one frame of lead: 2+ stages or none
    const size_t sb = ((size_t)tg * 4 * fe + 127) & ~(size_t)127;
    size_t smem = fixed + (size_t)stages * sb;
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

template <int FDT, int R>
static int run_persistent_t(const CfgDev& cfg, const v2e_pack* packs, int n_seq, void* workspace,
                            cudaStream_t st) {
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
    for (int s = 0; s < n_seq; ++s) {
        fill_seq(prm.seq[s], packs[s]);
        for (int n = 1; n < packs[s].sc.num_frames; ++n) {
            const v2e_frame_scalars& f = packs[s].sc.fr[n];
            prm.fr[s][n] = FrameSc{f.t_bin_prev, f.s_large, f.s_small, f.sdt, f.shot_c};
        }
    }
    static bool attr_set = false;  // one process drives one device (multi-GPU)
    auto kern = k_persistent<FDT, R>;
    int rc = V2E_OK;
    if (!attr_set) {
        rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          g_smem_optin));
        if (rc != V2E_OK) return rc;
        attr_set = true;
    }
    void* args[] = {(void*)&prm};
    const dim3 grid((unsigned)(pl.tiles_per_seq * n_seq));
    rc = cuda_ok(cudaLaunchCooperativeKernel((const void*)kern, grid, dim3(kPThreads), args, pl.smem, st));
    count_launch();
    return rc;
}

int run_persistent(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs,
                   int n_seq, void* workspace, cudaStream_t st) {
#define V2E_DISPATCH(FDT)                                                               \
    return rounds == 7 ? run_persistent_t<FDT, 7>(cfg, packs, n_seq, workspace, st)     \
                       : run_persistent_t<FDT, 10>(cfg, packs, n_seq, workspace, st)
    switch (frame_dtype) {
        case V2E_DT_U8: V2E_DISPATCH(V2E_DT_U8);
        case V2E_DT_BF16: V2E_DISPATCH(V2E_DT_BF16);
        default: V2E_DISPATCH(V2E_DT_F32);
    }
#undef V2E_DISPATCH
}

}  // namespace v2e
