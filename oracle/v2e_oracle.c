/*
 * oracle/v2e_oracle.c — CPU restatement of the V2E2V emulator + voxel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker the CUDA path is diffed against; nothing
 * under v2e2v_b200/ may import, link or call it.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it.
 *
 * Parity status: PINNED.  The reference ships no tests or golden vectors, so the
 * oracle is pinned against outputs of the reference itself: tests/golden/make_golden.py imports
 * the unmodified reference from /root/reference, runs it on seeded inputs with injected noise
 * and stores states / voxel grids / event rows in tests/golden (npz files); tests/test_oracle_golden.py
 * requires this file to reproduce them (state planes bit-exact, voxel grid rtol 1e-5).
 *
 * Every function cites the reference lines it restates (paths relative to the reference repo).
 * All arithmetic is IEEE f32 with one rounding per operation — build with -ffp-contract=off
 * (see oracle/Makefile); f64 only where torch's type promotion makes the reference use f64.
 *
 * All pointers in v2e_pack / v2e_config are HOST pointers here.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/v2e2v_b200.h"

#ifdef _OPENMP
#include <omp.h>
#endif

/* lin_log, v2e/emulator_utils.py:13-37: fp64 piecewise lin/log, rounded to 1e-8, cast to f32. */
static float oracle_lin_log(double x) {
    const double f = (1. / 20.) * log(20.);
    double y = (x <= 20.) ? x * f : log(x);
    const double rounding = 1e8;
    y = nearbyint(y * rounding) / rounding; /* torch.round = half-to-even */
    return (float)y;
}

/* rescale_intensity_frame, v2e/emulator_utils.py:40-45: (x + 20) / 275 in f32. */
static float oracle_inten01(float x) { return (x + 20.f) / 275.f; }

/* Build the 2x256 LUT the product path consumes (exact for integer-valued frames). */
void v2e_oracle_build_lut(float* lut /* [2][256] */) {
    for (int v = 0; v < 256; ++v) {
        lut[v] = oracle_lin_log((double)v);
        lut[256 + v] = oracle_inten01((float)v);
    }
}

static inline float bf16_to_f32(uint16_t h) {
    uint32_t u = ((uint32_t)h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

static inline uint16_t f32_to_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* nan */
    uint32_t lsb = (u >> 16) & 1u;
    u += 0x7fffu + lsb;
    return (uint16_t)(u >> 16);
}

static inline float load_frame(const v2e_config* cfg, const void* frames, int64_t idx) {
    switch (cfg->frame_dtype) {
        case V2E_DT_U8: return (float)((const uint8_t*)frames)[idx];
        case V2E_DT_BF16: return bf16_to_f32(((const uint16_t*)frames)[idx]);
        default: return ((const float*)frames)[idx];
    }
}

static int cmp_rows(const void* a, const void* b) {
    const float* r = (const float*)a;
    const float* q = (const float*)b;
    if (r[0] != q[0]) return r[0] < q[0] ? -1 : 1;
    if (r[2] != q[2]) return r[2] < q[2] ? -1 : 1;
    if (r[1] != q[1]) return r[1] < q[1] ? -1 : 1;
    return 0;
}

int v2e_oracle_voxel_normalize(float* v, int64_t n, int mode, float hot_pixel_thresh);

/*
 * One pack of one sequence (B = 1): EventEmulator.forward, v2e/v2e_model.py:290-536, with the
 * host scalar prep (:303-318) taken from pk->sc.  Per-pixel restatement.
 * Returns 0 or a V2E_ERR_* code.  `threads` <= 0 -> all available (OpenMP), 1 -> serial.
 */
int v2e_oracle_pack(const v2e_config* cfg, const v2e_pack* pk, int threads) {
    if (!cfg || !pk) return V2E_ERR_BAD_ARG;
    const int H = cfg->height, W = cfg->width, nb = cfg->num_bins;
    const int T = pk->sc.num_frames;
    const int64_t P = (int64_t)H * W;
    if (T < 2 || T > V2E_MAX_FRAMES || nb < 1 || nb > V2E_MAX_BINS) return V2E_ERR_BAD_ARG;
    const int leak = cfg->leak_on && cfg->noise_mode != V2E_NOISE_OFF;
    const int shot = cfg->shot_on && cfg->noise_mode != V2E_NOISE_OFF;
    if ((leak || shot) && cfg->noise_mode != V2E_NOISE_INJECTED) return V2E_ERR_UNSUPPORTED;
    if (leak && !pk->leak_randn) return V2E_ERR_BAD_ARG;
    if (shot && !pk->shot_rand) return V2E_ERR_BAD_ARG;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#else
    (void)threads;
#endif
    float* base = pk->base_log_frame;
    float* lp = pk->lp_log_frame0;
    float* tmem = pk->timestamp_mem;
    const float Tr = pk->sc.Tr;
    const int raw = cfg->output_mode == V2E_OUT_RAW;

    float* vox = NULL;
    if (!raw) {
        vox = (float*)calloc((size_t)nb * P, sizeof(float));
        if (!vox) return V2E_ERR_BAD_ARG;
    }
    int32_t* cnt = (int32_t*)malloc((size_t)P * sizeof(int32_t));
    float* dif = (float*)malloc((size_t)P * sizeof(float));
    int64_t nev = 0, nraw = 0;
    int rc = V2E_OK;

    /* first pack: _init (v2e_model.py:156-174,239-240); else rollover (:329-330) */
    if (pk->sc.first_pack) {
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < P; ++p) {
            const float x0 = load_frame(cfg, pk->frames, p);
            const float l0 = oracle_lin_log((double)x0);
            base[p] = l0;
            if (cfg->cutoff_on) lp[p] = l0;
            tmem[p] = 0.f + (-Tr);
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t p = 0; p < P; ++p) {
            float tm = tmem[p];
            if (tm > 0.f) tm = tm - pk->sc.bins_m1;
            if (tm < 0.f) tm = -Tr;
            tmem[p] = tm;
        }
    }

    for (int n = 1; n < T && rc == V2E_OK; ++n) {
        const v2e_frame_scalars* fs = &pk->sc.fr[n];
        int32_t mx = 0;
        /* ---- phase A, per pixel: lin-log, low-pass, leak, diff, polarity, count ------------ */
#pragma omp parallel for schedule(static) reduction(max : mx)
        for (int64_t p = 0; p < P; ++p) {
            const float x = load_frame(cfg, pk->frames, (int64_t)n * P + p);
            const float L = oracle_lin_log((double)x);
            float nw = L;
            if (cfg->cutoff_on) {
                /* low_pass_filter, emulator_utils.py:73-93 */
                const float I = oracle_inten01(x);
                const int y = (int)(p / W), xx = (int)(p % W);
                const int small = ((y & 1) == 0) && ((xx & 1) == 0);
                const float s = small ? fs->s_small : fs->s_large;
                float eps = isinf(s) ? 1.f : I * s;
                if (eps > 1.f) eps = 1.f; /* torch.clamp(max=1) */
                const float a = (1.f - eps) * lp[p];
                const float b = eps * L;
                lp[p] = a + b;
                nw = lp[p];
            }
            if (leak) {
                /* subtract_leak_current, emulator_utils.py:113-125 */
                const float z = pk->leak_randn[(int64_t)(n - 1) * P + p];
                const float clr =
                    (cfg->leak_rate * pk->noise_rate[p]) * (1.f - cfg->leak_jitter * z);
                const float dl = (fs->sdt * clr) * pk->pos_thres[p];
                base[p] = base[p] - dl;
            }
            /* v2e_model.py:375,394-404 */
            float diff = nw - base[p];
            if (!(fabsf(diff) > 1e-6f)) diff = 0.f;
            const float C = diff > 0.f ? pk->pos_thres[p] : (diff < 0.f ? pk->neg_thres[p] : 0.f);
            const int32_t c = (int32_t)floorf(fabsf(diff) / (C + 1e-9f));
            cnt[p] = c;
            dif[p] = diff;
            if (c > mx) mx = c;
        }
        if (pk->max_iters) pk->max_iters[n] = mx;
        /* ---- global: num_iters, ts_step, gate (v2e_model.py:405-417,469) ------------------ */
        const int32_t ni = mx > 0 ? mx : 1;
        const float ts_step = (1.f / (float)ni) * pk->sc.duration;
        const int gate = Tr > ts_step;
        /* shot coefficient (rate/2)*dt/num_iters, emulator_utils.py:174; dtype per trap T5 */
        const int f64 = (pk->sc.shot_f64_mask >> n) & 1u;
        const double a64 = fs->shot_c / (double)ni;
        const float a32 = (float)fs->shot_c / (float)ni;
        if (shot && mx > pk->shot_imax) { rc = V2E_ERR_BAD_ARG; break; }
        /* ---- phase B, per pixel: iteration loop (v2e_model.py:449-522) -------------------- */
        int64_t nev_f = 0;
        /* raw rows are appended serially (omp if(!raw)); order is canonicalised by the sort */
#pragma omp parallel for schedule(static) reduction(+ : nev_f) if (!raw)
        for (int64_t p = 0; p < P; ++p) {
            const float diff = dif[p];
            const float pol = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
            const float C = diff > 0.f ? pk->pos_thres[p] : (diff < 0.f ? pk->neg_thres[p] : 0.f);
            const int32_t c = cnt[p];
            int32_t final = 0;
            float tm = tmem[p];
            /* shot thresholds, emulator_utils.py:173-188 */
            double on64 = 2., off64 = -1.;
            float on32 = 2.f, off32 = -1.f;
            if (shot && mx > 0) {
                const float x = load_frame(cfg, pk->frames, (int64_t)n * P + p);
                const float fac = cfg->shot_inten_m1 * oracle_inten01(x) + 1.f;
                const float ppp = cfg->rn_pos * pk->pos_thres[p];
                const float npp = cfg->rn_neg * pk->neg_thres[p];
                if (f64) {
                    const double snf = a64 * (double)fac;
                    on64 = 1. - snf * (double)ppp;
                    off64 = snf * (double)npp;
                } else {
                    const float snf = a32 * fac;
                    on32 = 1.f - snf * ppp;
                    off32 = snf * npp;
                }
            }
            for (int32_t i = 0; i < mx; ++i) {
                int m = c >= i + 1;
                if (shot) {
                    const float r = pk->shot_rand[((int64_t)(n - 1) * pk->shot_imax + i) * P + p];
                    int on, off;
                    if (f64) { on = (double)r > on64; off = (double)r < off64; }
                    else { on = r > on32; off = r < off32; }
                    m = m || (on && pol > 0.f) || (off && pol < 0.f);
                }
                const float ts = fs->t_bin_prev + ts_step * (float)(i + 1);
                if (gate) {
                    /* v2e_model.py:469-472 */
                    const float since = (m ? ts : 0.f) - tm;
                    m = since > Tr;
                    if (m) tm = ts;
                }
                if (!m) continue;
                final += 1;
                nev_f += 1;
                if (!raw) {
                    /* v2e_model.py:478-502 */
                    const float ti = floorf(ts);
                    const float d = ts - ti;
                    const int b = (int)ti;
                    if (b >= 0 && b < nb) vox[(int64_t)b * P + p] += pol * (1.0f - d);
                    if (b >= 0 && b + 1 < nb) vox[(int64_t)(b + 1) * P + p] += pol * d;
                } else {
                    /* v2e_model.py:504-518: rows [t, x, y, pol, b] */
                    if (nraw < pk->raw_capacity) {
                        float* row = pk->raw_events + 5 * nraw;
                        row[0] = ts;
                        row[1] = (float)(p % W);
                        row[2] = (float)(p / W);
                        row[3] = pol;
                        row[4] = 0.f;
                    }
                    nraw++;
                }
            }
            tmem[p] = tm;
            if (pk->dbg_pol) pk->dbg_pol[(int64_t)(n - 1) * P + p] = (int8_t)pol;
            if (pk->dbg_cnt) pk->dbg_cnt[(int64_t)(n - 1) * P + p] = c;
            if (pk->dbg_final) pk->dbg_final[(int64_t)(n - 1) * P + p] = final;
            /* v2e_model.py:522 */
            base[p] = base[p] + (pol * (float)final) * C;
        }
        nev += nev_f;
    }

    if (rc == V2E_OK && raw && nraw > 1) {
        /* v2e_model.py:527-534 sorts by t (unstable) then by batch; the canonical order here is
         * (t, y, x), one of the orders the reference may produce */
        const int64_t m = nraw < pk->raw_capacity ? nraw : pk->raw_capacity;
        qsort(pk->raw_events, (size_t)m, 5 * sizeof(float), cmp_rows);
    }
    if (rc == V2E_OK && !raw) {
        const int64_t N = (int64_t)nb * P;
        /* event_preprocess_pytorch(mode='std', filter_hot_pixel=False), v2e_model.py:526 */
        if (cfg->output_mode == V2E_OUT_VOXEL) v2e_oracle_voxel_normalize(vox, N, 0, 0.f);
        if (cfg->voxel_dtype == V2E_DT_BF16) {
            uint16_t* o = (uint16_t*)pk->voxel;
            for (int64_t k = 0; k < N; ++k) o[k] = f32_to_bf16_rn(vox[k]);
        } else {
            memcpy(pk->voxel, vox, (size_t)N * sizeof(float));
        }
    }
    if (rc == V2E_OK && pk->num_events) *pk->num_events = nev;
    free(cnt); free(dif); free(vox);
    return rc;
}

/*
 * events_to_voxel_grid, utils/event_process.py:15-63 (numpy; torch_semantics = 0) and
 * events_to_voxel_grid_pytorch, :66-129 (torch_semantics = 1).  events: host [n,4] rows
 * (t,x,y,p) as f64 (events_f64) or f32; NOT mutated (the reference overwrites column 0 and maps
 * p == 0 to -1 in place).  voxel: host f32 [num_bins,H,W].  Accumulation is sequential in f32 - every
 * left-bin add in event order, then every right-bin add - like the reference's two np.add.at /
 * index_add_ calls on CPU.
 */
int v2e_oracle_events_to_voxel(const void* events, int events_f64, int64_t n, int num_bins,
                               int width, int height, int torch_semantics, float* voxel) {
    if (num_bins <= 0 || width <= 0 || height <= 0 || !voxel) return V2E_ERR_BAD_ARG;
    const int64_t P = (int64_t)width * height;
    memset(voxel, 0, (size_t)num_bins * P * sizeof(float));
    if (n == 0) return V2E_OK;
    const double* e64 = (const double*)events;
    const float* e32 = (const float*)events;
    /* deltaT and the rescaled stamps in the events' dtype (:33-40 / :92-98) */
    const double first64 = events_f64 ? e64[0] : 0, last64 = events_f64 ? e64[4 * (n - 1)] : 0;
    double dT64 = last64 - first64;
    if (dT64 == 0) dT64 = 1.0;
    const float first32 = events_f64 ? 0 : e32[0], last32 = events_f64 ? 0 : e32[4 * (n - 1)];
    float dT32 = last32 - first32;
    if (dT32 == 0) dT32 = 1.0f;
    /* every left-bin add in event order, then every right-bin add (:51-57 / :111-124) */
    for (int side = 0; side < 2; ++side)
        for (int64_t k = 0; k < n; ++k) {
            double ts, x, y, pol;
            if (events_f64) {
                ts = (double)(num_bins - 1) * (e64[4 * k] - first64) / dT64;
                x = e64[4 * k + 1]; y = e64[4 * k + 2]; pol = e64[4 * k + 3];
            } else {
                ts = (double)((float)(num_bins - 1) * (e32[4 * k] - first32) / dT32);   /* f32 ops */
                x = e32[4 * k + 1]; y = e32[4 * k + 2]; pol = e32[4 * k + 3];
            }
            if (pol == 0) pol = -1;
            const int64_t px = (int64_t)x + (int64_t)y * width;
            if (!torch_semantics) {
                /* numpy (:41-57): tis = ts.astype(uint) truncates; dts = ts - tis is float64 even
                 * for f32 events (float32 - uint64 promotes); the values are float64 and
                 * np.add.at adds them to the f32 grid in the (f32,f64)->f64 loop, rounding the
                 * sum to f32 on the way back */
                const double tf = trunc(ts);
                const int64_t ti = (int64_t)tf;
                const double d = ts - tf;
                const double left = pol * (1.0 - d), right = pol * d;
                if (ti < 0) continue;
                if (side == 0 && ti < num_bins) {
                    float* v = voxel + px + ti * P;
                    *v = (float)((double)*v + left);
                }
                if (side == 1 && ti + 1 < num_bins) {
                    float* v = voxel + px + (ti + 1) * P;
                    *v = (float)((double)*v + right);
                }
            } else {
                /* torch (:99-124): tis = floor(ts), dts = ts - tis in the events' dtype;
                 * pols.float() * (1.0 - dts.float()) and pols.float() * dts.float() in f32;
                 * index_add_ (CPU) adds in f32 in event order; `tis >= 0` guard */
                const double tf = floor(ts);
                const int64_t ti = (int64_t)tf;
                const float d = events_f64 ? (float)(ts - tf) : (float)ts - (float)tf;
                const float pf = (float)pol;
                const float left = pf * (1.0f - d), right = pf * d;
                if (ti < 0) continue;
                if (side == 0 && ti < num_bins) voxel[px + ti * P] += left;
                if (side == 1 && ti + 1 < num_bins) voxel[px + (ti + 1) * P] += right;
            }
        }
    return V2E_OK;
}

/*
 * event_preprocess / event_preprocess_pytorch, utils/event_process.py:132-154 / 157-177.
 * In place on host f32 [n].  mode 0 'std', 1 'maxmin'; hot_pixel_thresh > 0 zeroes |v| > thresh.
 * Sums are accumulated in f64 (the reference's f32 reduction order is unspecified).
 */
int v2e_oracle_voxel_normalize(float* v, int64_t n, int mode, float hot_pixel_thresh) {
    if (!v || n < 0) return V2E_ERR_BAD_ARG;
    if (hot_pixel_thresh > 0.f)
        for (int64_t k = 0; k < n; ++k)
            if (fabsf(v[k]) > hot_pixel_thresh) v[k] = 0.f;
    if (mode == 1) {
        if (n == 0) return V2E_OK;
        float mn = v[0], mxv = v[0];
        for (int64_t k = 1; k < n; ++k) {
            if (v[k] < mn) mn = v[k];
            if (v[k] > mxv) mxv = v[k];
        }
        const float den = (mxv - mn) + 1e-8f;
        for (int64_t k = 0; k < n; ++k) v[k] = (v[k] - mn) / den;
        return V2E_OK;
    }
    double s1 = 0., s2 = 0.;
    int64_t nnz = 0;
    for (int64_t k = 0; k < n; ++k) {
        if (v[k] != 0.f) nnz++;
        s1 += (double)v[k];
        s2 += (double)v[k] * (double)v[k];
    }
    if (nnz > 0) {
        const float mean = (float)s1 / (float)nnz;
        const float sd = sqrtf((float)s2 / (float)nnz - mean * mean);
        const float den = sd + 1e-8f;
        for (int64_t k = 0; k < n; ++k) {
            const float mk = v[k] != 0.f ? 1.f : 0.f;
            v[k] = (mk * (v[k] - mean)) / den;
        }
    }
    return V2E_OK;
}
