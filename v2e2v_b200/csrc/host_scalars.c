/*
 * host_scalars.c — per-pack scalar preparation on the host (no CUDA).
 *
 * Replaces the scalar part of EventEmulator.forward (v2e/v2e_model.py:303-318,335,353,521) and
 * the scalar factors of low_pass_filter / subtract_leak_current / generate_shot_noise
 * (v2e/emulator_utils.py:73-86,120,174).  torch evaluates those expressions on 0-dim tensors,
 * so which of them run in f32 and which in f64 depends on the dtype of the time stamps
 * (SURVEY.md traps T5, T6, T9); this file reproduces that promotion with plain IEEE arithmetic.
 * Compile with -ffp-contract=off: every operation must round separately.
 *
 * The canonical configuration (1000 fps, 10-frame packs, 1 ms refractory) sits on a rounding
 * knife-edge (Tr vs duration/num_iters), so "close" is not good enough here;
 * tests/test_host_scalars.py checks bit-equality with the torch expressions.
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "../../include/v2e2v_b200.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

int v2e_host_pack_scalars(const v2e_host_params* hp, const double* t_frames,
                          const float* time_frames, v2e_pack_scalars* out) {
    if (!hp || !t_frames || !time_frames || !out) return V2E_ERR_BAD_ARG;
    const int T = hp->num_frames, nb = hp->num_bins;
    if (T < 2 || T > V2E_MAX_FRAMES || nb < 1 || nb > V2E_MAX_BINS || hp->n_stamps < T)
        return V2E_ERR_BAD_ARG;
    memset(out, 0, sizeof(*out));

    /* Tr = einsum('i,ij->ij', (nb-1)*refr_f32, 1/(t_end - t_start)).float()  (:317)
     * (nb-1)*refr is an f32 product; the reciprocal is taken in the stamps' dtype. */
    const float a = (float)(nb - 1) * hp->refractory_period_s;
    if (hp->t_range_f64) {
        const double r = 1.0 / (hp->t_end - hp->t_start);
        out->Tr = (float)((double)a * r);
    } else {
        const float r = 1.0f / ((float)hp->t_end - (float)hp->t_start);
        out->Tr = a * r;
    }
    /* duration = (nb-1)/(T-1) is a python double; ts_step = rcp(f32(num_iters)) * f32(duration)
     * (:313,417 with Tensor.__rtruediv__) */
    out->duration = (float)((double)(nb - 1) / (double)(T - 1));
    out->bins_m1 = (float)(nb - 1);
    out->first_pack = hp->first_pack;
    out->num_frames = T;

    /* t_previous (:327,521): value + dtype */
    double tprev = hp->first_pack ? hp->t_start : hp->t_prev;
    int tprev_f64 = hp->first_pack ? hp->t_range_f64 : hp->t_prev_f64;
    if (!tprev_f64) tprev = (double)(float)tprev;

    /* time-order check (:335): t_float_frames[1] <= t_previous, compared after promotion */
    {
        const double t1 = hp->stamps_f64 ? t_frames[1] : (double)(float)t_frames[1];
        if (t1 <= tprev) return V2E_ERR_TIME_ORDER;
    }

    const int cutoff = hp->cutoff_hz > 0;
    /* tau = 1/(math.pi*2*cutoff_hz*q)  (emulator_utils.py:74,85), python doubles */
    const double tau_l = (cutoff && hp->ql > 0) ? 1.0 / (M_PI * 2 * hp->cutoff_hz * hp->ql) : 0.0;
    const double tau_s = (cutoff && hp->qs > 0) ? 1.0 / (M_PI * 2 * hp->cutoff_hz * hp->qs) : 0.0;
    const double half_rate = hp->shot_noise_rate_hz / 2;

    uint32_t mask = 0;
    for (int n = 1; n < T; ++n) {
        v2e_frame_scalars* fr = &out->fr[n];
        fr->t_bin_prev = time_frames[n - 1];
        /* IIR step: delta = t[n]-t[n-1] in the stamps' dtype; delta/tau in that dtype; rounded
         * to f32 when multiplied into the f32 intensity plane (:259, emulator_utils.py:75,86) */
        if (hp->stamps_f64) {
            const double d = t_frames[n] - t_frames[n - 1];
            fr->s_large = (cutoff && hp->ql > 0) ? (float)(d / tau_l) : INFINITY;
            fr->s_small = (cutoff && hp->qs > 0) ? (float)(d / tau_s) : INFINITY;
        } else {
            const float d = (float)t_frames[n] - (float)t_frames[n - 1];
            fr->s_large = (cutoff && hp->ql > 0) ? d / (float)tau_l : INFINITY;
            fr->s_small = (cutoff && hp->qs > 0) ? d / (float)tau_s : INFINITY;
        }
        /* delta_time = t[n] - t_previous (:353): f64 if either side is f64 */
        const int dt_f64 = hp->stamps_f64 || tprev_f64;
        double dtm;
        if (dt_f64) {
            const double tn = hp->stamps_f64 ? t_frames[n] : (double)(float)t_frames[n];
            dtm = tn - tprev;
        } else {
            dtm = (double)((float)t_frames[n] - (float)tprev);
        }
        fr->sdt = (float)dtm; /* exact when dtm is already f32 */
        /* (rate/2)*delta_time (emulator_utils.py:174): python double times a 0-dim tensor keeps
         * the tensor's dtype */
        if (dt_f64) {
            fr->shot_c = half_rate * dtm;
            mask |= 1u << n;
        } else {
            fr->shot_c = (double)((float)half_rate * (float)dtm);
        }
        /* t_previous = t_float_frames[n] (:521) */
        tprev = hp->stamps_f64 ? t_frames[n] : (double)(float)t_frames[n];
        tprev_f64 = hp->stamps_f64;
    }
    out->shot_f64_mask = mask;
    return V2E_OK;
}

/* K consecutive packs of one sequence prepared by one host call: what the caller's pack loop
 * (test.py:115-121) does K times around v2e_host_pack_scalars - t_previous carried from pack to
 * pack (v2e_model.py:521), first_pack only for the first - plus the per-pack pointers. */
int v2e_host_clip_prepare(const v2e_host_params* hp0, const double* const* stamp_rows,
                          size_t row_offset_bytes, const float* time_frames, const v2e_pack* tmpl,
                          const void* const* frames, size_t frame_offset_bytes,
                          size_t voxel_stride_bytes, size_t meta_stride_bytes, int n_packs,
                          int out_stride, double* t_prev_out, int32_t* failed_at, v2e_pack* out) {
    if (!hp0 || !stamp_rows || !time_frames || !tmpl || !frames || !out || n_packs < 1 ||
        out_stride < 1)
        return V2E_ERR_BAD_ARG;
    v2e_host_params hp = *hp0;
    hp.stamps_f64 = 1;
    hp.t_range_f64 = 1;
    const int T = hp.num_frames;
    if (T < 2 || T > V2E_MAX_FRAMES || hp.n_stamps < T) return V2E_ERR_BAD_ARG;
    for (int i = 0; i < n_packs; ++i) {
        if (!stamp_rows[i] || !frames[i]) return V2E_ERR_BAD_ARG;
        const double* row = (const double*)((const char*)stamp_rows[i] + row_offset_bytes);
        hp.t_start = row[0];
        hp.t_end = row[hp.n_stamps - 1];
        v2e_pack* p = &out[(size_t)i * out_stride];
        /* (everything but the scalars, which the call below fills) */
        memcpy(p, tmpl, offsetof(v2e_pack, sc));
        p->frames = (const char*)frames[i] + frame_offset_bytes;
        p->voxel = (char*)tmpl->voxel + (size_t)i * voxel_stride_bytes;
        p->num_events = (int64_t*)((char*)tmpl->num_events + (size_t)i * meta_stride_bytes);
        p->status = tmpl->status ? (int32_t*)((char*)tmpl->status + (size_t)i * meta_stride_bytes) : 0;
        p->philox_offset = tmpl->philox_offset + (uint64_t)i;
        const int rc = v2e_host_pack_scalars(&hp, row, time_frames, &p->sc);
        if (rc != V2E_OK) {
            if (failed_at) *failed_at = i;
            return rc;
        }
        hp.first_pack = 0;
        hp.t_prev = row[T - 1];
        hp.t_prev_f64 = 1;
    }
    if (t_prev_out) *t_prev_out = hp.t_prev;
    return V2E_OK;
}
