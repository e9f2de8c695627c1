/*
 * v2e2v_b200.h — C ABI of the B200-native V2E (video -> events) emulator + voxel-grid path.
 *
 * This is the drop-in boundary for the data-parallel hot path of lsying009/V2E2V.  Every entry
 * point below replaces a piece of the reference's pure-PyTorch implementation; the reference has
 * no FFI of its own (it is ~10 000 ATen calls per pack), so the binding a maintainer adds is the
 * ctypes stub shown in INTEGRATION.md (v2e2v_b200/_cabi.py is that stub, productised).
 *
 *   reference interface (file:line, relative to the reference repo)      entry point here
 *   ------------------------------------------------------------------   -----------------------
 *   EventEmulator.forward            v2e/v2e_model.py:290-536             v2e_pack_forward
 *     lin_log                        v2e/emulator_utils.py:13-37            (fused, LUT + fp64 path)
 *     rescale_intensity_frame        v2e/emulator_utils.py:40-45            (fused)
 *     IIR_temporal_filtering         v2e/v2e_model.py:252-275               (fused)
 *     low_pass_filter                v2e/emulator_utils.py:48-102           (fused)
 *     subtract_leak_current          v2e/emulator_utils.py:105-125          (fused)
 *     generate_shot_noise            v2e/emulator_utils.py:161-207          (fused)
 *     event_preprocess_pytorch       utils/event_process.py:157-177         (fused epilogue)
 *   forward() scalar prep            v2e/v2e_model.py:303-318,353           v2e_host_pack_scalars
 *   the caller's pack loop           test.py:115-121, model_v2e2v.py:101    v2e_host_clip_prepare (host half),
 *                                                                          v2e_clip_forward, v2e_batch_clip_forward
 *   events_to_voxel_grid[_pytorch]   utils/event_process.py:15-63,66-129  v2e_events_to_voxel
 *   event_preprocess[_pytorch]       utils/event_process.py:132-177       v2e_voxel_normalize
 *
 * Conventions
 *   - plain C, no torch types; every pointer marked "device" is a CUDA device pointer owned by
 *     the caller (PyTorch owns all memory; the library allocates nothing persistent).
 *   - all launches are asynchronous on the cudaStream_t passed as `stream` (a void* here so the
 *     header has no CUDA dependency); nothing synchronises the device.
 *   - return value: 0 on success, a negative V2E_ERR_* code otherwise; never throws.
 *   - one "sequence" == one reference EventEmulator instance with batch_size 1 (the only
 *     well-defined batch semantics of the reference, see DESIGN.md "B>1").
 */
#ifndef V2E2V_B200_H
#define V2E2V_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2E_ABI_VERSION 2

#define V2E_MAX_FRAMES 32   /* max frames per pack (num_pack_frames); reference default is 10 */
#define V2E_MAX_BINS 8      /* max voxel bins; reference default is 5                        */
#define V2E_MAX_SEQ_PER_LAUNCH 16

/* error codes */
#define V2E_OK 0
#define V2E_ERR_BAD_ARG (-1)        /* null pointer / bad dims / misaligned pointer             */
#define V2E_ERR_UNSUPPORTED (-2)    /* valid request this build cannot serve                    */
#define V2E_ERR_TIME_ORDER (-3)     /* frame time not later than previous (ValueError upstream) */
#define V2E_ERR_CUDA (-4)           /* a CUDA runtime call failed; see v2e_last_cuda_error      */
#define V2E_ERR_WORKSPACE (-5)      /* workspace too small                                      */
#define V2E_ERR_NO_DEVICE (-6)      /* no sm_100 device                                         */

/* v2e_pack.status values written by the device */
#define V2E_STATUS_OK 0
#define V2E_STATUS_NONINTEGER_FRAMES 1 /* frames_integer was asserted but a value was not 0..255
                                          integer: the outputs of this pack are invalid          */
#define V2E_STATUS_NOT_CONVERGED 2     /* fixed-point engine: internal error (the chain of per-
                                          frame maxima did not settle): outputs are invalid      */

/* noise_mode */
#define V2E_NOISE_OFF 0       /* leak_rate_hz == 0 and shot_noise_rate_hz == 0                   */
#define V2E_NOISE_INJECTED 1  /* caller supplies the randn / rand draws the reference would make */
#define V2E_NOISE_PHILOX 2    /* in-kernel counter-based RNG (Philox4x32)                        */

/* frame_dtype / voxel_dtype */
#define V2E_DT_F32 0
#define V2E_DT_U8 1
#define V2E_DT_BF16 2

/* output_mode */
#define V2E_OUT_VOXEL 0       /* normalised voxel grid  (reference output_mode 'voxel_grid')     */
#define V2E_OUT_RAW 1         /* raw event rows [t,x,y,pol,b] (reference output_mode 'raw')      */
#define V2E_OUT_VOXEL_UNNORM 2/* voxel grid before event_preprocess_pytorch (debug / parity)     */

/* engine selection for v2e_pack_forward */
#define V2E_ENGINE_AUTO 0
#define V2E_ENGINE_STEPWISE 1   /* two launches per frame, state in HBM/L2; any size             */
#define V2E_ENGINE_PERSISTENT 2 /* one cooperative launch per pack, state on chip, lock-step     */
#define V2E_ENGINE_RUNAHEAD 3   /* one cooperative launch per pack, streaming decoupled from the
                                   per-frame max: only the pixels that fire wait for it          */
#define V2E_ENGINE_FIXPOINT 4   /* one cooperative launch per pack, the per-frame maxima are
                                   guessed and verified: pixels never wait for each other        */

/* Per-emulator static configuration (v2e/v2e_model.py:36-153). */
typedef struct v2e_config {
    int32_t abi_version;     /* V2E_ABI_VERSION                                                  */
    int32_t height, width;
    int32_t num_bins;        /* <= V2E_MAX_BINS                                                  */
    int32_t cutoff_on;       /* cutoff_hz > 0  (v2e_model.py:255)                                */
    int32_t leak_on;         /* leak_rate_hz > 0 (v2e_model.py:362)                              */
    int32_t shot_on;         /* shot_noise_rate_hz > 0 (v2e_model.py:438)                        */
    int32_t noise_mode;      /* V2E_NOISE_*                                                      */
    int32_t frame_dtype;     /* V2E_DT_F32 | V2E_DT_U8 | V2E_DT_BF16                             */
    int32_t voxel_dtype;     /* V2E_DT_F32 | V2E_DT_BF16                                         */
    int32_t output_mode;     /* V2E_OUT_*                                                        */
    int32_t engine;          /* V2E_ENGINE_*                                                     */
    float leak_rate;         /* f32(leak_rate_hz)          emulator_utils.py:118                 */
    float leak_jitter;       /* f32(leak_jitter_fraction)  emulator_utils.py:118                 */
    float rn_pos;            /* f32(1/pos_thres_nominal)   v2e_model.py:217                      */
    float rn_neg;            /* f32(1/neg_thres_nominal)   v2e_model.py:218                      */
    float shot_inten_m1;     /* f32(SHOT_NOISE_INTEN_FACTOR-1) emulator_utils.py:175             */
    int32_t philox_rounds;   /* 7 or 10 (0 -> default 10)                                        */
    int32_t frames_integer;  /* 1: the caller guarantees integer-valued frames in 0..255 (what
                                every reader of the reference produces; always true for
                                V2E_DT_U8).  Lets the run-ahead engine drop the fp64 lin_log path
                                from its streaming loop; a violation is detected and reported
                                through v2e_pack.status, never served silently.               */
    int32_t reserved_cfg;
    const float* lut;        /* device, [2][256]: lin_log(v), (v+20)/275 for v = 0..255          */
} v2e_config;

/* Host-prepared per-frame scalars; index n pairs frame n with stamp n (v2e_model.py:350-353). */
typedef struct v2e_frame_scalars {
    float t_bin_prev;  /* time_frames[n-1]                   v2e_model.py:314,419               */
    float s_large;     /* f32(dts/tau(ql)); +inf if ql<=0    emulator_utils.py:73-77            */
    float s_small;     /* f32(dts/tau(qs)); +inf if qs<=0    emulator_utils.py:84-89            */
    float sdt;         /* f32(t[n] - t_previous)             emulator_utils.py:120              */
    double shot_c;     /* (rate/2)*(t[n]-t_previous) in the dtype torch evaluates it in         */
} v2e_frame_scalars;

typedef struct v2e_pack_scalars {
    float Tr;              /* refractory period in bin units  v2e_model.py:317                   */
    float duration;        /* f32((num_bins-1)/(T-1))         v2e_model.py:313,417               */
    float bins_m1;         /* f32(num_bins-1)                 v2e_model.py:329                   */
    int32_t first_pack;    /* 1: first call after reset()     v2e_model.py:325-327               */
    uint32_t shot_f64_mask;/* bit n set: frame n's shot thresholds are evaluated in fp64         */
    int32_t num_frames;    /* T = frames.size(1)                                                 */
    int32_t reserved0, reserved1;
    v2e_frame_scalars fr[V2E_MAX_FRAMES];
} v2e_pack_scalars;

/* One pack of one sequence.  P = height*width.  Planes are contiguous f32[P]. */
typedef struct v2e_pack {
    const void* frames;       /* device [T,H,W] of frame_dtype (frame 0 read only on first pack) */
    float* base_log_frame;    /* device in/out   v2e_model.py:170,522                            */
    float* lp_log_frame0;     /* device in/out   v2e_model.py:174,264  (unused if !cutoff_on)    */
    float* timestamp_mem;     /* device in/out   v2e_model.py:240,329-330,472                    */
    const float* pos_thres;   /* device          v2e_model.py:183-198                            */
    const float* neg_thres;   /* device          v2e_model.py:200-214                            */
    const float* noise_rate;  /* device          v2e_model.py:232-236  (unused if !leak_on)      */
    void* voxel;              /* device out [num_bins,H,W] of voxel_dtype (voxel modes)          */
    float* raw_events;        /* device out [raw_capacity,5] rows t,x,y,pol,b (raw mode)         */
    int64_t raw_capacity;
    int64_t* num_events;      /* device out, scalar (also = rows written in raw mode)            */
    int32_t* status;          /* device out, optional scalar: 0 ok, V2E_STATUS_NONINTEGER_FRAMES  */
    int32_t* max_iters;       /* device out [T] optional: per-frame max event count (n>=1)       */
    const float* leak_randn;  /* device [T-1,P]  injected randn (emulator_utils.py:113)          */
    const float* shot_rand;   /* device [T-1,shot_imax,P] injected rand (emulator_utils.py:191)  */
    int32_t shot_imax;
    int32_t reserved;
    int8_t* dbg_pol;          /* device out [T-1,P] optional: polarity map                       */
    int32_t* dbg_cnt;         /* device out [T-1,P] optional: event_counts  (v2e_model.py:404)   */
    int32_t* dbg_final;       /* device out [T-1,P] optional: final_evts_frame (v2e_model.py:475)*/
    float* dbg_leak_randn;    /* device out [T-1,P] optional: normals drawn in PHILOX mode       */
    uint8_t* dbg_shot;        /* device out [T-1,dbg_shot_imax,P] optional: shot decisions       */
    int32_t dbg_shot_imax;
    int32_t reserved2;
    uint64_t philox_seed;
    uint64_t philox_offset;   /* pack counter; distinct per pack                                 */
    v2e_pack_scalars sc;
} v2e_pack;

/* -- queries -------------------------------------------------------------------------------- */
int v2e_abi_version(void);
/* Human-readable string for a V2E_ERR_* code. */
const char* v2e_error_string(int code);
/* Last cudaError_t (as int) recorded by this thread's most recent V2E_ERR_CUDA, and its name. */
int v2e_last_cuda_error(void);
const char* v2e_last_cuda_error_string(void);
/* Bytes of zero-initialised device workspace v2e_pack_forward needs for n_seq sequences. */
size_t v2e_workspace_bytes(const v2e_config* cfg, int n_seq);
/* Which engine AUTO resolves to for this config on the current device (V2E_ENGINE_*), or <0. */
int v2e_resolve_engine(const v2e_config* cfg, int n_seq);
/* Kernels launched by this process through this library so far (for bench.py gpu_launches). */
int64_t v2e_launch_count(void);

/* -- host-side scalar preparation (v2e_model.py:303-318,353; emulator_utils.py:73-86,174) ----
 * Pure host code (no CUDA).  t_frames: the stamps torch would index as t_float_frames[n]
 * (n < n_stamps, n_stamps >= num_frames), already resolved for the 2-column case by the caller.
 * stamps_f64: dtype of t_float_frames (1: float64, 0: float32, passed widened to double).
 * t_prev / t_prev_f64: value and dtype of self.t_previous (ignored when first_pack, where it is
 * t_start/t_start_f64 = t_frames[0,0] in the caller's original dtype).
 * t_start/t_end (+dtype flag): t_frames[0,0], t_frames[0,-1] used for Tr.
 * time_frames: f32[num_frames] = torch.linspace(0, duration*(T-1), T) computed by the caller.
 * Returns V2E_ERR_TIME_ORDER when t_float_frames[1] <= t_previous (v2e_model.py:335). */
typedef struct v2e_host_params {
    int32_t num_frames, num_bins;
    int32_t n_stamps;
    int32_t stamps_f64;
    int32_t first_pack;
    int32_t t_prev_f64;
    int32_t t_range_f64;
    int32_t reserved;
    double t_prev;
    double t_start, t_end;
    double cutoff_hz, ql, qs;
    double shot_noise_rate_hz;
    float refractory_period_s;
    float reserved_f;
} v2e_host_params;
int v2e_host_pack_scalars(const v2e_host_params* hp, const double* t_frames,
                          const float* time_frames, v2e_pack_scalars* out);

/* K consecutive packs of ONE sequence prepared by one host call (pure host code): the scalars
 * of each pack as v2e_host_pack_scalars gives them with t_previous carried from pack to pack
 * (v2e_model.py:521; hp0->first_pack / t_prev / t_prev_f64 describe the state before pack 0),
 * plus the per-pack pointers.  The caller's loop `for pack in clip` (test.py:115-121) costs tens
 * of microseconds of interpreter time per pack, as much as a third of a kernel; this is the same
 * work in well under a microsecond.  Serves the usual stamp layout only: every row holds
 * hp0->n_stamps >= num_frames float64 stamps (stamps_f64 / t_range_f64 / t_start / t_end of hp0
 * are ignored and derived per row).  For pack i:
 *   row        = (double*)((char*)stamp_rows[i] + row_offset_bytes)
 *   o          = out[i * out_stride]   (out_stride > 1: the sequence is one of a batch whose
 *                                       step records are interleaved, v2e_batch_clip_forward)
 *   o          = *tmpl, then
 *   o.frames   = (char*)frames[i] + frame_offset_bytes
 *   o.voxel    = tmpl->voxel + i * voxel_stride_bytes
 *   o.num_events / o.status = tmpl's + i * meta_stride_bytes
 *   o.philox_offset = tmpl->philox_offset + i,   o.sc = the pack's scalars.
 * On V2E_ERR_TIME_ORDER (or any other error) *failed_at is the pack index, nothing else is
 * reported and *t_prev_out is not written.  On success *t_prev_out = the last pack's last stamp
 * (a float64 t_previous). */
int v2e_host_clip_prepare(const v2e_host_params* hp0, const double* const* stamp_rows,
                          size_t row_offset_bytes, const float* time_frames, const v2e_pack* tmpl,
                          const void* const* frames, size_t frame_offset_bytes,
                          size_t voxel_stride_bytes, size_t meta_stride_bytes, int n_packs,
                          int out_stride, double* t_prev_out, int32_t* failed_at, v2e_pack* out);

/* -- the hot path ----------------------------------------------------------------------------
 * Emulate one pack for each of n_seq independent sequences (1 <= n_seq <= 16 per call).
 * Replaces EventEmulator.forward (v2e/v2e_model.py:290-536) for B=1 per sequence.  State planes
 * are updated in place; `voxel` / `raw_events` / `num_events` are written.  `workspace` must be
 * v2e_workspace_bytes() bytes, zero-initialised once by the caller; its synchronisation words
 * are left zeroed (the fixed-point engine keeps a hint for the next pack in it: reuse the same
 * workspace for consecutive packs of a sequence, any content of the hint is correct). */
int v2e_pack_forward(const v2e_config* cfg, const v2e_pack* packs, int n_seq,
                     void* workspace, size_t workspace_bytes, void* stream);

/* -- a clip: consecutive packs of ONE sequence in one call --------------------------------------
 * Replaces the Python loop `for pack in clip: emulator(pack_frames, pack_stamps)` around
 * EventEmulator.forward (test.py:115-121 / model_v2e2v.py:101 drive it that way): packs[i+1]
 * continues from the state packs[i] leaves, so all entries name the same state planes and are
 * launched in order on `stream` with no host work in between.  Each pack has its own frames /
 * voxel / num_events / status / scalars (v2e_host_pack_scalars, prepared ahead).
 * Chains: when the fixed-point engine's specialised kernel serves the request (reference pack
 * shape, voxel output, no debug maps, noise off or PHILOX leak + shot), up to
 * V2E_MAX_SEQ_PER_LAUNCH consecutive packs go out as ONE launch whose kernel loops over them;
 * anything else is one launch per pack.  Results are identical either way.
 * events: optional array of 2*n_packs cudaEvent_t (NULL: none), recorded before and after each
 * pack's launch; for a chain the first pack's pair brackets the whole launch (its start is
 * recorded right before the kernel) and the other pairs of the chain are empty intervals after
 * it, so that the intervals still add up to the device time of the clip.
 * Same workspace rules as v2e_pack_forward. */
int v2e_clip_forward(const v2e_config* cfg, const v2e_pack* packs, int n_packs,
                     void* workspace, size_t workspace_bytes, void* stream, void* const* events);

/* -- a clip of a batch: consecutive steps of n_seq independent sequences in one call -------------
 * packs[i * n_seq + s] is pack i of sequence s (s < n_seq <= V2E_MAX_SEQ_PER_LAUNCH); step i is
 * what v2e_pack_forward(packs + i * n_seq, n_seq) does; consecutive steps are chained into one
 * launch when the specialised fixed-point kernel serves the request (up to 128 pack records per
 * launch: 16 steps of 8 sequences, 8 steps of 16), otherwise one launch per step, enqueued in
 * order on `stream` with no host work in between (the reference runs the sequences of a batch one emulator after the other,
 * one pack at a time: configs.py:62-63, test.py:115-121).  events: optional 2*n_packs
 * cudaEvent_t, recorded before and after each step (NULL: none). */
int v2e_batch_clip_forward(const v2e_config* cfg, const v2e_pack* packs, int n_packs, int n_seq,
                           void* workspace, size_t workspace_bytes, void* stream,
                           void* const* events);

/* -- standalone event -> voxel grid (utils/event_process.py:15-63 / 66-129) ------------------
 * events: device f32 or f64 [n,4] rows (t,x,y,p), time-sorted.  Unlike the reference this does
 * NOT mutate `events`.  voxel: device f32 [num_bins,H,W], overwritten.  t_first/t_last are
 * events[0,0] / events[-1,0] supplied by the caller when known (pass have_range=0 to let the
 * kernel read them).  torch_semantics=1 applies the `tis >= 0` guard of the _pytorch variant
 * and floor() instead of the numpy unsigned cast. */
int v2e_events_to_voxel(const void* events, int events_f64, int64_t n_events,
                        int num_bins, int width, int height,
                        int have_range, double t_first, double t_last, int torch_semantics,
                        float* voxel, void* stream);

/* -- standalone voxel normalisation (utils/event_process.py:132-154 / 157-177) ---------------
 * In place on device f32 [n] (n = num_bins*H*W).  mode: 0 'std', 1 'maxmin'.
 * hot_pixel_thresh > 0 zeroes |v| > thresh first (20/num_bins torch variant, 25/num_bins numpy
 * variant).  workspace: >= V2E_NORMALIZE_WORKSPACE_BYTES of device scratch (no initialisation
 * needed).  Sums are reduced in a fixed order: results are run-to-run deterministic. */
#define V2E_NORMALIZE_WORKSPACE_BYTES 40960
int v2e_voxel_normalize(float* voxel, int64_t n, int mode, float hot_pixel_thresh,
                        void* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* V2E2V_B200_H */
