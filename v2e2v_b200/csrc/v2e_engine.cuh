// v2e_engine.cuh — host-side declarations shared by the engines and the C-ABI layer.
#pragma once

#include <atomic>
#include <cstring>

#include "v2e_device.cuh"

namespace v2e {

constexpr int kMaxPartials = 2048;
constexpr size_t kTimelineBytes = 160 * 32 * 8 + 2048 + 1024;  // development aid: [cta][32] clock stamps + [128][2] launch start/end ring + [128] rounds
// fixed-point engine: ring of launch records for multi-sequence launches (8 slots x 16 sequences
// x (SeqDev + 32 FrameSc) <= 8 * 16 * 1280 bytes)
constexpr size_t kFxBlkBytes = 8 * 16 * 1280;
constexpr int kFxBlkSlots = 128;   // launch records (<= 1280 B each) the fixed-point engine's ring holds
constexpr int kRaMaxTiles = 160;  // run-ahead engine: tiles (CTAs) per sequence, <= SM count

// Per-sequence work area inside the caller's workspace.  Zero on entry, zero on exit.
struct SeqWork {
    // per-frame barrier word: low half = CTAs arrived (persistent engine), high half = max
    // event count of the frame (atomicMax); slot V2E_MAX_FRAMES is the statistics barrier
    unsigned long long bar[V2E_MAX_FRAMES + 4];
    unsigned long long raw_count;         // raw mode: rows appended
    unsigned long long nev;               // persistent engine: event counter
    unsigned exit_count;                  // persistent engine: CTAs that left (last one cleans)
    unsigned pad;
    double partials[3 * kMaxPartials];    // per-CTA (sum, sum of squares, nnz)
    // run-ahead engine: flag-free grid barriers.  slot[n][tile] = (tile max of frame n) + 1 and
    // stat[tile][k] = ~bits of (sum, sum of squares, nnz, events); 0 means "not arrived yet".
    unsigned ra_slot[V2E_MAX_FRAMES * kRaMaxTiles];
    unsigned long long ra_stat[4 * kRaMaxTiles];
};
using StepWork = SeqWork;

// Per-sequence work area of the fixed-point engine.  Everything but `guess` is zero on entry and
// on exit; `guess` is a hint that survives from pack to pack (any content is correct, see
// v2e_fixpoint.cu).
constexpr int kFxMaxRounds = V2E_MAX_FRAMES + 4;   // resolve rounds of one launch (pre-pass included)
struct FxWork {
    // grid barrier: `arrive` counts the CTAs that reached a barrier in this launch (monotone:
    // round r is complete at tiles * (r + 1)).  What a round publishes is accumulated into that
    // round's own words with order-free integer atomics before the arrival: the per-frame maxima
    // (max) and the statistics partials (nnz, sum, three 40-bit limbs of the sum of squares,
    // events: adds that cannot overflow 64 bits).  The last CTA to leave zeroes what was used.
    unsigned arrive;
    unsigned exit_count;
    unsigned pad[2];
    unsigned guess[V2E_MAX_FRAMES];  // last pack's per-frame max event count + 1 (0: none yet)
    unsigned gmax[kFxMaxRounds][V2E_MAX_FRAMES];
    unsigned long long stat[kFxMaxRounds][8];
    // pre-pass rounds of the specialised kernel: what they exchange is only a hint, so each tile
    // publishes ONE self-validating 64-bit word (round tag | nine 6-bit saturated maxima) with a
    // relaxed store and polls everybody else's: no atomics, no fences
    unsigned long long hint[2][kRaMaxTiles];
};

struct StepLayout {
    size_t work_off, vox_off, cntpol_off, fx_off, fxstate_off, fxblk_off, tl_off, total;
    int n_partials;
};

inline StepLayout step_layout(const CfgDev& cfg, int n_seq) {
    StepLayout l;
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    l.work_off = 0;
    l.vox_off = up(sizeof(SeqWork) * (size_t)n_seq);
    l.cntpol_off = up(l.vox_off + (size_t)n_seq * cfg.nb * cfg.P * sizeof(float));
    l.fx_off = up(l.cntpol_off + (size_t)n_seq * cfg.P * sizeof(int));
    // (two sets of barrier words per sequence, [set][seq]: the packs of a chain alternate)
    l.fxstate_off = up(l.fx_off + sizeof(FxWork) * (size_t)(2 * n_seq));
    l.fxblk_off = up(l.fxstate_off + (size_t)n_seq * 3 * cfg.P * sizeof(float));
    l.tl_off = up(l.fxblk_off + kFxBlkBytes);
    l.total = up(l.tl_off + kTimelineBytes);
    const size_t N = (size_t)cfg.nb * cfg.P;
    size_t np = (N + 256 * 16 - 1) / (256 * 16);
    if (np < 1) np = 1;
    if (np > (size_t)kMaxPartials) np = kMaxPartials;
    l.n_partials = (int)np;
    return l;
}

inline unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

inline void fill_seq(SeqDev& d, const v2e_pack& p) {
    d.frames = p.frames;
    d.base = p.base_log_frame;
    d.lp = p.lp_log_frame0;
    d.tmem = p.timestamp_mem;
    d.pos = p.pos_thres;
    d.neg = p.neg_thres;
    d.nra = p.noise_rate;
    d.voxel = p.voxel;
    d.raw_events = p.raw_events;
    d.raw_capacity = p.raw_capacity;
    d.num_events = (unsigned long long*)p.num_events;
    d.status = p.status;
    d.max_iters = p.max_iters;
    d.leak_randn = p.leak_randn;
    d.shot_rand = p.shot_rand;
    d.shot_imax = p.shot_imax;
    d.dbg_shot_imax = p.dbg_shot_imax;
    d.dbg_pol = p.dbg_pol;
    d.dbg_cnt = p.dbg_cnt;
    d.dbg_final = p.dbg_final;
    d.dbg_leak_randn = p.dbg_leak_randn;
    d.dbg_shot = p.dbg_shot;
    d.philox_ctr = splitmix64(p.philox_seed) + p.philox_offset;
    d.Tr = p.sc.Tr;
    d.duration = p.sc.duration;
    d.bins_m1 = p.sc.bins_m1;
    d.first_pack = p.sc.first_pack;
    d.shot_f64_mask = p.sc.shot_f64_mask;
    d.T = p.sc.num_frames;
}

inline FrameSc make_frame_sc(const CfgDev& cfg, const v2e_frame_scalars& f) {
    FrameSc s;
    s.t_bin_prev = f.t_bin_prev;
    s.s_large = f.s_large;
    s.s_small = f.s_small;
    s.sdt = f.sdt;
    s.shot_c = f.shot_c;
    const float k = (float)f.shot_c * 256.0f * 1.0001f;
    s.kp = k * cfg.rn_pos;
    s.kn = k * cfg.rn_neg;
    return s;
}

// launch accounting + CUDA error capture (v2e_api.cu)
void count_launch(int n = 1);
int cuda_ok(cudaError_t e);

// engines
int run_stepwise(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st);
int run_persistent(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs,
                   int n_seq, void* workspace, cudaStream_t st);
int run_runahead(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st);
// 1 if the persistent / run-ahead engine can hold n_seq sequences of this config on chip
int persistent_fits(const CfgDev& cfg, int frame_dtype, int n_seq);
int runahead_fits(const CfgDev& cfg, int frame_dtype, int n_seq);
// n_chain > 1: one launch for n_chain consecutive packs of each of the n_seq sequences,
// packs[i * n_seq + s] (V2E_ERR_UNSUPPORTED, with nothing enqueued, when the request does not qualify); ev_launch: optional cudaEvent_t recorded
// right before the kernel launch (after the launch records were copied)
int run_fixpoint(const CfgDev& cfg, int frame_dtype, int rounds, const v2e_pack* packs, int n_seq,
                 void* workspace, cudaStream_t st, int n_chain = 1, void* ev_launch = nullptr);
int fixpoint_fits(const CfgDev& cfg, int frame_dtype, int n_seq);

}  // namespace v2e
