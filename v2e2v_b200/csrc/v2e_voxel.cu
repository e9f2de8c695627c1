// v2e_voxel.cu — standalone event-list -> voxel grid and voxel normalisation.
//
//   v2e_events_to_voxel  : events_to_voxel_grid (utils/event_process.py:15-63, numpy) and
//                          events_to_voxel_grid_pytorch (:66-129)
//   v2e_voxel_normalize  : event_preprocess (:132-154) / event_preprocess_pytorch (:157-177)
//
// The scatter is HBM/L2-atomic bound: 16 or 32 B read per event, two f32 REDs per event into a
// grid that is written once.  Events arrive time-sorted, i.e. spatially random, so the grid
// tile a CTA would privatise in shared memory is hit by ~N/tiles events from the whole list;
// the scatter therefore uses L2 reductions (RED.ADD.F32, no return value) and keeps the loads
// fully coalesced.  (The fused emulator never scatters: each pixel owns its bins on chip.)
#include "v2e_engine.cuh"

namespace v2e {

constexpr int kVThreads = 256;
constexpr int kNormParts = 1024;

template <typename ET, bool TORCH>
__global__ void __launch_bounds__(kVThreads)
k_events_to_voxel(const ET* __restrict__ ev, long long n, int nb, int W, int H, int have_range,
                  double t_first, double t_last, float* __restrict__ vox) {
    const long long P = (long long)W * H;
    ET first, last;
    if (have_range) { first = (ET)t_first; last = (ET)t_last; }
    else { first = ev[0]; last = ev[4 * (n - 1)]; }
    ET dT = last - first;
    if (dT == (ET)0) dT = (ET)1;
    const ET scale = (ET)(nb - 1);
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        // one 16 B (f32) or 2x16 B (f64) vector load per event row
        ET t, x, y, pol;
        if (sizeof(ET) == 4) {
            const float4 r = __ldg((const float4*)ev + k);
            t = (ET)r.x; x = (ET)r.y; y = (ET)r.z; pol = (ET)r.w;
        } else {
            const double2 a = __ldg((const double2*)ev + 2 * k);
            const double2 b = __ldg((const double2*)ev + 2 * k + 1);
            t = (ET)a.x; x = (ET)a.y; y = (ET)b.x; pol = (ET)b.y;
        }
        // ts = (nb-1)*(t-first)/deltaT, each op rounded separately (:40 / :98)
        ET ts;
        if (sizeof(ET) == 4) ts = (ET)fdiv(fmul((float)scale, fsub((float)t, (float)first)), (float)dT);
        else ts = (ET)__ddiv_rn(__dmul_rn((double)scale, __dsub_rn((double)t, (double)first)), (double)dT);
        if (pol == (ET)0) pol = (ET)-1;
        const ET tf = TORCH ? floor(ts) : trunc(ts);
        const long long ti = (long long)tf;
        ET d, left, right;
        if (sizeof(ET) == 4) {
            d = (ET)fsub((float)ts, (float)tf);
            left = (ET)fmul((float)pol, fsub(1.0f, (float)d));
            right = (ET)fmul((float)pol, (float)d);
        } else {
            d = (ET)__dsub_rn((double)ts, (double)tf);
            left = (ET)__dmul_rn((double)pol, __dsub_rn(1.0, (double)d));
            right = (ET)__dmul_rn((double)pol, (double)d);
        }
        if (ti < 0) continue;
        const long long px = (long long)x + (long long)y * W;
        if (px < 0 || px >= P) continue;  // the reference would index out of bounds
        if (ti < nb) atomicAdd(vox + px + ti * P, (float)left);
        if (ti + 1 < nb) atomicAdd(vox + px + (ti + 1) * P, (float)right);
    }
}

// ---- normalisation -------------------------------------------------------------------------
struct NormPart { double s1, s2, nz, pad; float mn, mx; };

__global__ void __launch_bounds__(kVThreads)
k_norm_partials(float* __restrict__ v, long long n, float hot, NormPart* __restrict__ parts) {
    __shared__ NormPart s_red[kVThreads / 32];
    double s1 = 0., s2 = 0., nz = 0.;
    float mn = INFINITY, mx = -INFINITY;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        float x = v[k];
        if (hot > 0.f && fabsf(x) > hot) { x = 0.f; v[k] = 0.f; }
        s1 += (double)x;
        s2 += (double)x * (double)x;
        nz += x != 0.f ? 1. : 0.;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
        nz += __shfl_down_sync(0xffffffffu, nz, o);
        mn = fminf(mn, __shfl_down_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_down_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = NormPart{s1, s2, nz, 0., mn, mx};
    __syncthreads();
    if (threadIdx.x == 0) {
        NormPart a = s_red[0];
        for (int i = 1; i < kVThreads / 32; ++i) {
            a.s1 += s_red[i].s1; a.s2 += s_red[i].s2; a.nz += s_red[i].nz;
            a.mn = fminf(a.mn, s_red[i].mn); a.mx = fmaxf(a.mx, s_red[i].mx);
        }
        parts[blockIdx.x] = a;
    }
}

__global__ void __launch_bounds__(kVThreads)
k_norm_apply(float* __restrict__ v, long long n, int mode, const NormPart* __restrict__ parts,
             int n_parts) {
    __shared__ float s_stat[3];
    if (threadIdx.x == 0) {
        NormPart a = parts[0];  // fixed-order combine: deterministic and identical in every CTA
        for (int i = 1; i < n_parts; ++i) {
            a.s1 += parts[i].s1; a.s2 += parts[i].s2; a.nz += parts[i].nz;
            a.mn = fminf(a.mn, parts[i].mn); a.mx = fmaxf(a.mx, parts[i].mx);
        }
        if (mode == 1) {
            s_stat[0] = a.mn;
            s_stat[1] = fadd(fsub(a.mx, a.mn), 1e-8f);
            s_stat[2] = 1.f;
        } else if (a.nz > 0.) {
            const float mean = fdiv((float)a.s1, (float)a.nz);
            const float var = fsub(fdiv((float)a.s2, (float)a.nz), fmul(mean, mean));
            s_stat[0] = mean;
            s_stat[1] = fadd(sqrtf(var), 1e-8f);
            s_stat[2] = 1.f;
        } else {
            s_stat[2] = 0.f;
        }
    }
    __syncthreads();
    if (s_stat[2] == 0.f) return;
    const float a = s_stat[0], den = s_stat[1];
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n;
         k += (long long)gridDim.x * blockDim.x) {
        const float x = v[k];
        if (mode == 1) v[k] = fdiv(fsub(x, a), den);
        else v[k] = fdiv(fmul(x != 0.f ? 1.f : 0.f, fsub(x, a)), den);
    }
}

}  // namespace v2e

using namespace v2e;

extern "C" {

int v2e_events_to_voxel(const void* events, int events_f64, int64_t n_events, int num_bins,
                        int width, int height, int have_range, double t_first, double t_last,
                        int torch_semantics, float* voxel, void* stream) {
    if (num_bins <= 0 || width <= 0 || height <= 0 || !voxel || n_events < 0) return V2E_ERR_BAD_ARG;
    if (n_events > 0 && (!events || ((uintptr_t)events & 15u))) return V2E_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)num_bins * width * height * sizeof(float);
    int rc = cuda_ok(cudaMemsetAsync(voxel, 0, bytes, st));
    if (rc != V2E_OK || n_events == 0) return rc;
    long long blocks = (n_events + kVThreads - 1) / kVThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
#define V2E_LAUNCH(ET, TORCH)                                                                   \
    k_events_to_voxel<ET, TORCH><<<(unsigned)blocks, kVThreads, 0, st>>>(                       \
        (const ET*)events, n_events, num_bins, width, height, have_range, t_first, t_last, voxel)
    if (events_f64) { if (torch_semantics) V2E_LAUNCH(double, true); else V2E_LAUNCH(double, false); }
    else { if (torch_semantics) V2E_LAUNCH(float, true); else V2E_LAUNCH(float, false); }
#undef V2E_LAUNCH
    count_launch();
    return cuda_ok(cudaGetLastError());
}

int v2e_voxel_normalize(float* voxel, int64_t n, int mode, float hot_pixel_thresh, void* workspace,
                        void* stream) {
    if (!voxel || n < 0 || !workspace || (mode != 0 && mode != 1)) return V2E_ERR_BAD_ARG;
    if (n == 0) return V2E_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long parts = (n + kVThreads * 8 - 1) / (kVThreads * 8);
    if (parts > kNormParts) parts = kNormParts;
    NormPart* p = (NormPart*)workspace;
    k_norm_partials<<<(unsigned)parts, kVThreads, 0, st>>>(voxel, n, hot_pixel_thresh, p);
    long long blocks = (n + kVThreads * 4 - 1) / (kVThreads * 4);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_norm_apply<<<(unsigned)blocks, kVThreads, 0, st>>>(voxel, n, mode, p, (int)parts);
    count_launch(2);
    return cuda_ok(cudaGetLastError());
}

}  // extern "C"
