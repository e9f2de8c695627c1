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

// One grid-stride pass over the grid: statistics of the (hot-pixel filtered) values.  128-bit
// loads when the grid allows; the filter writes back only what it changes.
struct NormAcc {
    double s1 = 0., s2 = 0., nz = 0.;
    float mn = INFINITY, mx = -INFINITY;
    __device__ __forceinline__ void add(float x) {
        s1 += (double)x;
        s2 += (double)x * (double)x;
        nz += x != 0.f ? 1. : 0.;
        mn = fminf(mn, x);
        mx = fmaxf(mx, x);
    }
};

__global__ void __launch_bounds__(kVThreads)
k_norm_partials(float* __restrict__ v, long long n, float hot, int vec4, NormPart* __restrict__ parts) {
    __shared__ NormPart s_red[kVThreads / 32];
    NormAcc a;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (vec4) {
        float4* v4 = (float4*)v;
        for (long long k = t0; k < n / 4; k += stride) {
            float4 x = v4[k];
            if (hot > 0.f) {
                const bool h0 = fabsf(x.x) > hot, h1 = fabsf(x.y) > hot, h2 = fabsf(x.z) > hot, h3 = fabsf(x.w) > hot;
                if (h0) x.x = 0.f;
                if (h1) x.y = 0.f;
                if (h2) x.z = 0.f;
                if (h3) x.w = 0.f;
                if (h0 || h1 || h2 || h3) v4[k] = x;
            }
            a.add(x.x); a.add(x.y); a.add(x.z); a.add(x.w);
        }
    } else {
        for (long long k = t0; k < n; k += stride) {
            float x = v[k];
            if (hot > 0.f && fabsf(x) > hot) { x = 0.f; v[k] = 0.f; }
            a.add(x);
        }
    }
    double s1 = a.s1, s2 = a.s2, nz = a.nz;
    float mn = a.mn, mx = a.mx;
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
        NormPart r = s_red[0];
        for (int i = 1; i < kVThreads / 32; ++i) {
            r.s1 += s_red[i].s1; r.s2 += s_red[i].s2; r.nz += s_red[i].nz;
            r.mn = fminf(r.mn, s_red[i].mn); r.mx = fmaxf(r.mx, s_red[i].mx);
        }
        parts[blockIdx.x] = r;
    }
}

__global__ void __launch_bounds__(kVThreads)
k_norm_apply(float* __restrict__ v, long long n, int mode, int vec4, const NormPart* __restrict__ parts,
             int n_parts) {
    // Every CTA combines the partials itself, in one fixed order (thread t takes parts t, t+256, ...,
    // then a fixed tree over the threads): deterministic, identical in every CTA, and a few hundred
    // nanoseconds instead of one thread walking a thousand records.
    __shared__ double s_d[3][kVThreads];
    __shared__ float s_f[2][kVThreads];
    __shared__ float s_stat[3];
    {
        double s1 = 0., s2 = 0., nz = 0.;
        float mn = INFINITY, mx = -INFINITY;
        for (int i = threadIdx.x; i < n_parts; i += kVThreads) {
            const NormPart q = parts[i];
            s1 += q.s1; s2 += q.s2; nz += q.nz;
            mn = fminf(mn, q.mn); mx = fmaxf(mx, q.mx);
        }
        s_d[0][threadIdx.x] = s1; s_d[1][threadIdx.x] = s2; s_d[2][threadIdx.x] = nz;
        s_f[0][threadIdx.x] = mn; s_f[1][threadIdx.x] = mx;
        __syncthreads();
        for (int o = kVThreads / 2; o > 0; o >>= 1) {
            if ((int)threadIdx.x < o) {
                s_d[0][threadIdx.x] += s_d[0][threadIdx.x + o];
                s_d[1][threadIdx.x] += s_d[1][threadIdx.x + o];
                s_d[2][threadIdx.x] += s_d[2][threadIdx.x + o];
                s_f[0][threadIdx.x] = fminf(s_f[0][threadIdx.x], s_f[0][threadIdx.x + o]);
                s_f[1][threadIdx.x] = fmaxf(s_f[1][threadIdx.x], s_f[1][threadIdx.x + o]);
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        const double a_s1 = s_d[0][0], a_s2 = s_d[1][0], a_nz = s_d[2][0];
        const float a_mn = s_f[0][0], a_mx = s_f[1][0];
        if (mode == 1) {
            s_stat[0] = a_mn;
            s_stat[1] = fadd(fsub(a_mx, a_mn), 1e-8f);
            s_stat[2] = 1.f;
        } else if (a_nz > 0.) {
            const float mean = fdiv((float)a_s1, (float)a_nz);
            const float var = fsub(fdiv((float)a_s2, (float)a_nz), fmul(mean, mean));
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
    auto f = [&](float x) {
        return mode == 1 ? fdiv(fsub(x, a), den) : fdiv(fmul(x != 0.f ? 1.f : 0.f, fsub(x, a)), den);
    };
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (vec4) {
        float4* v4 = (float4*)v;
        for (long long k = t0; k < n / 4; k += stride) {
            float4 x = v4[k];
            x.x = f(x.x); x.y = f(x.y); x.z = f(x.z); x.w = f(x.w);
            v4[k] = x;
        }
    } else {
        for (long long k = t0; k < n; k += stride) v[k] = f(v[k]);
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
    const int vec4 = (n % 4 == 0) && (((uintptr_t)voxel & 15u) == 0);
    long long parts = (n + kVThreads * 16 - 1) / (kVThreads * 16);
    if (parts > kNormParts) parts = kNormParts;
    NormPart* p = (NormPart*)workspace;
    k_norm_partials<<<(unsigned)parts, kVThreads, 0, st>>>(voxel, n, hot_pixel_thresh, vec4, p);
    long long blocks = (n + kVThreads * 16 - 1) / (kVThreads * 16);
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_norm_apply<<<(unsigned)blocks, kVThreads, 0, st>>>(voxel, n, mode, vec4, p, (int)parts);
    count_launch(2);
    return cuda_ok(cudaGetLastError());
}

}  // extern "C"
