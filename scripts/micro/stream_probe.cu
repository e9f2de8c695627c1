// stream_probe.cu — development aid: what the memory system gives the pass-1 access pattern
// (15 read planes + 8 written planes of P floats, one persistent CTA per SM), with a little
// arithmetic per pixel.  Variants: pixels per lane (2/4), threads per CTA, stores on/off.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int NR = 15, NW = 8;
struct Planes { const float* r[NR]; float* w[NW]; };

template <int VEC, bool STORE>
__global__ void __launch_bounds__(1024, 1) k_stream(Planes pl, int P, int tile_px) {
    const int tile0 = blockIdx.x * tile_px;
    const int tile_n = min(tile_px, P - tile0);
    for (int c0 = threadIdx.x * VEC; c0 < tile_n; c0 += blockDim.x * VEC) {
        const int p = tile0 + c0;
        float acc[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
        float x[NR][VEC];
#pragma unroll
        for (int i = 0; i < NR; ++i) {
            if (VEC == 4) { const float4 t = __ldg((const float4*)(pl.r[i] + p)); x[i][0] = t.x; x[i][1] = t.y; x[i][2] = t.z; x[i][3] = t.w; }
            else { const float2 t = __ldg((const float2*)(pl.r[i] + p)); x[i][0] = t.x; x[i][1] = t.y; }
        }
#pragma unroll
        for (int i = 0; i < NR; ++i)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = acc[v] * 0.5f + x[i][v];
        if (STORE) {
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                if (VEC == 4) *(float4*)(pl.w[i] + p) = make_float4(acc[0], acc[1], acc[2], acc[3]);
                else *(float2*)(pl.w[i] + p) = make_float2(acc[0], acc[1]);
            }
        } else if (acc[0] == 12345.f) pl.w[0][p] = acc[0];
    }
}

int main() {
    const int P = 1280 * 720, SETS = 6;
    float* buf[SETS][NR + NW];
    for (int s = 0; s < SETS; ++s)
        for (int i = 0; i < NR + NW; ++i) { cudaMalloc(&buf[s][i], P * 4); cudaMemset(buf[s][i], 0, P * 4); }
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int tile_px = ((P + 147) / 148 + 63) & ~63;
    auto run = [&](const char* name, auto kern, int threads) {
        float best = 1e9f, sum = 0;
        const int N = 30;
        for (int rep = 0; rep < 3; ++rep) {
            cudaStreamSynchronize(st);
            cudaEventRecord(e0, st);
            for (int i = 0; i < N; ++i) {
                Planes pl;
                for (int k = 0; k < NR; ++k) pl.r[k] = buf[i % SETS][k];
                for (int k = 0; k < NW; ++k) pl.w[k] = buf[i % SETS][NR + k];
                kern<<<148, threads, 0, st>>>(pl, P, tile_px);
            }
            cudaEventRecord(e1, st);
            cudaStreamSynchronize(st);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
            sum += ms;
        }
        const double us = 1e3 * best / N;
        printf("%-28s threads %4d: %.1f us/launch (incl ~2.5 us launch gap) -> %.0f GB/s of %d read + %d written planes (%s)\n",
               name, threads, us, (double)P * 4 * (NR + NW) / us / 1e3, NR, NW, cudaGetErrorString(cudaGetLastError()));
    };
    for (int threads : {512, 768, 1024}) {
        run("vec2 load+store", k_stream<2, true>, threads);
        run("vec4 load+store", k_stream<4, true>, threads);
        run("vec2 load only", k_stream<2, false>, threads);
        run("vec4 load only", k_stream<4, false>, threads);
    }
    return 0;
}
