// param_probe.cu — development aid: back-to-back launch cost against kernel parameter size
// (cooperative launch, 148 x 768 threads, 227 KB shared memory, empty body)
#include <cstdio>
#include <cuda_runtime.h>

template <int N> struct Blob { unsigned w[N / 4]; };

template <int N>
__global__ void __launch_bounds__(768, 1) k_blob(const __grid_constant__ Blob<N> b, long long* out) {
    if (threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        out[blockIdx.x] = t + b.w[blockIdx.x % (N / 4)];
    }
}

template <int N>
void run(cudaStream_t st, long long* d, bool coop) {
    const int smem = 232448;
    cudaFuncSetAttribute(k_blob<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    static Blob<N> b;
    float best = 1e9f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int i = 0; i < 64; ++i) {
            b.w[0] = i;
            void* args[] = {&b, &d};
            if (coop) cudaLaunchCooperativeKernel((const void*)k_blob<N>, dim3(148), dim3(768), args, smem, st);
            else k_blob<N><<<148, 768, smem, st>>>(b, d);
        }
        cudaEventRecord(e1, st);
        cudaStreamSynchronize(st);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%s params %5d B: %.2f us/launch  (%s)\n", coop ? "coop " : "plain", N, 1e3 * best / 64,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    long long* d;
    cudaMalloc(&d, 1 << 16);
    cudaStream_t st;
    cudaStreamCreate(&st);
    for (int coop = 0; coop < 2; ++coop) {
        run<64>(st, d, coop);
        run<512>(st, d, coop);
        run<1024>(st, d, coop);
        run<2048>(st, d, coop);
        run<3072>(st, d, coop);
        run<3968>(st, d, coop);
        run<4064>(st, d, coop);
    }
    return 0;
}
