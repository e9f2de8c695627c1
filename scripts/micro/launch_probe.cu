// launch_probe.cu — development aid: (1) %globaltimer against CUDA events, (2) back-to-back launch
// cost of a 148 x 768-thread, 227 KB-shared-memory kernel: plain launch vs cooperative launch
// vs programmatic dependent launch.   nvcc -arch=sm_100a -o launch_probe launch_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>

__global__ void __launch_bounds__(768, 1) k_spin(long long* out, long long spin_ns, int pdl) {
    extern __shared__ unsigned char smem[];
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do { asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1)); } while (t1 - t0 < spin_ns);
    const long long c1 = clock64();
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) {
        smem[0] = 1;
        out[3 * blockIdx.x] = t0;
        out[3 * blockIdx.x + 1] = t1;
        out[3 * blockIdx.x + 2] = c1 - c0;
    }
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * 3 * 8 * 64);
    const int smem = 232448;
    cudaFuncSetAttribute(k_spin, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int mode = 0; mode < 3; ++mode) {
        for (long long spin : {0ll, 20000ll, 100000ll}) {
            const int N = 64;
            float best = 1e9f;
            long long h[148 * 3 * 64];
            for (int rep = 0; rep < 5; ++rep) {
                cudaStreamSynchronize(st);
                cudaEventRecord(e0, st);
                for (int i = 0; i < N; ++i) {
                    long long* o = d + (size_t)i * 148 * 3;
                    int pdl = mode == 2;
                    void* args[] = {&o, &spin, &pdl};
                    if (mode == 1) {
                        cudaLaunchCooperativeKernel((const void*)k_spin, dim3(148), dim3(768), args, smem, st);
                    } else if (mode == 2) {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(148); cfg.blockDim = dim3(768); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        cudaLaunchKernelEx(&cfg, k_spin, o, spin, pdl);
                    } else {
                        k_spin<<<148, 768, smem, st>>>(o, spin, 0);
                    }
                }
                cudaEventRecord(e1, st);
                cudaStreamSynchronize(st);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            // kernel span by globaltimer: first CTA start -> last CTA end, averaged over launches
            double span = 0, gap = 0, cyc = 0;
            long long prev_end = 0;
            for (int i = 0; i < N; ++i) {
                long long s = h[(size_t)i * 444], e = h[(size_t)i * 444 + 1];
                for (int b = 0; b < 148; ++b) {
                    if (h[(size_t)i * 444 + 3 * b] < s) s = h[(size_t)i * 444 + 3 * b];
                    if (h[(size_t)i * 444 + 3 * b + 1] > e) e = h[(size_t)i * 444 + 3 * b + 1];
                }
                span += (double)(e - s);
                cyc += (double)h[(size_t)i * 444 + 2];
                if (i) gap += (double)(s - prev_end);
                prev_end = e;
            }
            printf("mode %s spin %6lld ns: events %.2f us/launch | globaltimer: kernel span %.2f us, gap between launches %.2f us, clock64 per spin %.0f cycles\n",
                   mode == 0 ? "plain" : mode == 1 ? "coop " : "pdl  ", spin, 1e3 * best / N, span / N / 1e3, gap / (N - 1) / 1e3, cyc / N);
        }
    }
    cudaError_t err = cudaGetLastError();
    printf("last error: %s\n", cudaGetErrorString(err));
    return 0;
}
