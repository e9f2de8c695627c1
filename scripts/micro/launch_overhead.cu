// Launch-gap microbenchmark: back-to-back launches of a trivial 148x1024 kernel with 200 KB of
// dynamic shared memory: regular vs cooperative vs programmatic-dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(1024, 1) k(int* p, int spin) {
    extern __shared__ int s[];
    s[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if (p && threadIdx.x == 0 && s[5] == 12345) p[blockIdx.x] = 1;
}
__global__ void __launch_bounds__(1024, 1) kpdl(int* p, int spin) {
    extern __shared__ int s[];
    s[threadIdx.x] = threadIdx.x;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __syncthreads();
    long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (p && threadIdx.x == 0 && s[5] == 12345) p[blockIdx.x] = 1;
}
int main() {
    const int smem = 200 * 1024, N = 300;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kpdl, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaStream_t st; cudaStreamCreate(&st);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int* p = nullptr;
    for (int spin : {0, 20000, 100000}) {
        for (int mode = 0; mode < 3; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0, st);
                for (int i = 0; i < N; ++i) {
                    if (mode == 0) {
                        k<<<148, 1024, smem, st>>>(p, spin);
                    } else if (mode == 1) {
                        void* args[] = {&p, (void*)&spin};
                        cudaLaunchCooperativeKernel((const void*)k, dim3(148), dim3(1024), args, smem, st);
                    } else {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(148); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem; cfg.stream = st;
                        cudaLaunchAttribute at[1];
                        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        at[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = at; cfg.numAttrs = 1;
                        cudaLaunchKernelEx(&cfg, kpdl, p, spin);
                    }
                }
                cudaEventRecord(e1, st);
                cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (ms < best) best = ms;
            }
            printf("spin %6d cycles  mode %s : %.2f us per launch (err %s)\n", spin,
                   mode == 0 ? "regular    " : (mode == 1 ? "cooperative" : "pdl        "), best * 1e3f / N,
                   cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
