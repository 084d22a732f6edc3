// Dev micro-benchmark: how fast can a chain of dependent kernels run on this GPU?
//   mode 0: plain stream order          mode 1: PDL, trigger at start, griddepcontrol.wait at start
//   mode 2: PDL, wait at start, trigger at end          mode 3: PDL, no wait, per-block flag hand-over, trigger after acquire
// Each block busy-waits `work_ns` of %globaltimer between its "load" and "store".  Prints us per kernel in a CUDA graph.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void k(int mode, int* ready, float* data, int work_ns) {
    if (mode == 3) {
        if (threadIdx.x == 0) {
            int v = 0;
            do { asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(ready + blockIdx.x) : "memory"); } while (v == 0);
            asm volatile("st.relaxed.gpu.global.s32 [%0], %1;" ::"l"(ready + blockIdx.x), "r"(0) : "memory");
        }
        __syncthreads();
        asm volatile("griddepcontrol.launch_dependents;");
    } else {
        if (mode == 1) asm volatile("griddepcontrol.launch_dependents;");
        if (mode != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float x = __ldcg(data + i);
    const unsigned long long t0 = gtime();
    while (gtime() - t0 < (unsigned long long)work_ns) x = x * 1.0000001f + 1e-9f;
    if (mode == 2) asm volatile("griddepcontrol.launch_dependents;");
    data[i] = x;
    if (mode == 3) {
        __syncthreads();
        if (threadIdx.x == 0) { asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(ready + blockIdx.x), "r"(1) : "memory"); }
    }
}

int main(int argc, char** argv) {
    const int T = 256, reps = 200;
    int* ready; float* data;
    const int maxb = 4096;
    cudaMalloc(&ready, maxb * sizeof(int)); cudaMalloc(&data, maxb * 64 * sizeof(float));
    int* ones = (int*)malloc(maxb * sizeof(int)); for (int i = 0; i < maxb; ++i) ones[i] = 1;
    cudaMemset(data, 0, maxb * 64 * sizeof(float));
    cudaStream_t s; cudaStreamCreate(&s);
    for (int blocks : {4, 512, 1024}) for (int work : {0, 3000, 6000}) for (int mode = 0; mode < 4; ++mode) {
        cudaMemcpy(ready, ones, maxb * sizeof(int), cudaMemcpyHostToDevice);
        cudaGraph_t g; cudaGraphExec_t ge;
        cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        for (int t = 0; t < T; ++t) {
            cudaLaunchConfig_t lc = {}; lc.gridDim = dim3(blocks); lc.blockDim = dim3(64); lc.stream = s;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
            lc.attrs = a; lc.numAttrs = mode ? 1 : 0;
            cudaLaunchKernelEx(&lc, k, mode, ready, data, work);
        }
        cudaStreamEndCapture(s, &g);
        cudaGraphInstantiate(&ge, g, 0);
        cudaGraphLaunch(ge, s); cudaStreamSynchronize(s);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        for (int r = 0; r < reps / 10; ++r) cudaGraphLaunch(ge, s);
        cudaEventRecord(e1, s); cudaStreamSynchronize(s);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("blocks %4d work %4d ns mode %d : %.3f us per kernel  (%s)\n", blocks, work, mode, ms * 1e3 / (T * (reps / 10)), cudaGetErrorString(cudaGetLastError()));
        cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
    }
    return 0;
}
