// chain of N tiny dependent kernels captured in a CUDA graph: plain vs programmatic dependent launch
#include <cuda_runtime.h>
#include <stdio.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <bool PDL>
__global__ void k_link(const float* in, float* out, int n) {
    if (PDL) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        asm volatile("griddepcontrol.launch_dependents;");
    }
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] * 1.0001f + 1.f;
}

template <bool PDL>
int run(int links, int grid, cudaStream_t st, float* a, float* b, int n, float* ms_out) {
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int l = 0; l < links; ++l) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
        const float* in = (l & 1) ? b : a; float* out = (l & 1) ? a : b;
        CK(cudaLaunchKernelEx(&cfg, k_link<PDL>, in, out, n));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < 200; ++i) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    *ms_out = ms / 200;
    return 0;
}

int main() {
    int n = 1 << 16;
    float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    for (int grid : {1, 32, 256}) {
        float m0, m1;
        if (run<false>(50, grid, st, a, b, n, &m0)) return 1;
        if (run<true>(50, grid, st, a, b, n, &m1)) return 1;
        printf("grid %3d: 50 links plain %.1f us (%.2f us/link)   PDL %.1f us (%.2f us/link)\n", grid, m0 * 1e3, m0 * 20, m1 * 1e3, m1 * 20);
    }
    float h[4]; CK(cudaMemcpy(h, a, 16, cudaMemcpyDeviceToHost)); printf("check %f\n", h[0]);
    return 0;
}
