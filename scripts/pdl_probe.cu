// pdl_probe.cu -- how long does a dependent kernel take to start behind its predecessor on this box?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o pdl_probe pdl_probe.cu && ./pdl_probe
// A chain of kernels, each of which spins for `busy` us in every block; per kernel: when its first block was on an SM
// (before griddepcontrol.wait), when it passed the wait, when its last block finished.  With and without the
// programmatic-launch attribute, 1 and 148 blocks, 256 and 1024 threads.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__global__ void k(unsigned long long* ts, int idx, int busy_ns, int trigger_early)
{
    if (trigger_early)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0 && ts[3 * idx] == 0ull)
        atomicCAS(ts + 3 * idx, 0ull, now());
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned long long t0 = now();
    if (threadIdx.x == 0 && ts[3 * idx + 1] == 0ull)
        atomicCAS(ts + 3 * idx + 1, 0ull, t0);
    while (now() - t0 < (unsigned long long)busy_ns) { }
    if (threadIdx.x == 0)
        atomicMax(ts + 3 * idx + 2, now());
}
int main()
{
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    unsigned long long *d, h[3 * 8];
    cudaMalloc(&d, sizeof h);
    const int N = 6;
    for (int pdl = 0; pdl < 2; pdl++)
        for (int blocks : { 1, 148, 592 })
            for (int threads : { 256, 1024 }) {
                for (int rep = 0; rep < 3; rep++) {
                    cudaMemsetAsync(d, 0, sizeof h, s);
                    for (int i = 0; i < N; i++) {
                        cudaLaunchConfig_t cfg = {};
                        cfg.gridDim = dim3(blocks);
                        cfg.blockDim = dim3(threads);
                        cfg.stream = s;
                        cudaLaunchAttribute a[1];
                        a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        a[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = a;
                        cfg.numAttrs = pdl ? 1 : 0;
                        cudaLaunchKernelEx(&cfg, k, d, i, 20000, 1);
                    }
                    cudaStreamSynchronize(s);
                    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
                    if (rep < 2)
                        continue;
                    printf("pdl=%d blocks=%3d threads=%4d: ", pdl, blocks, threads);
                    for (int i = 1; i < N; i++)
                        printf("[on SM %+.1f, past wait %+.1f] ", ((double)h[3 * i] - (double)h[3 * (i - 1) + 2]) * 1e-3,
                            ((double)h[3 * i + 1] - (double)h[3 * (i - 1) + 2]) * 1e-3);
                    printf("us after the previous kernel's last block\n");
                }
            }
    cudaError_t e = cudaGetLastError();
    printf("%s\n", cudaGetErrorString(e));
    return 0;
}
