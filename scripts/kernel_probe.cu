// kernel_probe.cu -- times the real k_scan_mask / k_label kernels in isolation for several grid
// shapes (rows per CTA), to pick launch parameters.  Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I domain_decomp_b200/csrc -o scripts/kernel_probe scripts/kernel_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ddc_kernels.cuh"
using namespace ddc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <typename F>
static float timeit(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; i++) f();
    cudaEventRecord(a);
    const int N = 5;
    for (int i = 0; i < N; i++) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    CK(cudaGetLastError());
    return ms / N;
}

int main(int argc, char** argv)
{
    const int NX = argc > 1 ? atoi(argv[1]) : 32768, NY = argc > 2 ? atoi(argv[2]) : 32768;
    const size_t n = (size_t)NX * NY;
    const int NG = (NX + 127) / 128, NB = NG * 16, gridx = (NG + 7) / 8;
    int32_t *mask, *pid; uint8_t* bits; unsigned* col; DevScalars* sc;
    CK(cudaMalloc(&mask, n * 4)); CK(cudaMalloc(&pid, n * 4)); CK(cudaMalloc(&bits, (size_t)NY * NB));
    CK(cudaMalloc(&col, (NX + 4) * 4)); CK(cudaMalloc(&sc, sizeof(DevScalars)));
    CK(cudaMemset(col, 0, (NX + 4) * 4)); CK(cudaMemset(sc, 0, sizeof(DevScalars)));
    k_generate_mask<<<148 * 32, 256>>>(mask, NX, NY, 0, 32, NX / 16, NX / 64, 131072);
    CK(cudaDeviceSynchronize());
    const double bytes = (double)n * 4;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_scan_mask<true>, 256, 0);
    printf("k_scan_mask<true>: %d CTAs/SM\n", occ);
    for (int gy : { 148 * occ / gridx, 2 * 148 * occ / gridx, 64, 128, 256, 512, 1024, 2048, 4096 }) {
        int rpc = ((NY + gy - 1) / gy + 7) / 8 * 8;
        dim3 grid(gridx, (NY + rpc - 1) / rpc);
        float ms = timeit([&] { k_scan_mask<true><<<grid, 256>>>(mask, NX, NY, 0, NB, rpc, bits, col, reinterpret_cast<int*>(col) + NX, PeerPush{}, PeerSync{}, nullptr, 0); });
        printf("  scan  grid %dx%-5d rpc %-5d %8.3f ms %8.1f GB/s\n", grid.x, grid.y, rpc, ms, bytes / ms / 1e6);
    }
    // a plausible strip / part layout for the label kernel: 128 strips x 128 parts
    const int S = 128, PPS = 128, P = S * PPS;
    std::vector<int> hs(NX), hp0(S + 1), hy0(P), hey(P);
    for (int x = 0; x < NX; x++) hs[x] = (int)((long long)x * S / NX);
    for (int s = 0; s <= S; s++) hp0[s] = s * PPS;
    for (int p = 0; p < P; p++) { int j = p % PPS; hy0[p] = (int)((long long)j * NY / PPS); hey[p] = (int)((long long)(j + 1) * NY / PPS) - hy0[p]; }
    int *ds, *dp0, *dy0, *dey;
    CK(cudaMalloc(&ds, NX * 4)); CK(cudaMalloc(&dp0, (S + 1) * 4)); CK(cudaMalloc(&dy0, P * 4)); CK(cudaMalloc(&dey, P * 4));
    CK(cudaMemcpy(ds, hs.data(), NX * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dp0, hp0.data(), (S + 1) * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy0, hy0.data(), P * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dey, hey.data(), P * 4, cudaMemcpyHostToDevice));
    NaiveParams nv { 128, 128, NX / 128, NY / 128 };
    Plan* dplan;
    CK(cudaMalloc(&dplan, sizeof(Plan))); CK(cudaMemset(dplan, 0, sizeof(Plan)));
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_label<true, true>, 256, 0);
    printf("k_label<true,true>: %d CTAs/SM\n", occ);
    for (int rpc : { 32, 64, 128, 256, 512, 1024, 2048 }) {
        dim3 grid(gridx, (NY + rpc - 1) / rpc);
        CK(cudaMemset(sc, 0, sizeof(DevScalars)));
        float ms = timeit([&] { k_label<true, true><<<grid, 256>>>(bits, NX, NY, 0, NB, rpc, ds, dp0, dy0, dey, nv, pid, sc, dplan); });
        printf("  label grid %dx%-5d rpc %-5d %8.3f ms %8.1f GB/s\n", grid.x, grid.y, rpc, ms, bytes / ms / 1e6);
    }
    return 0;
}
