// median_probe.cu -- how long does ONE median of the cut kernels take?  A single thread evaluates the medians
// along a few root -> leaf paths of a realistic column histogram (the synthetic coastline, 32768 columns) from
// shared memory, each timed with clock64().  Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I domain_decomp_b200/csrc -o scripts/median_probe scripts/median_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "ddc_kernels.cuh"
using namespace ddc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__global__ void k_colcount(const int32_t* mask, int NX, int rows, unsigned* col)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)NX * rows; i += (size_t)gridDim.x * blockDim.x)
        if (mask[i] > 0)
            atomicAdd(col + i % NX, 1u);
}

struct Rec { long long cyc; int iters, c0, c1, cut, n; };

template <int VARIANT>
__global__ void __launch_bounds__(1024) k_probe(const unsigned* col, int NX, int P, int levels, Rec* out /* [paths][levels] */, int npaths)
{
    extern __shared__ __align__(16) unsigned smem[];
    __shared__ unsigned wsum[PFX_WS];
    unsigned* pfx = smem;
    unsigned* bitmap = smem + (((size_t)NX + 1 + 3) & ~(size_t)3);
    const uint4* g = reinterpret_cast<const uint4*>(col);
    block_prefix_tiles([&](int base, uint4 (&v)[PFX_Q]) {
#pragma unroll
        for (int q = 0; q < PFX_Q; q++) { const int i = base + q * 4096; v[q] = i < NX ? g[i >> 2] : make_uint4(0, 0, 0, 0); }
    }, NX, pfx, wsum, bitmap);
    const Hist H = make_hist(pfx, bitmap, NX);
    const FastHist F = make_fast_hist(H);
    __syncthreads();
    if (threadIdx.x >= (unsigned)npaths * 32 || (threadIdx.x & 31)) return;  // one lane of one warp per path
    const int path = threadIdx.x >> 5;
    const int nleaves = leaves_below(P, levels);
    int k = (int)((long long)path * (nleaves - 1) / max(1, npaths - 1));
    RcbSet set = { 0, NX, 0, P };
    for (int l = levels; l > 0 && set.n > 1; l--) {
        const int nlo = (set.n - 1) / 2 + 1;
        int it = 0;
        const long long t0 = clock64();
        int cut;
        if (VARIANT == 0) cut = median_boundary(H, set.lo, set.hi - 1, nlo, set.n, &it);
        else cut = median_boundary_fast(F, set.lo, set.hi - 1, nlo, set.n, &it);
        const long long t1 = clock64();
        out[path * levels + (levels - l)] = { t1 - t0, it, set.lo, set.hi, cut, set.n };
        const int below = leaves_below(nlo, l - 1);
        if (k < below) { set.hi = cut; set.n = nlo; }
        else { k -= below; set.lo = cut; set.plo += nlo; set.n -= nlo; }
    }
}

int main()
{
    const int NX = 32768, rows = 2048, P = 16384, levels = 7, npaths = 8;
    int32_t* mask; unsigned* col; Rec* out;
    CK(cudaMalloc(&mask, (size_t)NX * rows * 4)); CK(cudaMalloc(&col, (NX + 4) * 4)); CK(cudaMemset(col, 0, (NX + 4) * 4));
    CK(cudaMalloc(&out, sizeof(Rec) * npaths * levels));
    const uint64_t m = 32768;
    k_generate_mask<<<148 * 8, 256>>>(mask, NX, rows, 12000, 32, m / 16, m / 64, 98000u);
    k_colcount<<<148 * 8, 256>>>(mask, NX, rows, col);
    CK(cudaDeviceSynchronize());
    const size_t smem = 4 * ((((size_t)NX + 1 + 3) & ~(size_t)3) + hist_bitmap_words(NX));
    CK(cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<Rec> h(npaths * levels);
    for (int variant = 0; variant < 2; variant++)
        for (int rep = 0; rep < 2; rep++) {
            CK(cudaMemset(out, 0, sizeof(Rec) * npaths * levels));
            if (variant == 0) k_probe<0><<<1, 1024, smem>>>(col, NX, P, levels, out, npaths);
            else k_probe<1><<<1, 1024, smem>>>(col, NX, P, levels, out, npaths);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), out, sizeof(Rec) * npaths * levels, cudaMemcpyDeviceToHost));
            if (rep == 0) continue;  // cold instruction cache
            printf("variant %s (warm run)\n", variant ? "median_boundary_fast" : "median_boundary (general)");
            long long tot = 0; int tit = 0, n = 0;
            for (int p = 0; p < npaths; p++) {
                printf("  path %d:", p);
                for (int l = 0; l < levels; l++) {
                    const Rec& r = h[p * levels + l];
                    printf(" [%lldcyc %dit w=%d]", r.cyc, r.iters, r.c1 - r.c0);
                    tot += r.cyc; tit += r.iters; n++;
                }
                printf("\n");
            }
            printf("  mean %.0f cycles per median, %.2f iterations, %.0f cycles per iteration-equivalent\n", (double)tot / n, (double)tit / n, (double)tot / tit);
        }
    return 0;
}
