// bw_probe.cu -- measures what HBM read / write bandwidth simple access patterns reach on this GPU,
// to set expectations for the mask-scan (read) and label (write) kernels.  Not part of the product.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bw_probe bw_probe.cu && ./bw_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

// V0: linear grid-stride read, U independent 16-byte loads per thread
template <int U, int MODE>
__global__ void __launch_bounds__(256) read_linear(const int4* __restrict__ p, size_t n, unsigned* out)
{
    unsigned acc = 0;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride * U) {
        int4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            size_t j = i + k * stride;
            v[k] = make_int4(0, 0, 0, 0);
            if (j < n) v[k] = MODE == 0 ? p[j] : (MODE == 1 ? __ldcs(p + j) : __ldg(p + j));
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc += (v[k].x > 0) + (v[k].y > 0) + (v[k].z > 0) + (v[k].w > 0);
    }
    if (acc == 0xffffffffu) *out = acc;
}

// V1: the mask-scan pattern: warp = 128-column group, U rows in flight per lane, rows_per_cta rows
template <int U>
__global__ void __launch_bounds__(256) read_rows(const int32_t* __restrict__ mask, int NX, int rows, int NG, int rpc, unsigned* out)
{
    const int lane = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= NG) return;
    const int r0 = blockIdx.y * rpc, r1 = min(rows, r0 + rpc), x = g * 128 + lane * 4;
    unsigned acc = 0;
    for (int r = r0; r < r1; r += U) {
        int4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            v[k] = make_int4(0, 0, 0, 0);
            if (r + k < r1) v[k] = __ldcs(reinterpret_cast<const int4*>(mask + (size_t)(r + k) * NX + x));
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc += (v[k].x > 0) + (v[k].y > 0) + (v[k].z > 0) + (v[k].w > 0);
    }
    if (acc == 0xffffffffu) *out = acc;
}

// V2: warp reads U consecutive 512-byte chunks of ONE row (4 KiB contiguous per warp), rows in sequence
template <int U>
__global__ void __launch_bounds__(256) read_rowwise(const int32_t* __restrict__ mask, int NX, int rows, int rpc, unsigned* out)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int xb = blockIdx.x * (128 * U) + lane * 4;   // column block of 128*U columns
    const int r0 = blockIdx.y * rpc, r1 = min(rows, r0 + rpc);
    unsigned acc = 0;
    for (int r = r0 + w; r < r1; r += 8) {
        int4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) {
            int x = xb + k * 128;
            v[k] = make_int4(0, 0, 0, 0);
            if (x < NX) v[k] = __ldcs(reinterpret_cast<const int4*>(mask + (size_t)r * NX + x));
        }
#pragma unroll
        for (int k = 0; k < U; k++) acc += (v[k].x > 0) + (v[k].y > 0) + (v[k].z > 0) + (v[k].w > 0);
    }
    if (acc == 0xffffffffu) *out = acc;
}

template <int MODE>
__global__ void __launch_bounds__(256) write_linear(int4* __restrict__ p, size_t n)
{
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int4 v = make_int4((int)i, 1, 2, 3);
        if (MODE == 0) p[i] = v; else __stcs(p + i, v);
    }
}
// the label pattern: warp = 128-column group, rows in sequence
__global__ void __launch_bounds__(256) write_rows(int32_t* __restrict__ pid, int NX, int rows, int NG, int rpc)
{
    const int lane = threadIdx.x & 31, g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g >= NG) return;
    const int r0 = blockIdx.y * rpc, r1 = min(rows, r0 + rpc), x = g * 128 + lane * 4;
    for (int r = r0; r < r1; r++)
        __stcs(reinterpret_cast<int4*>(pid + (size_t)r * NX + x), make_int4(r, x, 2, 3));
}

template <typename F>
static void timeit(const char* name, double bytes, F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 2; i++) f();
    cudaEventRecord(a);
    const int N = 5;
    for (int i = 0; i < N; i++) f();
    cudaEventRecord(b); CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= N;
    printf("%-48s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6);
}

int main()
{
    const int NX = 32768, NY = 32768;
    const size_t n = (size_t)NX * NY;
    int32_t *a, *b; unsigned* out;
    CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&out, 4));
    CK(cudaMemset(a, 1, n * 4)); CK(cudaMemset(b, 0, n * 4));
    const double bytes = (double)n * 4;
    const size_t n4 = n / 4;
    timeit("cudaMemcpy D2D (read+write bytes)", 2 * bytes, [&] { cudaMemcpyAsync(b, a, n * 4, cudaMemcpyDeviceToDevice); });
    for (int ctas : { 148 * 4, 148 * 8, 148 * 16, 148 * 32 }) {
        char nm[96];
        snprintf(nm, sizeof nm, "read linear U=4 plain  grid=%d", ctas);
        timeit(nm, bytes, [&] { read_linear<4, 0><<<ctas, 256>>>((const int4*)a, n4, out); });
        snprintf(nm, sizeof nm, "read linear U=8 ldcs   grid=%d", ctas);
        timeit(nm, bytes, [&] { read_linear<8, 1><<<ctas, 256>>>((const int4*)a, n4, out); });
    }
    const int NG = NX / 128, gridx = NG / 8;
    for (int gy : { 23, 46, 75, 256, 1024 }) {
        int rpc = ((NY + gy - 1) / gy + 7) / 8 * 8;
        dim3 grid(gridx, (NY + rpc - 1) / rpc);
        char nm[96];
        snprintf(nm, sizeof nm, "read rows U=8 (scan pattern) grid=%dx%d", grid.x, grid.y);
        timeit(nm, bytes, [&] { read_rows<8><<<grid, 256>>>(a, NX, NY, NG, rpc, out); });
        snprintf(nm, sizeof nm, "read rows U=4 (scan pattern) grid=%dx%d", grid.x, grid.y);
        timeit(nm, bytes, [&] { read_rows<4><<<grid, 256>>>(a, NX, NY, NG, rpc, out); });
        snprintf(nm, sizeof nm, "read rowwise U=8 (4KiB/warp) grid=%dx%d", grid.x, grid.y);
        timeit(nm, bytes, [&] { read_rowwise<8><<<grid, 256>>>(a, NX, NY, rpc, out); });
    }
    for (int ctas : { 148 * 8, 148 * 32 }) {
        char nm[96];
        snprintf(nm, sizeof nm, "write linear plain grid=%d", ctas);
        timeit(nm, bytes, [&] { write_linear<0><<<ctas, 256>>>((int4*)b, n4); });
        snprintf(nm, sizeof nm, "write linear stcs  grid=%d", ctas);
        timeit(nm, bytes, [&] { write_linear<1><<<ctas, 256>>>((int4*)b, n4); });
    }
    for (int gy : { 18, 23, 75, 256, 1024 }) {
        int rpc = ((NY + gy - 1) / gy + 7) / 8 * 8;
        dim3 grid(gridx, (NY + rpc - 1) / rpc);
        char nm[96];
        snprintf(nm, sizeof nm, "write rows (label pattern) grid=%dx%d", grid.x, grid.y);
        timeit(nm, bytes, [&] { write_rows<<<grid, 256>>>(b, NX, NY, NG, rpc); });
    }
    return 0;
}
