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


// ---- TMA-style 1-D bulk copies (cp.async.bulk): does a bulk pipeline beat LDG / STG streams? ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* smem, const void* gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem)), "l"(gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem, const void* smem, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem), "r"(smem_u32(smem)), "r"(bytes) : "memory");
}
// persistent CTAs, STAGES x CH bytes of shared memory, one elected thread issues the bulk loads
template <int STAGES, int CH>
__global__ void __launch_bounds__(256) read_bulk(const char* __restrict__ p, size_t nchunks, unsigned* out)
{
    extern __shared__ __align__(128) char sm[];
    __shared__ uint64_t bar[STAGES];
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES; i++) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned acc = 0;
    const size_t first = blockIdx.x, step = gridDim.x;
    // prologue
    if (threadIdx.x == 0)
        for (int i = 0; i < STAGES; i++) {
            const size_t c = first + (size_t)i * step;
            if (c < nchunks) { mbar_expect_tx(&bar[i], CH); bulk_load(sm + i * CH, p + c * CH, CH, &bar[i]); }
        }
    int it = 0;
    for (size_t c = first; c < nchunks; c += step, it++) {
        const int st = it % STAGES;
        const uint32_t par = (it / STAGES) & 1;
        while (!mbar_try_wait(&bar[st], par)) { }
        const int4* q = reinterpret_cast<const int4*>(sm + st * CH);
#pragma unroll 4
        for (int i = threadIdx.x; i < CH / 16; i += 256) {
            const int4 v = q[i];
            acc += (v.x > 0) + (v.y > 0) + (v.z > 0) + (v.w > 0);
        }
        __syncthreads(); // everyone is done with the stage
        const size_t nc = c + (size_t)STAGES * step;
        if (threadIdx.x == 0 && nc < nchunks) { mbar_expect_tx(&bar[st], CH); bulk_load(sm + st * CH, p + nc * CH, CH, &bar[st]); }
    }
    if (acc == 0xffffffffu) *out = acc;
}
template <int STAGES, int CH>
__global__ void __launch_bounds__(256) write_bulk(char* __restrict__ p, size_t nchunks)
{
    extern __shared__ __align__(128) char sm[];
    const size_t first = blockIdx.x, step = gridDim.x;
    int it = 0;
    for (size_t c = first; c < nchunks; c += step, it++) {
        const int st = it % STAGES;
        if (it >= STAGES) { // the bulk store that last read this stage must have finished reading it
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(STAGES - 1) : "memory");
            __syncthreads();
        }
        int4* q = reinterpret_cast<int4*>(sm + st * CH);
#pragma unroll 4
        for (int i = threadIdx.x; i < CH / 16; i += 256) q[i] = make_int4((int)c, i, 2, 3);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0) { bulk_store(p + c * CH, q, CH); asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
    }
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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

    {
        char nm[96];
#define RB(ST, CHK, MULT)                                                                           \
    {                                                                                              \
        cudaFuncSetAttribute(read_bulk<ST, CHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CHK);   \
        snprintf(nm, sizeof nm, "read bulk %d x %d KiB grid=148*%d", ST, CHK / 1024, MULT);        \
        timeit(nm, bytes, [&] { read_bulk<ST, CHK><<<148 * MULT, 256, ST * CHK>>>((const char*)a, (n * 4) / CHK, out); }); \
        cudaFuncSetAttribute(write_bulk<ST, CHK>, cudaFuncAttributeMaxDynamicSharedMemorySize, ST * CHK);  \
        snprintf(nm, sizeof nm, "write bulk %d x %d KiB grid=148*%d", ST, CHK / 1024, MULT);       \
        timeit(nm, bytes, [&] { write_bulk<ST, CHK><<<148 * MULT, 256, ST * CHK>>>((char*)b, (n * 4) / CHK); }); \
    }
        RB(4, 16384, 1) RB(4, 16384, 2) RB(4, 16384, 3) RB(4, 32768, 1) RB(6, 32768, 1) RB(8, 8192, 3) RB(3, 16384, 4)
        CK(cudaDeviceSynchronize());
    }
    return 0;
}
