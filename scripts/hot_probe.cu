// hot_probe.cu -- variants of the two HBM-bound kernels (mask scan, labelling) timed in isolation, to find
// out what separates them from the plain read / write streams of bw_probe.cu.  Not part of the product.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -I domain_decomp_b200/csrc -o scripts/hot_probe scripts/hot_probe.cu
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

__device__ __forceinline__ uint64_t policy_evict_last()
{
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_v4_hint(void* p, const uint4& v, uint64_t pol)
{
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ unsigned ld_u8_hint(const uint8_t* p, uint64_t pol)
{
    unsigned v;
    asm volatile("ld.global.nc.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
// ---- scan variants -----------------------------------------------------------------------------
// STAGE 0: bit-map bytes stored straight from registers (1 byte per even lane and row)
//       2: per-warp shared staging, __syncwarp only, 16-byte stores (one lane per row)
// BITS / ATOM: write the bit map / do the column atomics (off = upper bounds)
template <int STAGE, bool BITS, bool ATOM, int U>
__global__ void __launch_bounds__(256) scan_v(const int32_t* __restrict__ mask, int NX, int rows, int NB,
    int rows_per_cta, uint8_t* __restrict__ bits, unsigned* __restrict__ colcount, int* __restrict__ yr)
{
    __shared__ __align__(16) uint8_t wst[8][U][16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = blockIdx.x * 8 + warp;
    if (g * 128 >= NX) return;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int x = g * 128 + lane * 4;
    unsigned c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    int ylo = 0x7fffffff, yhi = -1;
    const size_t pitch4 = (size_t)NX >> 2;
    for (int r = r0; r < r1; r += U) {
        const int4* p = reinterpret_cast<const int4*>(mask + (size_t)r * NX + x);
        int4 v[U];
#pragma unroll
        for (int k = 0; k < U; k++) v[k] = __ldcs(p + k * pitch4);
        unsigned packed = 0;
#pragma unroll
        for (int k = 0; k < U; k++) packed |= ocean_nibble(v[k]) << (4 * k);
        c0 += __popc(packed & 0x11111111u); c1 += __popc(packed & 0x22222222u);
        c2 += __popc(packed & 0x44444444u); c3 += __popc(packed & 0x88888888u);
        if (BITS) {
            const unsigned other = __shfl_down_sync(0xffffffffu, packed, 1);
            const unsigned even = (packed & 0x0f0f0f0fu) | ((other & 0x0f0f0f0fu) << 4);
            const unsigned odd = ((packed >> 4) & 0x0f0f0f0fu) | (other & 0xf0f0f0f0u);
            if (STAGE == 0) {
                if (!(lane & 1)) {
                    uint8_t* brow = bits + (size_t)g * 16 + (lane >> 1);
#pragma unroll
                    for (int k = 0; k < U; k++)
                        brow[(size_t)(r + k) * NB] = (uint8_t)(((k & 1) ? odd : even) >> (8 * (k >> 1)));
                }
            } else {
                if (!(lane & 1)) {
#pragma unroll
                    for (int k = 0; k < U; k++)
                        wst[warp][k][lane >> 1] = (uint8_t)(((k & 1) ? odd : even) >> (8 * (k >> 1)));
                }
                __syncwarp();
                if (lane < U) {
                    if (STAGE == 3)
                        st_v4_hint(bits + (size_t)(r + lane) * NB + (size_t)g * 16, *reinterpret_cast<const uint4*>(&wst[warp][lane][0]), policy_evict_last());
                    else
                        *reinterpret_cast<uint4*>(bits + (size_t)(r + lane) * NB + (size_t)g * 16)
                            = *reinterpret_cast<const uint4*>(&wst[warp][lane][0]);
                }
                __syncwarp();
            }
        }
        const unsigned any = __reduce_or_sync(0xffffffffu, packed);
        if (any) { ylo = min(ylo, r + ((__ffs(any) - 1) >> 2)); yhi = r + ((31 - __clz(any)) >> 2); }
    }
    if (ATOM) {
        if (c0) atomicAdd(colcount + x, c0);
        if (c1) atomicAdd(colcount + x + 1, c1);
        if (c2) atomicAdd(colcount + x + 2, c2);
        if (c3) atomicAdd(colcount + x + 3, c3);
    } else if (c0 + c1 + c2 + c3 == 0xffffffffu) colcount[0] = 1;
    if (lane == 0 && yhi >= 0) {
        if (-ylo > yr[0]) atomicMax(&yr[0], -ylo);
        if (yhi > yr[1]) atomicMax(&yr[1], yhi);
    }
}

// ---- label variants ---------------------------------------------------------------------------
// MODE 0: bits -> pid with ONE constant part per thread column (no cursors, no `changes`): what the
//         store stream costs when fed from the bit map
//      1: like the product but without the `changes` test
template <int MODE>
__global__ void __launch_bounds__(256) label_v(const uint8_t* __restrict__ bits, int NX, int rows, int NB,
    int rows_per_cta, const int* __restrict__ strip_of_col, int32_t* __restrict__ pid)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g * 128 >= NX) return;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int x = g * 128 + lane * 4;
    const int p0 = strip_of_col[x], p1 = strip_of_col[x + 1], p2 = strip_of_col[x + 2], p3 = strip_of_col[x + 3];
    const uint8_t* brow = bits + (size_t)g * 16 + (lane >> 1);
    const int sh = (lane & 1) * 4;
    for (int r = r0; r < r1; r += 8) {
        unsigned nb[8];
#pragma unroll
        for (int k = 0; k < 8; k++) nb[k] = MODE == 1 ? ld_u8_hint(brow + (size_t)(r + k) * NB, policy_evict_last()) : (unsigned)__ldg(brow + (size_t)(r + k) * NB);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const unsigned nib = nb[k] >> sh;
            __stcs(reinterpret_cast<int4*>(pid + (size_t)(r + k) * NX + x),
                make_int4((nib & 1u) ? p0 : -1, (nib & 2u) ? p1 : -1, (nib & 4u) ? p2 : -1, (nib & 8u) ? p3 : -1));
        }
    }
}
// MODE 2: one uint4 bit-map load per lane and 8 ROWS x 128 columns?  No: lane = row pair.  A warp takes
//         128 columns x 32 rows; lane l first loads the 16 bit-map bytes of row l (one 16-byte load
//         instead of 8 single-byte loads per 8 rows), the nibbles are then exchanged by shuffles.
__global__ void __launch_bounds__(256) label_w(const uint8_t* __restrict__ bits, int NX, int rows, int NB,
    int rows_per_cta, const int* __restrict__ strip_of_col, int32_t* __restrict__ pid)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (g * 128 >= NX) return;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(rows, r0 + rows_per_cta);
    const int x = g * 128 + lane * 4;
    const int p0 = strip_of_col[x], p1 = strip_of_col[x + 1], p2 = strip_of_col[x + 2], p3 = strip_of_col[x + 3];
    for (int r = r0; r < r1; r += 32) {
        // lane l holds the 128 column bits of row r + l
        const uint4 mine = __ldg(reinterpret_cast<const uint4*>(bits + (size_t)(r + lane) * NB + (size_t)g * 16));
        const int wsel = lane >> 3, sh = (lane & 7) * 4; // my 4 columns: word wsel, bits sh .. sh + 3
#pragma unroll 8
        for (int k = 0; k < 32; k++) {
            const unsigned w0 = __shfl_sync(0xffffffffu, mine.x, k), w1 = __shfl_sync(0xffffffffu, mine.y, k);
            const unsigned w2 = __shfl_sync(0xffffffffu, mine.z, k), w3 = __shfl_sync(0xffffffffu, mine.w, k);
            const unsigned w = wsel == 0 ? w0 : (wsel == 1 ? w1 : (wsel == 2 ? w2 : w3));
            const unsigned nib = w >> sh;
            __stcs(reinterpret_cast<int4*>(pid + (size_t)(r + k) * NX + x),
                make_int4((nib & 1u) ? p0 : -1, (nib & 2u) ? p1 : -1, (nib & 4u) ? p2 : -1, (nib & 8u) ? p3 : -1));
        }
    }
}

int main(int argc, char** argv)
{
    const int NX = 32768, NY = 32768;
    const size_t n = (size_t)NX * NY;
    const int NG = (NX + 127) / 128, NB = NG * 16, gridx = (NG + 7) / 8;
    int32_t *mask, *pid; uint8_t *bits, *bits2; unsigned *col, *col2; int* yr; DevScalars* sc; Plan* plan;
    CK(cudaMalloc(&mask, n * 4)); CK(cudaMalloc(&pid, n * 4)); CK(cudaMalloc(&bits, (size_t)NY * NB)); CK(cudaMalloc(&bits2, (size_t)NY * NB));
    CK(cudaMalloc(&col, (NX + 8) * 4)); CK(cudaMalloc(&col2, (NX + 8) * 4)); CK(cudaMalloc(&sc, sizeof(DevScalars))); CK(cudaMalloc(&plan, sizeof(Plan)));
    CK(cudaMalloc(&yr, 8));
    CK(cudaMemset(col, 0, (NX + 8) * 4)); CK(cudaMemset(col2, 0, (NX + 8) * 4)); CK(cudaMemset(sc, 0, sizeof(DevScalars))); CK(cudaMemset(plan, 0, sizeof(Plan)));
    CK(cudaMemset(yr, 0, 8));
    k_generate_mask<<<148 * 32, 256>>>(mask, NX, NY, 0, 32, NX / 16, NX / 64, 131072);
    CK(cudaDeviceSynchronize());
    const double bytes = (double)n * 4;
    // reference outputs from the product kernel
    {
        dim3 grid(gridx, NY / 128);
        k_scan_mask<true><<<grid, 256>>>(mask, NX, NY, 0, NB, 128, bits, col, yr, PeerPush{}, PeerSync{}, nullptr, 0);
        CK(cudaDeviceSynchronize());
    }
    auto check = [&](const char* what) {
        std::vector<uint8_t> a((size_t)NY * NB), b((size_t)NY * NB);
        CK(cudaMemcpy(a.data(), bits, a.size(), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(b.data(), bits2, b.size(), cudaMemcpyDeviceToHost));
        std::vector<unsigned> c(NX), d(NX);
        CK(cudaMemcpy(c.data(), col, NX * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), col2, NX * 4, cudaMemcpyDeviceToHost));
        printf("    check %s: bits %s, counts %s\n", what, a == b ? "same" : "DIFFER", c == d ? "same" : "DIFFER");
    };
    for (int rpc : { 32, 64, 128 }) {
        dim3 grid(gridx, NY / rpc);
        float ms = timeit([&] { k_scan_mask<true><<<grid, 256>>>(mask, NX, NY, 0, NB, rpc, bits2, col2, yr, PeerPush{}, PeerSync{}, nullptr, 0); });
        printf("scan product            rpc %-4d %8.3f ms %8.1f GB/s\n", rpc, ms, bytes / ms / 1e6);
#define SV(ST, B, A, U)                                                                              \
    {                                                                                              \
        float ms = timeit([&] { scan_v<ST, B, A, U><<<grid, 256>>>(mask, NX, NY, NB, rpc, bits2, col2, yr); }); \
        printf("scan stage %d bits %d atom %d U %d rpc %-4d %8.3f ms %8.1f GB/s\n", ST, B, A, U, rpc, ms, bytes / ms / 1e6); \
    }
        SV(0, false, false, 8) SV(0, false, true, 8) SV(0, true, false, 8) SV(0, true, true, 8) SV(2, true, true, 8) SV(2, true, false, 8)
    }
    {   // correctness of the staged variant
        dim3 grid(gridx, NY / 128);
        CK(cudaMemset(col2, 0, (NX + 8) * 4)); CK(cudaMemset(bits2, 0, (size_t)NY * NB));
        scan_v<2, true, true, 8><<<grid, 256>>>(mask, NX, NY, NB, 128, bits2, col2, yr);
        CK(cudaDeviceSynchronize());
        check("scan_v<2>");
    }
    // label
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
    for (int rpc : { 32, 64, 128 }) {
        dim3 grid(gridx, NY / rpc);
        CK(cudaMemset(sc, 0, sizeof(DevScalars)));
        float ms = timeit([&] { k_label<true, true><<<grid, 256>>>(bits, NX, NY, 0, NB, rpc, ds, dp0, dy0, dey, nv, pid, sc, plan); });
        printf("label product   rpc %-4d %8.3f ms %8.1f GB/s\n", rpc, ms, bytes / ms / 1e6);
        ms = timeit([&] { label_v<0><<<grid, 256>>>(bits, NX, NY, NB, rpc, ds, pid); });
        printf("label const     rpc %-4d %8.3f ms %8.1f GB/s\n", rpc, ms, bytes / ms / 1e6);
        ms = timeit([&] { label_w<<<grid, 256>>>(bits, NX, NY, NB, rpc, ds, pid); });
        printf("label const w   rpc %-4d %8.3f ms %8.1f GB/s\n", rpc, ms, bytes / ms / 1e6);
    }
    // the pair scan -> label back to back, with and without L2 evict_last on the bit map
    for (int rev = 0; rev < 2; rev++) {
        dim3 gs(gridx, NY / 128), gl(gridx, NY / 32);
        float ms = timeit([&] {
            scan_v<2, true, true, 8><<<gs, 256>>>(mask, NX, NY, NB, 128, bits2, col2, yr);
            label_v<0><<<gl, 256>>>(bits2, NX, NY, NB, 32, ds, pid);
        });
        printf("pair plain               %8.3f ms\n", ms);
        ms = timeit([&] {
            scan_v<3, true, true, 8><<<gs, 256>>>(mask, NX, NY, NB, 128, bits2, col2, yr);
            label_v<1><<<gl, 256>>>(bits2, NX, NY, NB, 32, ds, pid);
        });
        printf("pair evict_last bits     %8.3f ms\n", ms);
        ms = timeit([&] { scan_v<3, true, true, 8><<<gs, 256>>>(mask, NX, NY, NB, 128, bits2, col2, yr); });
        printf("scan evict_last alone    %8.3f ms\n", ms);
        ms = timeit([&] { label_v<1><<<gl, 256>>>(bits2, NX, NY, NB, 32, ds, pid); });
        printf("label evict_last alone   %8.3f ms\n", ms);
    }
    return 0;
}
