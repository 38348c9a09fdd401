// ddc_host_emu.h -- TEST INFRASTRUCTURE: host stand-ins of the CUDA device language (-DDDC_HOST_EMU builds).
//
// With this header the kernel sources (ddc_median.cuh, ddc_neighbours.cuh, ddc_kernels.cuh) compile as
// plain host C++.  Scalar intrinsics are defined here; the execution model -- thread / block indices,
// block barriers, warp collectives, dynamic shared memory -- is only DECLARED and comes from whoever links:
//   * oracle/emu_median_harness.cpp: a single "thread", for the scalar fuzzers;
//   * oracle/emu/cuda_emu.cpp: every CUDA thread of a block is a fiber (ucontext), barriers and warp
//     collectives are rendezvous between fibers, blocks run one after the other.  That runs the real
//     kernels -- whole decompositions of small masks -- on a machine without a GPU.
// Nothing of this is part of the product: nvcc never sees this file.
#pragma once
#ifndef DDC_HOST_EMU
#error "ddc_host_emu.h is for -DDDC_HOST_EMU test builds only"
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

// ---- language extensions ------------------------------------------------------------------------
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

// ---- built-in variables -------------------------------------------------------------------------
struct ddc_emu_dim {
    unsigned x, y, z;
};
extern ddc_emu_dim threadIdx, blockIdx, blockDim, gridDim;

// ---- vector types -------------------------------------------------------------------------------
struct __attribute__((aligned(16))) int4 {
    int x, y, z, w;
};
struct __attribute__((aligned(16))) uint4 {
    unsigned x, y, z, w;
};
struct __attribute__((aligned(8))) uint2 {
    unsigned x, y;
};
inline int4 make_int4(int x, int y, int z, int w) { return int4 { x, y, z, w }; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4 { x, y, z, w }; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2 { x, y }; }

// ---- execution model: provided by the runtime that is linked in -----------------------------------
void __syncthreads();
int __syncthreads_or(int predicate);
// every lane named in `mask` deposits v; returns a pointer to the 32 deposited values (valid until the lane's
// next collective) and, through *present, which lanes took part
const unsigned long long* ddc_emu_warp_gather(unsigned mask, unsigned long long v, unsigned* present);
void* ddc_emu_dyn_smem();

// ---- scalar intrinsics --------------------------------------------------------------------------
inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
// the device's fast division is accurate to 2 ulp: test builds can perturb the quotient by that much
extern int ddc_emu_fdividef_ulps;
inline float __fdividef(float a, float b)
{
    float q = a / b;
    if (ddc_emu_fdividef_ulps && q > 0.0f && std::isfinite(q)) { // +-ulps units in the last place
        int32_t bits;
        std::memcpy(&bits, &q, 4);
        bits += ddc_emu_fdividef_ulps;
        std::memcpy(&q, &bits, 4);
    }
    return q;
}
// correctly rounded double operations (test builds are compiled with -ffp-contract=off)
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }

// integer min / max as the device overloads them (mixed signedness promotes like the built-in operators)
template <typename A, typename B>
inline auto min(A a, B b) -> typename std::common_type<A, B>::type
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a < (T)b ? (T)a : (T)b;
}
template <typename A, typename B>
inline auto max(A a, B b) -> typename std::common_type<A, B>::type
{
    typedef typename std::common_type<A, B>::type T;
    return (T)a > (T)b ? (T)a : (T)b;
}

// loads / stores with cache hints are plain accesses
template <typename T>
inline T __ldcs(const T* p) { return *p; }
template <typename T>
inline T __ldcg(const T* p) { return *p; }
template <typename T>
inline T __ldg(const T* p) { return *p; }
template <typename T>
inline void __stcs(T* p, const T& v) { *p = v; }
inline void __threadfence() { }
inline void __threadfence_system() { }
inline void __threadfence_block() { }

// atomics: fibers never run concurrently
template <typename T, typename U>
inline T atomicAdd(T* p, U v)
{
    const T old = *p;
    *p = (T)(old + (T)v);
    return old;
}
template <typename T, typename U>
inline T atomicMax(T* p, U v)
{
    const T old = *p;
    if ((T)v > old)
        *p = (T)v;
    return old;
}
template <typename T, typename U>
inline T atomicMin(T* p, U v)
{
    const T old = *p;
    if ((T)v < old)
        *p = (T)v;
    return old;
}
template <typename T, typename U, typename V>
inline T atomicCAS(T* p, U cmp, V v)
{
    const T old = *p;
    if (old == (T)cmp)
        *p = (T)v;
    return old;
}
template <typename T, typename U>
inline T atomicOr(T* p, U v)
{
    const T old = *p;
    *p = (T)(old | (T)v);
    return old;
}

// ---- warp collectives on top of ddc_emu_warp_gather ---------------------------------------------
inline unsigned ddc_emu_lane() { return (threadIdx.x + threadIdx.y * blockDim.x) & 31u; }
template <typename T>
inline unsigned long long ddc_emu_pack(T v)
{
    unsigned long long u = 0;
    static_assert(sizeof(T) <= 8, "collectives move at most 8 bytes");
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <typename T>
inline T ddc_emu_unpack(unsigned long long u)
{
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}
template <typename T>
inline T __shfl_sync(unsigned mask, T v, int src)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, ddc_emu_pack(v), &present);
    return ddc_emu_unpack<T>(all[src & 31]);
}
template <typename T>
inline T __shfl_up_sync(unsigned mask, T v, unsigned delta)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, ddc_emu_pack(v), &present);
    const unsigned lane = ddc_emu_lane();
    return lane >= delta ? ddc_emu_unpack<T>(all[lane - delta]) : v;
}
template <typename T>
inline T __shfl_down_sync(unsigned mask, T v, unsigned delta)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, ddc_emu_pack(v), &present);
    const unsigned lane = ddc_emu_lane();
    return lane + delta < 32 ? ddc_emu_unpack<T>(all[lane + delta]) : v;
}
template <typename T>
inline T __shfl_xor_sync(unsigned mask, T v, int lanemask)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, ddc_emu_pack(v), &present);
    return ddc_emu_unpack<T>(all[(ddc_emu_lane() ^ (unsigned)lanemask) & 31]);
}
inline unsigned __ballot_sync(unsigned mask, int predicate)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, predicate ? 1ull : 0ull, &present);
    unsigned b = 0;
    for (int l = 0; l < 32; l++)
        if ((present >> l & 1u) && all[l])
            b |= 1u << l;
    return b;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v)
{
    unsigned present;
    const unsigned long long* all = ddc_emu_warp_gather(mask, v, &present);
    unsigned r = 0;
    for (int l = 0; l < 32; l++)
        if (present >> l & 1u)
            r |= (unsigned)all[l];
    return r;
}
// __any_sync(__activemask(), p): "among the lanes that happen to be converged here".  The emulation runs one
// lane at a time, so the converged set is the lane itself (the kernels only use the result as a hint to stop
// checking early).  With an explicit mask it is a real collective.
inline unsigned __activemask() { return 0u; }
inline int __any_sync(unsigned mask, int predicate)
{
    if (mask == 0u)
        return predicate != 0;
    return __ballot_sync(mask, predicate) != 0u;
}
