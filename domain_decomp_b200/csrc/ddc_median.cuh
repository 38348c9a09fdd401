// ddc_median.cuh -- the scalar core of the cut kernels: Zoltan_RB_find_median restated on a histogram, the
// bit map of non-empty bins it queries, and the barrier-free walk of the RCB tree.  No thread / warp / block
// primitive appears in this file, so the very same source also compiles as plain host C++ (define
// DDC_HOST_EMU, see oracle/emu_median_harness.cpp): the CPU suite fuzzes these functions against the
// oracle's literal double-precision Zoltan loop on millions of histograms, including the effect of an
// approximate __fdividef.  Included by ddc_kernels.cuh.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef DDC_HOST_EMU
#include "ddc_host_emu.h" // oracle/emu: host stand-ins of the device language (test builds only)
#define DDC_NOINLINE
#else
#define DDC_NOINLINE __noinline__
#endif

namespace ddc {

// A histogram as the median search sees it: pfx[i] = number of dots in bins [0, i), plus an optional
// three-level bit map of the non-empty bins (l0: one bit per bin, l1: one bit per l0 word, l2: one
// bit per l1 word) that answers "nearest non-empty bin at or below / above" in a handful of
// dependent loads instead of a binary search over the prefix sums.  On a land-sea mask whole runs
// of rows of a strip are land, so these queries are most of what Zoltan's median loop does.
struct Hist {
    const unsigned* pfx;
    const unsigned *l0, *l1, *l2; // l0 == nullptr: no bit map, search the prefix sums
    int nl2; // words of l2
};
__device__ __forceinline__ Hist make_hist(const unsigned* pfx, const unsigned* bitmap, int n)
{
    Hist H;
    H.pfx = pfx;
    const int tiles = (n + 32767) / 32768;
    H.l0 = bitmap;
    H.l1 = bitmap ? bitmap + tiles * 1024 : nullptr;
    H.l2 = bitmap ? bitmap + tiles * (1024 + 32) : nullptr;
    H.nl2 = tiles;
    return H;
}
constexpr int HIST_TILE = 32768; // bins per tile of block_prefix_tiles (1024 threads)
__host__ __device__ inline size_t hist_bitmap_words(int n) // words behind pfx for n bins
{
    const size_t tiles = ((size_t)n + HIST_TILE - 1) / HIST_TILE;
    return tiles * (1024 + 32 + 1);
}
// largest non-empty bin in [a, t], -1 if none
__device__ __forceinline__ int bitmap_prev(const Hist& H, int a, int t)
{
    if (t < a)
        return -1;
    int w = t >> 5;
    unsigned m = H.l0[w] & (0xffffffffu >> (31 - (t & 31)));
    if (!m) {
        int w1 = w >> 5;
        m = H.l1[w1] & ((1u << (w & 31)) - 1u);
        if (!m) {
            int w2 = w1 >> 5;
            unsigned m2 = H.l2[w2] & ((1u << (w1 & 31)) - 1u);
            while (!m2) {
                if (w2 == 0)
                    return -1;
                m2 = H.l2[--w2];
            }
            w1 = (w2 << 5) + 31 - __clz(m2);
            m = H.l1[w1];
        }
        w = (w1 << 5) + 31 - __clz(m);
        m = H.l0[w];
    }
    const int j = (w << 5) + 31 - __clz(m);
    return j >= a ? j : -1;
}
// smallest non-empty bin in [t, b], -1 if none
__device__ __forceinline__ int bitmap_next(const Hist& H, int t, int b)
{
    if (t > b)
        return -1;
    int w = t >> 5;
    unsigned m = H.l0[w] & (0xffffffffu << (t & 31));
    if (!m) {
        int w1 = w >> 5;
        m = (w & 31) == 31 ? 0u : H.l1[w1] & (0xffffffffu << ((w & 31) + 1));
        if (!m) {
            int w2 = w1 >> 5;
            unsigned m2 = (w1 & 31) == 31 ? 0u : H.l2[w2] & (0xffffffffu << ((w1 & 31) + 1));
            while (!m2) {
                if (++w2 >= H.nl2)
                    return -1;
                m2 = H.l2[w2];
            }
            w1 = (w2 << 5) + __ffs(m2) - 1;
            m = H.l1[w1];
        }
        w = (w1 << 5) + __ffs(m) - 1;
        m = H.l0[w];
    }
    const int j = (w << 5) + __ffs(m) - 1;
    return j <= b ? j : -1;
}

// ------------------------------------------------------------------------------------------------
// Zoltan_RB_find_median on a histogram (device restatement; see DESIGN.md "median")
// ------------------------------------------------------------------------------------------------
// pfx[i] = number of dots in bins [0, i).  All doubles are combined with explicit
// round-to-nearest intrinsics so that no FMA contraction can change Zoltan's arithmetic.
__device__ __forceinline__ unsigned hcnt(const unsigned* pfx, int a, int b)
{
    return b < a ? 0u : pfx[b + 1] - pfx[a];
}
__device__ inline int last_nonempty(const unsigned* pfx, int a, int b)
{
    if (b < a || pfx[b + 1] == pfx[a])
        return -1;
    const unsigned target = pfx[b + 1];
    if (pfx[b] != target)
        return b; // bin b itself holds a dot (the common case on a real coastline)
    int lo = a, hi = b;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (pfx[mid + 1] >= target)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
__device__ inline int first_nonempty(const unsigned* pfx, int a, int b)
{
    if (b < a || pfx[b + 1] == pfx[a])
        return -1;
    const unsigned base = pfx[a];
    if (pfx[a + 1] != base)
        return a; // bin a itself holds a dot
    int lo = a, hi = b;
    while (lo < hi) {
        const int mid = lo + ((hi - lo) >> 1);
        if (pfx[mid + 1] > base)
            hi = mid;
        else
            lo = mid + 1;
    }
    return lo;
}
// the same two queries on a Hist: the bit map when there is one
__device__ __forceinline__ int last_nonempty(const Hist& H, int a, int b)
{
    return H.l0 ? bitmap_prev(H, a, b) : last_nonempty(H.pfx, a, b);
}
__device__ __forceinline__ int first_nonempty(const Hist& H, int a, int b)
{
    return H.l0 ? bitmap_next(H, a, b) : first_nonempty(H.pfx, a, b);
}

// floor() of Zoltan's interpolated guess
//     tmp_half = valuemin + (targetlo - weightlo) / (weight - weightlo - weighthi) * (valuemax - valuemin)
// clamped to [alo - 1, ahi].  Only floor(tmp_half) matters, and FP64 instructions have a long
// latency on this part, so the offset from valuemin is first formed in FP32 from exact integers
// (targetlo = Ti + Tfrac with Ti an integer, so targetlo - weightlo starts from an exact integer
// difference): its relative error is below 2^-21, i.e. below range * 2^-21 bins.  Unless the result
// lies that close to an integer -- where the rounding of the real IEEE sequence could matter -- its
// floor IS the floor of the exact sequence.  Otherwise (and for ranges FP32 cannot hold) the exact
// sequence is evaluated in FP64: correctly rounded subtraction, division, multiplication, addition.
__device__ __forceinline__ int guess_bin(int vmin, int range, unsigned Ti, float Tfrac, double T, unsigned wlo,
    unsigned den, int alo, int ahi)
{
    int t;
    bool sure = false;
    if (range < (1 << 22) && den != 0u) {
        const float numf = (float)(Ti - wlo) + Tfrac;
        const float rangef = (float)range;
        const float off = __fdividef(numf, (float)den) * rangef;
        const float fl = floorf(off), fr = off - fl;
        const float guard = rangef * 0x1p-20f + 0x1p-18f;
        sure = fr > guard && fr < 1.0f - guard;
        t = vmin + (int)fl;
    }
    if (!sure) {
        const double tmp = __dadd_rn((double)vmin,
            __dmul_rn(__ddiv_rn(__dsub_rn(T, (double)wlo), (double)den), (double)range));
        if (tmp < (double)alo)
            return alo - 1;
        if (tmp >= (double)ahi)
            return ahi;
        return (int)floor(tmp);
    }
    // tmp < alo -> alo - 1;  tmp >= ahi -> ahi;  else floor(tmp)
    return min(max(t, alo - 1), ahi);
}

// Integer boundary ceil(cut) of the weighted-median cut of bins [c0, c1] for a set of num_parts
// parts whose lower child receives nlo parts.  *iters += median iterations.
//
// Zoltan keeps weightlo / weighthi / totallo / totalhi as doubles, but with unit weights they are
// exact integers: here they are integers read off the prefix sums (dots in [c0, t] = weightlo +
// totallo, dots in (t, c1] = weighthi + totalhi).  Comparing such an integer m with a target T is
// done against ceil(T): m < T <=> m < ceil(T); and Zoltan's tolerance test fl(T - m) <= 1.0 (for
// m < T, where the subtraction is exact because m is a multiple of ulp(T)) <=> m >= ceil(T) - 1.
// FP64 is only touched for the tie rule that ends a search, for an ambiguous guess, and to form
// the targets when the part counts are not split in half (fractionlo != 0.5).
// (Out of line: the general form serves the configurations off the fast path -- histograms in global memory or
//  beyond 32768 bins, more than 1024 leaves -- from one copy of the code; median_boundary_fast below is the one
//  the benchmark grids run.)
__device__ DDC_NOINLINE inline int median_boundary(const Hist& H, int c0, int c1, int nlo, int num_parts,
    int* iters)
{
    const unsigned* pfx = H.pfx;
    const unsigned base = pfx[c0];
    const unsigned Wn = c1 < c0 ? 0u : pfx[c1 + 1] - base;
    if (Wn == 0u) { // no dot at all: integer midpoint of the inherited range (policy Q2)
        *iters += 1;
        return c0 + ((c1 + 1 - c0) >> 1);
    }
    // targetlo = fl(fl(nlo / num_parts) * W), targethi = fl(W - targetlo)
    double T, Thi;
    unsigned Ti, ceilT, ceilThi;
    float Tfrac;
    if (2 * nlo == num_parts) { // fractionlo = 0.5: both targets are W / 2 exactly
        T = Thi = 0.5 * (double)Wn;
        Ti = Wn >> 1;
        Tfrac = (Wn & 1u) ? 0.5f : 0.0f;
        ceilT = ceilThi = (Wn + 1u) >> 1;
    } else {
        T = __dmul_rn(__ddiv_rn((double)nlo, (double)num_parts), (double)Wn);
        Thi = __dsub_rn((double)Wn, T);
        const double fT = floor(T);
        Ti = (unsigned)fT;
        Tfrac = (float)(T - fT);
        ceilT = (unsigned)ceil(T);
        ceilThi = (unsigned)ceil(Thi);
    }
    const int first = first_nonempty(H, c0, c1), last = last_nonempty(H, c0, c1);
    int vmin = first, vmax = last; // valuemin, valuemax (always bin indices)
    int alo = first, ahi = last; // the active bins
    int B;
    unsigned wlo = 0, whi = 0; // weightlo, weighthi
    int it = 0;
    for (;;) {
        const int t = guess_bin(vmin, vmax - vmin, Ti, Tfrac, T, wlo, Wn - wlo - whi, alo, ahi);
        it++;
        B = t;
        const unsigned cum = pfx[t + 1] - base; // dots in [c0, t] = weightlo + totallo
        if (cum < ceilT) { // lower half TOO SMALL (weightlo + totallo < targetlo)
            const int vhi = first_nonempty(H, t + 1, ahi);
            if (vhi < 0)
                break;
            const unsigned moved = pfx[vhi + 1] - base; // weightlo + wthi
            if (moved >= ceilT) { // the bin that crosses the target: Zoltan's tie rules, in FP64
                const double over = __dsub_rn((double)moved, T), under = __dsub_rn(T, (double)cum);
                if (moved - cum == 1u ? over < under // a single dot moves only if strictly better
                                      : !(over > under)) // a whole column moves unless strictly worse
                    B = vhi;
                break;
            }
            B = vhi;
            wlo = moved;
            if (moved >= ceilT - 1u) // targetlo - weightlo <= tolerance (the weight of one dot)
                break;
            vmin = vhi;
            alo = vhi + 1;
        } else {
            const unsigned above = Wn - cum; // dots in (t, c1] = weighthi + totalhi
            if (above >= ceilThi)
                break; // both halves just right
            // upper half TOO SMALL
            const int vlo = last_nonempty(H, alo, t);
            if (vlo < 0)
                break;
            const unsigned moved = Wn - (pfx[vlo] - base); // weighthi + wtlo = dots in [vlo, c1]
            if (moved >= ceilThi) {
                const double over = __dsub_rn((double)moved, Thi), under = __dsub_rn(Thi, (double)above);
                if (moved - above == 1u ? over < under : !(over > under))
                    B = vlo - 1;
                break;
            }
            B = vlo - 1;
            whi = moved;
            if (moved >= ceilThi - 1u)
                break;
            vmax = vlo;
            ahi = vlo - 1;
        }
    }
    *iters += it;
    // AVERAGE_CUTS over all dots of the set, then ceil() (ZoltanPartitioner.cpp:177-180)
    const int L = last_nonempty(H, c0, B), U = first_nonempty(H, B + 1, c1);
    if (L >= 0 && U >= 0)
        return (L + U + 1) >> 1; // ceil(0.5 * (L + U))
    if (L >= 0)
        return L + 1; // policy Q2
    return U; // policy Q2 (U >= 0 because Wn > 0)
}

// ------------------------------------------------------------------------------------------------
// The same search for the case every benchmark grid is in: prefix sums and bit map in SHARED memory and at
// most 32768 bins (one tile of the bit map: its third level is ONE word, kept in a register).
// ------------------------------------------------------------------------------------------------
// median_boundary above costs ~660 SASS instructions per median (ncu, r2c: 86 branches, generic-pointer address
// arithmetic, FP64 tie rules, both the bit-map and the binary-search paths compiled in) and a median is one
// dependent chain: a cut kernel is 7 such chains in a row.  Here the accesses are 32-bit ld.shared, the bit-map
// queries are straight-line, and when the part counts split in half (always, for a power-of-two part count) the
// targets are W / 2 and Zoltan's FP64 tie rules become exact integer comparisons: (moved - T) < (T - cum) <=>
// moved + cum < W (every term is a multiple of 1/2 below 2^32, so the doubles were exact).  Same boundaries, same
// iteration counts: fuzzed against the oracle next to median_boundary (tests/test_device_median_host.py).
#ifndef DDC_HOST_EMU
typedef unsigned sm_ptr; // a shared-memory address
__device__ __forceinline__ unsigned sm_ld(sm_ptr base, int i)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + 4u * (unsigned)i));
    return v;
}
__device__ __forceinline__ sm_ptr sm_addr(const unsigned* p) { return (sm_ptr)__cvta_generic_to_shared(p); }
#else
typedef const unsigned* sm_ptr;
inline unsigned sm_ld(sm_ptr base, int i) { return base[i]; }
inline sm_ptr sm_addr(const unsigned* p) { return p; }
#endif
struct FastHist {
    sm_ptr pfx, l0, l1;
    unsigned l2;
};
// Zoltan's guess evaluated exactly (the real IEEE sequence in FP64): out of line, it is rarely needed
__device__ DDC_NOINLINE inline int guess_bin_exact(int vmin, int range, double T, unsigned wlo, unsigned den, int alo, int ahi)
{
    const double tmp = __dadd_rn((double)vmin, __dmul_rn(__ddiv_rn(__dsub_rn(T, (double)wlo), (double)den), (double)range));
    if (tmp < (double)alo)
        return alo - 1;
    if (tmp >= (double)ahi)
        return ahi;
    return (int)floor(tmp);
}
// guess_bin with the exact evaluation out of line; Wn / half: the target is Wn / 2 when half, else T
__device__ __forceinline__ int guess_bin_fast(int vmin, int range, unsigned Ti, float Tfrac, double T, bool half, unsigned Wn,
    unsigned wlo, unsigned den, int alo, int ahi)
{
    // As much weight set aside below as above (always so in the first iteration) with targets of W / 2: the
    // quotient (W/2 - wlo) / (W - 2 wlo) is 0.5 EXACTLY in IEEE double, so tmp = vmin + range / 2 without any
    // rounding and its floor is an integer expression.  (This case used to fall to the exact FP64 evaluation
    // whenever range was even: the FP32 estimate of an exact integer is never "sure".)
    if (half && Wn - wlo - wlo == den && den != 0u)
        return min(max(vmin + (range >> 1), alo - 1), ahi);
    if (range < (1 << 22) && den != 0u) {
        const float numf = (float)(Ti - wlo) + Tfrac;
        const float rangef = (float)range;
        const float off = __fdividef(numf, (float)den) * rangef;
        const float fl = floorf(off), fr = off - fl;
        const float guard = rangef * 0x1p-20f + 0x1p-18f;
        if (fr > guard && fr < 1.0f - guard)
            return min(max(vmin + (int)fl, alo - 1), ahi);
    }
    return guess_bin_exact(vmin, range, half ? 0.5 * (double)Wn : T, wlo, den, alo, ahi);
}
constexpr int FAST_HIST_BINS = 32768;
// H must have a bit map and at most FAST_HIST_BINS bins; call after the histogram is complete
__device__ __forceinline__ FastHist make_fast_hist(const Hist& H)
{
    FastHist F;
    F.pfx = sm_addr(H.pfx);
    F.l0 = sm_addr(H.l0);
    F.l1 = sm_addr(H.l1);
    F.l2 = H.l2[0];
    return F;
}
// largest non-empty bin in [a, t], -1 if none
__device__ __forceinline__ int fast_prev(const FastHist& F, int a, int t)
{
    if (t < a)
        return -1;
    int w = t >> 5;
    unsigned m = sm_ld(F.l0, w) & (0xffffffffu >> (31 - (t & 31)));
    if (!m) {
        int w1 = w >> 5;
        unsigned m1 = sm_ld(F.l1, w1) & ((1u << (w & 31)) - 1u);
        if (!m1) {
            const unsigned m2 = F.l2 & ((1u << w1) - 1u);
            if (!m2)
                return -1;
            w1 = 31 - __clz(m2);
            m1 = sm_ld(F.l1, w1);
        }
        w = (w1 << 5) + 31 - __clz(m1);
        m = sm_ld(F.l0, w);
    }
    const int j = (w << 5) + 31 - __clz(m);
    return j >= a ? j : -1;
}
// smallest non-empty bin in [t, b], -1 if none
__device__ __forceinline__ int fast_next(const FastHist& F, int t, int b)
{
    if (t > b)
        return -1;
    int w = t >> 5;
    unsigned m = sm_ld(F.l0, w) & (0xffffffffu << (t & 31));
    if (!m) {
        int w1 = w >> 5;
        unsigned m1 = (w & 31) == 31 ? 0u : sm_ld(F.l1, w1) & (0xffffffffu << ((w & 31) + 1));
        if (!m1) {
            const unsigned m2 = w1 == 31 ? 0u : F.l2 & (0xffffffffu << (w1 + 1));
            if (!m2)
                return -1;
            w1 = __ffs(m2) - 1;
            m1 = sm_ld(F.l1, w1);
        }
        w = (w1 << 5) + __ffs(m1) - 1;
        m = sm_ld(F.l0, w);
    }
    const int j = (w << 5) + __ffs(m) - 1;
    return j <= b ? j : -1;
}
// Both neighbours of a guess at once: the largest non-empty bin in [a, t] and the smallest one in [t2, b].  The two
// l0 words are requested before either result is looked at -- a warp issues in order, so a branch on the first
// load would hold the second one back for a shared-memory round trip; a median is one dependent chain, and the
// round trips are what it is made of.
__device__ __forceinline__ void fast_prev_next(const FastHist& F, int a, int t, int t2, int b, int* prev, int* next)
{
    const bool hp = t >= a, hn = t2 <= b;
    const int wp = hp ? t >> 5 : 0, wn = hn ? t2 >> 5 : 0;
    unsigned mp = sm_ld(F.l0, wp), mn = sm_ld(F.l0, wn);
    mp = hp ? mp & (0xffffffffu >> (31 - (t & 31))) : 0u;
    mn = hn ? mn & (0xffffffffu << (t2 & 31)) : 0u;
    int p, n;
    if (mp)
        p = (wp << 5) + 31 - __clz(mp);
    else
        p = hp ? fast_prev(F, a, (wp << 5) - 1) : -1; // (continues below this word: l1 / l2)
    if (mn)
        n = (wn << 5) + __ffs(mn) - 1;
    else
        n = hn ? fast_next(F, (wn << 5) + 32, b) : -1;
    *prev = p >= a ? p : -1;
    *next = (n >= 0 && n <= b) ? n : -1;
}
__device__ inline int median_boundary_fast(const FastHist& F, int c0, int c1, int nlo, int num_parts, int* iters)
{
    const unsigned base = sm_ld(F.pfx, c0);
    const unsigned Wn = c1 < c0 ? 0u : sm_ld(F.pfx, c1 + 1) - base;
    if (Wn == 0u) { // no dot at all: integer midpoint of the inherited range (policy Q2)
        *iters += 1;
        return c0 + ((c1 + 1 - c0) >> 1);
    }
    const bool half = 2 * nlo == num_parts;
    double T = 0.0, Thi = 0.0;
    unsigned Ti, ceilT, ceilThi;
    float Tfrac;
    if (half) { // fractionlo = 0.5: both targets are W / 2 exactly
        Ti = Wn >> 1;
        Tfrac = (Wn & 1u) ? 0.5f : 0.0f;
        ceilT = ceilThi = (Wn + 1u) >> 1;
    } else {
        T = __dmul_rn(__ddiv_rn((double)nlo, (double)num_parts), (double)Wn);
        Thi = __dsub_rn((double)Wn, T);
        const double fT = floor(T);
        Ti = (unsigned)fT;
        Tfrac = (float)(T - fT);
        ceilT = (unsigned)ceil(T);
        ceilThi = (unsigned)ceil(Thi);
    }
    int first, last;
    fast_prev_next(F, c0, c1, c0, c1, &last, &first);
    int vmin = first, vmax = last, alo = first, ahi = last;
    int B;
    unsigned wlo = 0, whi = 0;
    int it = 0;
    for (;;) {
        const int t = guess_bin_fast(vmin, vmax - vmin, Ti, Tfrac, T, half, Wn, wlo, Wn - wlo - whi, alo, ahi);
        it++;
        B = t;
        // everything that depends on the guess only is requested together: the dots up to t and BOTH neighbours of
        // the guess (which one is needed depends on the dots up to t) ...
        const unsigned cum = sm_ld(F.pfx, t + 1) - base; // dots in [c0, t] = weightlo + totallo
        int vlo, vhi;
        fast_prev_next(F, alo, t, t + 1, ahi, &vlo, &vhi);
        // ... and so are the dots up to both neighbours
        const unsigned p_hi = sm_ld(F.pfx, (vhi >= 0 ? vhi : t) + 1), p_lo = sm_ld(F.pfx, vlo >= 0 ? vlo : c0);
        if (cum < ceilT) { // lower half TOO SMALL
            if (vhi < 0)
                break;
            const unsigned moved = p_hi - base; // weightlo + wthi
            if (moved >= ceilT) { // the bin that crosses the target: Zoltan's tie rules
                bool move;
                if (half) // (moved - W/2) < (W/2 - cum) <=> moved + cum < W: exact in integers as it was in FP64
                    move = moved - cum == 1u ? moved + cum < Wn : !(moved + cum > Wn);
                else {
                    const double over = __dsub_rn((double)moved, T), under = __dsub_rn(T, (double)cum);
                    move = moved - cum == 1u ? over < under : !(over > under);
                }
                if (move)
                    B = vhi;
                break;
            }
            B = vhi;
            wlo = moved;
            if (moved >= ceilT - 1u)
                break;
            vmin = vhi;
            alo = vhi + 1;
        } else {
            const unsigned above = Wn - cum; // dots in (t, c1] = weighthi + totalhi
            if (above >= ceilThi)
                break;
            if (vlo < 0)
                break;
            const unsigned moved = Wn - (p_lo - base); // weighthi + wtlo = dots in [vlo, c1]
            if (moved >= ceilThi) {
                bool move;
                if (half)
                    move = moved - above == 1u ? moved + above < Wn : !(moved + above > Wn);
                else {
                    const double over = __dsub_rn((double)moved, Thi), under = __dsub_rn(Thi, (double)above);
                    move = moved - above == 1u ? over < under : !(over > under);
                }
                if (move)
                    B = vlo - 1;
                break;
            }
            B = vlo - 1;
            whi = moved;
            if (moved >= ceilThi - 1u)
                break;
            vmax = vlo;
            ahi = vlo - 1;
        }
    }
    *iters += it;
    int L, U;
    fast_prev_next(F, c0, B, B + 1, c1, &L, &U);
    if (L >= 0 && U >= 0)
        return (L + U + 1) >> 1;
    if (L >= 0)
        return L + 1;
    return U;
}

// The RCB recursion without level barriers.  A set is the cell range [lo, hi) along the cut
// dimension plus the parts [plo, plo + n) it still has to produce; a set with n > 1 is split by the
// weighted median into a lower child with ceil(n / 2) parts (Zoltan_Divide_Machine) and an upper
// child with the rest, a set with n == 1 is final.  Because a set of n parts has min(n, 2^l) leaves
// l levels further down (the part counts of one level differ by at most one), the path from the root
// to the k-th leaf can be walked WITHOUT knowing the other branches: every thread walks the path
// of its own leaf and evaluates the medians along it.  Threads whose paths share a set evaluate the
// same median redundantly (same instructions, same result), nothing is exchanged, and no thread
// waits for the slowest median of a level -- Zoltan's interpolation search needs 2-3 iterations on
// average but tens for an occasional set, and with per-level barriers every level paid for its worst.
struct RcbSet {
    int lo, hi, plo, n;
};
__device__ __forceinline__ int leaves_below(int n, int levels)
{
    return levels >= 31 ? n : min(n, 1 << levels);
}
#ifndef DDC_HOST_EMU
__device__ __forceinline__ unsigned long long walk_clock()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#else
inline unsigned long long walk_clock() { return 0ull; }
#endif
// walk `levels` levels down from `set` towards leaf number k (0-based among the leaves below
// `set`); iterations of a median are counted by the thread whose leaf is the first of its upper child.
// lvl_ts (diagnostics, normally nullptr): a time stamp after every level.
__device__ inline RcbSet rcb_walk(const Hist& H, RcbSet set, int levels, int k, int* iters,
    unsigned long long* lvl_ts = nullptr)
{
    for (int l = levels; l > 0 && set.n > 1; l--) {
        if (lvl_ts && levels - l < 8)
            lvl_ts[levels - l] = walk_clock(); // start of level (levels - l)
        const int nlo = (set.n - 1) / 2 + 1;
        int it = 0;
        const int cut = median_boundary(H, set.lo, set.hi - 1, nlo, set.n, &it);
        const int below = leaves_below(nlo, l - 1);
        if (k < below) {
            set.hi = cut;
            set.n = nlo;
        } else {
            if (k == below)
                *iters += it;
            k -= below;
            set.lo = cut;
            set.plo += nlo;
            set.n -= nlo;
        }
    }
    return set;
}
} // namespace ddc
