// ddc_neighbours.cuh -- the scalar part of the neighbour / halo kernel K7: the reference's edge tests
// (DomainUtils.cpp:15-35, Partitioner.cpp:20-80) and the structured search over x-sorted strips of
// y-sorted parts, one thread per (list, part).  Like ddc_median.cuh this file compiles as plain host C++
// too (define DDC_HOST_EMU; oracle/emu_median_harness.cpp fuzzes it against the oracle's literal O(P^2)
// discovery on random and degenerate tilings).  Included by ddc_kernels.cuh.
#pragma once
#include <stdint.h>

#ifdef DDC_HOST_EMU
#include "ddc_host_emu.h" // oracle/emu: host stand-ins of the device language (test builds only)
#endif

namespace ddc {

struct DevScalars {
    int changes; // any ocean cell OF THIS RANK whose RCB part differs from its naive block
    int changes_all; // the same over all ranks (written by K5)
    int overflow; // neighbour lists did not fit their capacity
    unsigned long long edge_cut; // sum of interior halo lengths
};

struct StripTable {
    int* x0; // [cap]
    int* x1; // [cap]
    int* p0; // [cap + 1]  first part of every strip, p0[S] = P
    int* S; // [1]
    int* always; // [1] != 0: no x structure, treat every strip as relevant (brute force)
};
struct BoxTable { // final boxes, SoA
    int* x0;
    int* y0;
    int* ex;
    int* ey;
};

struct Dom {
    int x1, y1, x2, y2;
};
// DomainUtils.cpp:15-35
__device__ __forceinline__ int domain_overlap(const Dom& d1, const Dom& d2, int edge)
{
    int overlap = 0;
    if (edge >= 2) { // BOTTOM, TOP: overlap along x
        if (d1.x2 >= d2.x1 && d1.x1 <= d2.x2)
            overlap = min(d1.x2, d2.x2) - max(d1.x1, d2.x1);
    } else { // LEFT, RIGHT: overlap along y
        if (d1.y2 >= d2.y1 && d1.y1 <= d2.y2)
            overlap = min(d1.y2, d2.y2) - max(d1.y1, d2.y1);
    }
    return overlap;
}
// Partitioner.cpp:20-53
__device__ __forceinline__ bool is_neighbour(const Dom& d1, const Dom& d2, int edge, bool is_px,
    bool is_py, int NX, int NY)
{
    if (edge == 3)
        return is_py ? d1.y2 == d2.y1 + NY : d1.y2 == d2.y1;
    if (edge == 2)
        return is_py ? d1.y1 == d2.y2 - NY : d1.y1 == d2.y2;
    if (edge == 0)
        return is_px ? d1.x1 == d2.x2 - NX : d1.x1 == d2.x2;
    return is_px ? d1.x2 == d2.x1 + NX : d1.x2 == d2.x1;
}
// Partitioner.cpp:55-80
__device__ __forceinline__ int halo_start(const Dom& d1, const Dom& d2, int edge)
{
    const int w2 = d2.x2 - d2.x1, h2 = d2.y2 - d2.y1;
    if (edge == 3)
        return max(d1.x1, d2.x1) - d2.x1;
    if (edge == 2)
        return (h2 - 1) * w2 + (max(d1.x1, d2.x1) - d2.x1);
    const int dy = max(d1.y1, d2.y1) - d2.y1;
    if (edge == 0)
        return (dy + 1) * w2 - 1;
    return dy * w2;
}

// The structured search (the normal case): the boxes are vertical strips, each holding y-sorted parts
// that tile [0, NY).  One THREAD per (list, part): binary-search the strips whose x range satisfies the
// edge's x condition (exact at strip level, every part of a strip has the strip's x range), binary-search
// the first part whose y range can match and scan the short run that does.  ids come out ascending by
// construction.  list l = periodic * 4 + edge; counts/offsets [8][P]; ids/halos/starts [8][cap].
template <bool FILL>
__device__ __forceinline__ void neighbours_structured(const BoxTable& bx, int P, int NX, int NY, int px, int py,
    const StripTable& st, int l, int me, int* __restrict__ counts, const int* __restrict__ offsets, int cap,
    int* __restrict__ ids, int* __restrict__ halos, int* __restrict__ starts, DevScalars* sc)
{
    const int per = l >> 2, edge = l & 3;
    const bool lr = edge < 2;
    int cnt = 0;
    unsigned long long cut = 0;
    const bool listed = me < P && (!per || (lr && px) || (!lr && py)); // get_neighbour_info_periodic's filter
    if (listed) {
        Dom d1;
        d1.x1 = bx.x0[me];
        d1.y1 = bx.y0[me];
        d1.x2 = d1.x1 + bx.ex[me];
        d1.y2 = d1.y1 + bx.ey[me];
        const bool wx = per && px, wy = per && py;
        const int base = FILL ? offsets[l * (P + 1) + me] : 0;
        const int S = *st.S;
        // strips are x-sorted and tile [0, NX): binary-search the first strip that can satisfy the
        // edge's x condition, then walk the (short) run of strips that do
        //   LEFT   strips ending   at xt = d1.x1 (+ NX)      RIGHT  strips starting at xt = d1.x2 (- NX)
        //   BOTTOM / TOP  strips with a positive x overlap: the first one ending after d1.x1
        const int xt = edge == 0 ? (wx ? d1.x1 + NX : d1.x1) : (wx ? d1.x2 - NX : d1.x2);
        int s = 0, hi = S;
        while (s < hi) {
            const int mid = (s + hi) >> 1;
            const bool ge = edge == 0 ? st.x1[mid] >= xt : (edge == 1 ? st.x0[mid] >= xt : st.x1[mid] > d1.x1);
            if (ge)
                hi = mid;
            else
                s = mid + 1;
        }
        for (; s < S; s++) {
            const int sx0 = st.x0[s], sx1 = st.x1[s];
            if (edge == 0 ? sx1 != xt : (edge == 1 ? sx0 != xt : sx0 >= d1.x2))
                break;
            if (!lr && !(d1.x2 >= sx0 && d1.x1 <= sx1 && min(d1.x2, sx1) - max(d1.x1, sx0) > 0))
                continue;
            const int q0 = st.p0[s], q1 = st.p0[s + 1];
            // first part of the strip whose y range can satisfy the y condition (parts are y-sorted)
            const int yt = edge == 3 ? (wy ? d1.y2 - NY : d1.y2) : (wy ? d1.y1 + NY : d1.y1);
            int lo = q0, qh = q1;
            while (lo < qh) {
                const int mid = (lo + qh) >> 1;
                const int y1m = bx.y0[mid], y2m = y1m + bx.ey[mid];
                // LEFT/RIGHT: first part ending above my first row; TOP: first part starting at or
                // above yt (d1.y2 == d2.y1); BOTTOM: first part ending at or above yt (d1.y1 == d2.y2)
                const bool ge = lr ? y2m > d1.y1 : (edge == 3 ? y1m >= yt : y2m >= yt);
                if (ge)
                    qh = mid;
                else
                    lo = mid + 1;
            }
            for (int q = lo; q < q1; q++) {
                Dom d2;
                d2.x1 = sx0;
                d2.x2 = sx1;
                d2.y1 = bx.y0[q];
                d2.y2 = d2.y1 + bx.ey[q];
                if (lr ? d2.y1 >= d1.y2 : (edge == 3 ? d2.y1 > yt : d2.y2 > yt))
                    break; // past the run that can match
                if (!per && q == me)
                    continue;
                if (!is_neighbour(d1, d2, edge, wx, wy, NX, NY))
                    continue;
                const int halo = domain_overlap(d1, d2, edge);
                if (halo <= 0)
                    continue;
                if (FILL) {
                    const size_t pos = (size_t)l * cap + base + cnt;
                    ids[pos] = q;
                    halos[pos] = halo;
                    starts[pos] = halo_start(d1, d2, edge);
                    if (!per)
                        cut += (unsigned long long)halo;
                }
                cnt++;
            }
        }
    }
    if (!FILL) {
        if (me < P)
            counts[l * P + me] = cnt;
    } else {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            cut += __shfl_xor_sync(0xffffffffu, cut, o);
        if ((threadIdx.x & 31u) == 0 && cut)
            atomicAdd(&sc->edge_cut, cut);
    }
}

} // namespace ddc
