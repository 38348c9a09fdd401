// emu_median_harness.cpp -- TEST INFRASTRUCTURE.  The scalar core of the CUDA cut kernels
// (domain_decomp_b200/csrc/ddc_median.cuh: median_boundary, the bit map queries, rcb_walk) compiled as
// host C++ (DDC_HOST_EMU) and fuzzed against the oracle's literal double-precision Zoltan loop
// (ddc_oracle.c: find_median_hist via orc_median_boundary) on random histograms.
//
// Why: the device code is NOT a transliteration of the oracle.  It keeps Zoltan's weights as integers,
// compares against ceil(target), replaces the tolerance test by an integer test, searches non-empty
// bins through a three-level bit map and forms the interpolated guess in FP32 with a guard band,
// falling back to the exact FP64 sequence near integers.  Every one of those steps must give the same
// boundary AND the same iteration count as the oracle for every histogram; the GPU parity tests check
// that on whole decompositions, this harness checks it on tens of millions of single medians, with the
// fast division perturbed by +-2 ulp (the documented accuracy of __fdividef).
#define DDC_HOST_EMU
#include "ddc_median.cuh"
#include "ddc_neighbours.cuh"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

// the execution model of a single "thread": no block, no warp
int ddc_emu_fdividef_ulps = 0;
ddc_emu_dim threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
void __syncthreads() { }
int __syncthreads_or(int p) { return p; }
const unsigned long long* ddc_emu_warp_gather(unsigned, unsigned long long v, unsigned* present)
{
    static unsigned long long lanes[32];
    for (auto& l : lanes)
        l = 0ull; // the other lanes of the warp contribute nothing
    lanes[0] = v;
    *present = 1u;
    return lanes;
}
void* ddc_emu_dyn_smem() { return nullptr; }

extern "C" int orc_median_boundary(const int64_t* pfx, int n, int c0, int c1, int nlo, int num_parts, long* iters);
extern "C" void orc_neighbours(const int32_t* boxes, int P, int NX, int NY, int px, int py, int32_t* counts,
    const int64_t* offsets, int32_t* ids, int32_t* halos, int32_t* starts);

namespace {
struct Rng {
    uint64_t s;
    uint64_t next()
    {
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    uint64_t below(uint64_t n) { return n ? next() % n : 0; }
    double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

// the three-level bit map exactly as block_prefix_tiles lays it out behind the prefix sums
void build_bitmap(const std::vector<unsigned>& h, std::vector<unsigned>& bm)
{
    const int n = (int)h.size();
    const int tiles = (n + ddc::HIST_TILE - 1) / ddc::HIST_TILE;
    bm.assign(ddc::hist_bitmap_words(n), 0u);
    unsigned* l0 = bm.data();
    unsigned* l1 = bm.data() + (size_t)tiles * 1024;
    unsigned* l2 = bm.data() + (size_t)tiles * (1024 + 32);
    for (int i = 0; i < n; i++)
        if (h[i])
            l0[i >> 5] |= 1u << (i & 31);
    for (int w = 0; w < tiles * 1024; w++)
        if (l0[w])
            l1[w >> 5] |= 1u << (w & 31);
    for (int w = 0; w < tiles * 32; w++)
        if (l1[w])
            l2[w >> 5] |= 1u << (w & 31);
}

void make_histogram(Rng& r, int shape, int nmax, std::vector<unsigned>& h)
{
    const int n = 1 + (int)r.below((uint64_t)nmax);
    h.assign(n, 0u);
    // the product's dots are mask cells: at most INT_MAX of them (ddc_set_mask_* refuses larger masks),
    // which is what lets the device keep 32-bit prefix sums
    const unsigned cap = (unsigned)std::max<long long>(1, 2147483647LL / n);
    struct Clamp {
        std::vector<unsigned>& h;
        unsigned cap;
        ~Clamp()
        {
            for (auto& v : h)
                v = std::min(v, cap);
        }
    } clamp { h, cap };
    switch (shape) {
    case 0: { // dense, counts up to a random scale
        const unsigned scales[] = { 1, 2, 3, 7, 100, 4096, 40000 };
        const unsigned m = scales[r.below(7)];
        for (auto& v : h)
            v = (unsigned)r.below(m + 1);
        break;
    }
    case 1: { // sparse
        const double p = 0.002 + 0.3 * r.unit() * r.unit();
        for (auto& v : h)
            if (r.unit() < p)
                v = 1 + (unsigned)r.below(1 + r.below(500));
        break;
    }
    case 2: { // ties: the same count in every bin, a few gaps
        const unsigned k = 1 + (unsigned)r.below(6);
        for (auto& v : h)
            v = k;
        for (int g = (int)r.below(4); g > 0; g--) {
            const int a = (int)r.below(n), len = 1 + (int)r.below(1 + n / 8);
            for (int i = a; i < n && i < a + len; i++)
                h[i] = 0;
        }
        break;
    }
    case 3: // single dots
        for (auto& v : h)
            v = r.unit() < 0.5 ? 1u : 0u;
        break;
    default: { // coastline-like: runs of land, runs of slowly varying ocean counts
        int i = 0;
        while (i < n) {
            const int len = 1 + (int)r.below(1 + n / 6);
            const bool land = r.unit() < 0.4;
            unsigned c = 1 + (unsigned)r.below(3000);
            for (int k = 0; k < len && i < n; k++, i++) {
                if (!land) {
                    c = (unsigned)std::max<long long>(0, (long long)c + (long long)r.below(41) - 20);
                    h[i] = c;
                }
            }
        }
    }
    }
}

// the leaves of the RCB tree of [lo, hi) x [plo, plo + n) after `levels` levels, via the ORACLE's median
void oracle_leaves(const std::vector<int64_t>& pfx, int nbins, int lo, int hi, int plo, int n, int levels,
    std::vector<ddc::RcbSet>& out, long* iters)
{
    if (levels == 0 || n <= 1) {
        out.push_back({ lo, hi, plo, n });
        return;
    }
    const int nlo = (n - 1) / 2 + 1;
    long it = 0;
    const int cut = orc_median_boundary(pfx.data(), nbins, lo, hi - 1, nlo, n, &it);
    *iters += it;
    oracle_leaves(pfx, nbins, lo, cut, plo, nlo, levels - 1, out, iters);
    oracle_leaves(pfx, nbins, cut, hi, plo + nlo, n - nlo, levels - 1, out, iters);
}
} // namespace

extern "C" {
// returns the number of mismatches; bad[0..9] describes the first one:
// {kind (1 median, 2 walk), n, c0, c1, nlo, num_parts, got, want, got_iters, want_iters}
__attribute__((visibility("default"))) long emu_fuzz(uint64_t seed, long histograms, int queries, int nmax, int shape,
    int ulps, int use_bitmap, long long* bad, long long* medians_checked)
{
    ddc_emu_fdividef_ulps = ulps;
    Rng r { seed * 0x2545F4914F6CDD1Dull + 1 };
    long mism = 0;
    long long checked = 0;
    std::vector<unsigned> h, pfx, bm;
    std::vector<int64_t> pfx64;
    for (long t = 0; t < histograms; t++) {
        make_histogram(r, shape, nmax, h);
        const int n = (int)h.size();
        pfx.assign((size_t)n + 1, 0u);
        pfx64.assign((size_t)n + 1, 0);
        for (int i = 0; i < n; i++) {
            pfx[i + 1] = pfx[i] + h[i];
            pfx64[i + 1] = pfx64[i] + h[i];
        }
        if (use_bitmap)
            build_bitmap(h, bm);
        const ddc::Hist H = ddc::make_hist(pfx.data(), use_bitmap ? bm.data() : nullptr, n);
        for (int q = 0; q < queries; q++) {
            int c0 = (int)r.below(n), c1 = (int)r.below(n);
            if (r.unit() < 0.3) { // whole histogram, as at the root
                c0 = 0;
                c1 = n - 1;
            }
            if (c1 < c0)
                std::swap(c0, c1);
            const int pow2[] = { 2, 4, 8, 16, 64, 1024, 16384 };
            const int np = r.unit() < 0.5 ? pow2[r.below(7)] : 2 + (int)r.below(63);
            const int nlo = r.unit() < 0.8 ? (np - 1) / 2 + 1 : 1 + (int)r.below(np - 1);
            int it = 0;
            const int got = ddc::median_boundary(H, c0, c1, nlo, np, &it);
            long wit = 0;
            const int want = orc_median_boundary(pfx64.data(), n, c0, c1, nlo, np, &wit);
            checked++;
            if (got != want || it != wit) {
                if (!mism) {
                    const long long b[10] = { 1, n, c0, c1, nlo, np, got, want, it, wit };
                    std::memcpy(bad, b, sizeof b);
                }
                mism++;
            }
            if (use_bitmap && n <= ddc::FAST_HIST_BINS) { // the shared-memory variant of the cut kernels
                const ddc::FastHist F = ddc::make_fast_hist(H);
                int fit = 0;
                const int fgot = ddc::median_boundary_fast(F, c0, c1, nlo, np, &fit);
                checked++;
                if (fgot != want || fit != wit) {
                    if (!mism) {
                        const long long b[10] = { 3, n, c0, c1, nlo, np, fgot, want, fit, wit };
                        std::memcpy(bad, b, sizeof b);
                    }
                    mism++;
                }
            }
        }
        // the barrier-free walk: every leaf reached on its own must be the oracle's recursion
        {
            const int P = 1 + (int)r.below(40), levels = (int)r.below(7);
            std::vector<ddc::RcbSet> want;
            long wit = 0;
            oracle_leaves(pfx64, n, 0, n, 0, P, levels, want, &wit);
            const int leaves = ddc::leaves_below(P, levels);
            long git = 0;
            bool ok = leaves == (int)want.size();
            for (int k = 0; ok && k < leaves; k++) {
                int it = 0;
                const ddc::RcbSet g = ddc::rcb_walk(H, { 0, n, 0, P }, levels, k, &it);
                git += it;
                ok = g.lo == want[k].lo && g.hi == want[k].hi && g.plo == want[k].plo && g.n == want[k].n;
            }
            ok = ok && git == wit; // every median's iterations are counted exactly once
            checked += (long long)want.size();
            if (!ok) {
                if (!mism) {
                    const long long b[10] = { 2, n, P, levels, leaves, (long long)want.size(), git, wit, 0, 0 };
                    std::memcpy(bad, b, sizeof b);
                }
                mism++;
            }
        }
    }
    *medians_checked = checked;
    return mism;
}

// K7's structured search (csrc/ddc_neighbours.cuh, one "thread" per (list, part)) against the oracle's literal
// O(P^2) discovery on random tilings: x-sorted strips of y-sorted parts.  degenerate: boundaries may coincide
// (zero-width strips, zero-height parts -- what the cut kernels produce when there are more parts than
// non-empty columns or rows).  bad[0..5] = {NX, NY, P, list, part, what (1 count, 2 entry, 3 edge cut)}.
__attribute__((visibility("default"))) long emu_fuzz_neighbours(uint64_t seed, long cases, int maxn, int maxstrips,
    int maxparts, int degenerate, long long* bad, long long* lists_checked)
{
    Rng r { seed * 0x9E3779B97F4A7C15ull + 7 };
    long mism = 0;
    long long checked = 0;
    auto boundaries = [&](int n, int extent, std::vector<int>& out) { // n intervals tiling [0, extent)
        out.assign((size_t)n + 1, 0);
        for (int i = 1; i < n; i++)
            out[i] = (int)r.below((uint64_t)extent + 1);
        out[n] = extent;
        std::sort(out.begin(), out.end());
        if (!degenerate) { // strictly increasing: possible only when n <= extent
            for (int i = 1; i <= n; i++)
                out[i] = std::max(out[i], out[i - 1] + 1);
            for (int i = n; i >= 1; i--)
                out[i] = std::min(out[i], extent - (n - i));
        }
    };
    for (long t = 0; t < cases; t++) {
        const int NX = 1 + (int)r.below((uint64_t)maxn), NY = 1 + (int)r.below((uint64_t)maxn);
        int S = 1 + (int)r.below((uint64_t)maxstrips);
        if (!degenerate)
            S = std::min(S, NX);
        std::vector<int> xb, yb;
        boundaries(S, NX, xb);
        std::vector<int> sx0(S), sx1(S), p0(S + 1), bx0, by0, bex, bey;
        for (int s = 0; s < S; s++) {
            int n = 1 + (int)r.below((uint64_t)maxparts);
            if (!degenerate)
                n = std::min(n, NY);
            boundaries(n, NY, yb);
            sx0[s] = xb[s];
            sx1[s] = xb[s + 1];
            p0[s] = (int)bx0.size();
            for (int j = 0; j < n; j++) {
                bx0.push_back(xb[s]);
                bex.push_back(xb[s + 1] - xb[s]);
                by0.push_back(yb[j]);
                bey.push_back(yb[j + 1] - yb[j]);
            }
        }
        const int P = (int)bx0.size();
        p0[S] = P;
        const int px = (int)r.below(2), py = (int)r.below(2);
        int Sv = S, always = 0;
        const ddc::StripTable st { sx0.data(), sx1.data(), p0.data(), &Sv, &always };
        const ddc::BoxTable bx { bx0.data(), by0.data(), bex.data(), bey.data() };
        // the oracle
        std::vector<int32_t> aos(4 * (size_t)P), wc(8 * (size_t)P);
        for (int p = 0; p < P; p++) {
            aos[4 * p] = bx0[p];
            aos[4 * p + 1] = by0[p];
            aos[4 * p + 2] = bex[p];
            aos[4 * p + 3] = bey[p];
        }
        orc_neighbours(aos.data(), P, NX, NY, px, py, wc.data(), nullptr, nullptr, nullptr, nullptr);
        int64_t woff[9] = { 0 };
        for (int l = 0; l < 8; l++) {
            int64_t tot = 0;
            for (int p = 0; p < P; p++)
                tot += wc[(size_t)l * P + p];
            woff[l + 1] = woff[l] + tot;
        }
        std::vector<int32_t> wi(woff[8] + 1), wh(woff[8] + 1), ws(woff[8] + 1);
        orc_neighbours(aos.data(), P, NX, NY, px, py, wc.data(), woff, wi.data(), wh.data(), ws.data());
        // the device code: count, exclusive scan per list, fill
        int cap = 1;
        for (int l = 0; l < 8; l++)
            cap = std::max<int>(cap, (int)(woff[l + 1] - woff[l]) + 8);
        std::vector<int> gc(8 * (size_t)P, -1), goff(8 * ((size_t)P + 1), 0), gi(8 * (size_t)cap, -1), gh(8 * (size_t)cap, -1),
            gs(8 * (size_t)cap, -1);
        ddc::DevScalars sc {};
        for (int l = 0; l < 8; l++)
            for (int me = 0; me < P; me++)
                ddc::neighbours_structured<false>(bx, P, NX, NY, px, py, st, l, me, gc.data(), nullptr, cap, nullptr, nullptr,
                    nullptr, &sc);
        bool ok = true;
        int badl = -1, badp = -1, what = 0;
        for (int l = 0; ok && l < 8; l++) {
            int run = 0;
            for (int me = 0; me < P; me++) {
                goff[(size_t)l * (P + 1) + me] = run;
                if (gc[(size_t)l * P + me] != wc[(size_t)l * P + me]) {
                    ok = false;
                    badl = l;
                    badp = me;
                    what = 1;
                    break;
                }
                run += gc[(size_t)l * P + me];
            }
            goff[(size_t)l * (P + 1) + P] = run;
        }
        if (ok) {
            for (int l = 0; l < 8; l++)
                for (int me = 0; me < P; me++)
                    ddc::neighbours_structured<true>(bx, P, NX, NY, px, py, st, l, me, gc.data(), goff.data(), cap, gi.data(),
                        gh.data(), gs.data(), &sc);
            unsigned long long cut = 0;
            for (int l = 0; ok && l < 8; l++) {
                const int n = (int)(woff[l + 1] - woff[l]);
                for (int k = 0; k < n; k++) {
                    const size_t g = (size_t)l * cap + k, w = (size_t)woff[l] + k;
                    if (gi[g] != wi[w] || gh[g] != wh[w] || gs[g] != ws[w]) {
                        ok = false;
                        badl = l;
                        badp = k;
                        what = 2;
                        break;
                    }
                    if (l < 4)
                        cut += (unsigned long long)wh[w];
                }
            }
            if (ok && cut != sc.edge_cut) {
                ok = false;
                what = 3;
            }
        }
        checked += 8;
        if (!ok) {
            if (!mism) {
                const long long b[6] = { NX, NY, P, badl, badp, what };
                std::memcpy(bad, b, sizeof b);
            }
            mism++;
        }
    }
    *lists_checked = checked;
    return mism;
}
}
