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

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace ddc {
int g_fdividef_ulps = 0;
}

extern "C" int orc_median_boundary(const int64_t* pfx, int n, int c0, int c1, int nlo, int num_parts, long* iters);

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
    ddc::g_fdividef_ulps = ulps;
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
}
