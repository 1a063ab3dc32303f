// Host check of rlap_b200/csrc/introsort.cuh against the real thing: std::sort of libstdc++ with the reference's
// comparators (preconditioner.cc:295-303). Prints "ok <cases> <heap sorts seen>" or the first mismatch.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <random>
#include <vector>

static long g_heap = 0;
#define RLAP_IS_HEAP_HOOK g_heap++
#include "../../rlap_b200/csrc/introsort.cuh"

struct El { uint64_t q; int id; };

// the warp primitives of introsort_loop_arrange_warp with the 32 lanes run one after the other
struct HostWarp {
    template <class F> unsigned ballot(F f) { unsigned m = 0; for (int l = 0; l < 32; l++) if (f(l)) m |= 1u << l; return m; }
    template <class F> void each(F f) { for (int l = 0; l < 32; l++) f(l); }
    void sync() {}
};

template <bool DESC>
static bool check(const std::vector<uint64_t>& q) {
    const int n = (int)q.size();
    std::vector<El> want(n);
    for (int i = 0; i < n; i++) want[i] = El{q[i], i};
    if (DESC) std::sort(want.begin(), want.end(), [](const El& a, const El& b) { return a.q > b.q; });
    else std::sort(want.begin(), want.end(), [](const El& a, const El& b) { return a.q < b.q; });
    std::vector<uint64_t> key(q);
    std::vector<uint32_t> tag(n);
    std::iota(tag.begin(), tag.end(), 0u);
    rlap::introsort_loop_arrange<DESC, uint64_t, uint32_t>(key.data(), tag.data(), n);
    // the caller's part: stable sort by key of the arrangement
    std::vector<int> pos(n);
    std::iota(pos.begin(), pos.end(), 0);
    std::stable_sort(pos.begin(), pos.end(), [&](int a, int b) { return DESC ? key[a] > key[b] : key[a] < key[b]; });
    for (int i = 0; i < n; i++) {
        if ((int)tag[pos[i]] != want[i].id || key[pos[i]] != want[i].q) {
            printf("mismatch n=%d desc=%d at %d: got id %u want %d\n", n, (int)DESC, i, tag[pos[i]], want[i].id);
            return false;
        }
    }
    // the warp version leaves the very same arrangement (not just the same sorted order)
    std::vector<uint64_t> key2(q);
    std::vector<uint32_t> tag2(n);
    std::iota(tag2.begin(), tag2.end(), 0u);
    HostWarp wp;
    rlap::introsort_loop_arrange_warp<DESC, uint64_t, uint32_t, HostWarp>(wp, key2.data(), tag2.data(), n);
    if (tag2 != tag || key2 != key) {
        int i = 0;
        while (tag2[i] == tag[i]) i++;
        printf("warp version differs n=%d desc=%d at %d: %u vs %u\n", n, (int)DESC, i, tag2[i], tag[i]);
        return false;
    }
    // register tiles (at most 32 neighbours): 8-bit keys = number of smaller weights, 8-bit tags
    if (n <= 32) {
        uint8_t k8[32], t8[32];
        for (int i = 0; i < n; i++) {
            int c = 0;
            for (int j = 0; j < n; j++) c += q[j] < q[i];
            k8[i] = (uint8_t)c; t8[i] = (uint8_t)i;
        }
        rlap::introsort_loop_arrange_warp<DESC, uint8_t, uint8_t, HostWarp>(wp, k8, t8, n);
        for (int i = 0; i < n; i++)
            if (t8[i] != tag[i]) { printf("8-bit version differs n=%d desc=%d at %d\n", n, (int)DESC, i); return false; }
    }
    return true;
}

// Musser's median-of-three killer: drives the partition loop into its depth limit (heap sort branch)
static std::vector<uint64_t> killer(int n) {
    std::vector<uint64_t> v(n);
    int k = n / 2;
    for (int i = 1; i <= k; i++) {
        if (i % 2 == 1) { v[i - 1] = i; v[i] = k + i; }
        v[k + i - 1] = 2 * i;
    }
    return v;
}

int main() {
    std::mt19937_64 gen(12345);
    for (int rep = 0; rep < 200000; rep++) {         // is_nth_bit against the bit-by-bit definition
        unsigned m = (unsigned)gen();
        if (rep % 3 == 0) m &= (unsigned)gen();
        if (m == 0) continue;
        int n = (int)(gen() % (unsigned)__builtin_popcount(m));
        unsigned t = m;
        for (int i = 0; i < n; i++) t &= t - 1;
        if (rlap::is_nth_bit(m, n) != __builtin_ffs((int)t) - 1) { printf("is_nth_bit(%08x, %d)\n", m, n); return 1; }
    }
    for (int n = 17; n <= 32; n++) {                 // the tile sizes, many inputs each
        for (int rep = 0; rep < 2000; rep++) {
            std::vector<uint64_t> q(n);
            const uint64_t span = 1 + rep % 6;
            for (int i = 0; i < n; i++) q[i] = (rep % 7 == 6) ? gen() : gen() % span;
            if (!check<false>(q) || !check<true>(q)) return 1;
        }
    }
    long cases = 0;
    const int sizes[] = {1, 2, 15, 16, 17, 18, 23, 31, 32, 33, 47, 64, 65, 66, 67, 80, 95, 96, 97, 98, 99, 100, 127, 128, 129, 130, 131, 160, 161, 255, 500, 1000, 1024, 2047, 2900, 5000, 20000, 70000};
    for (int n : sizes) {
        for (int rep = 0; rep < (n <= 1024 ? 400 : 12); rep++) {
            std::vector<uint64_t> q(n);
            const int mode = rep % 8;
            for (int i = 0; i < n; i++) {
                switch (mode) {
                    case 0: q[i] = 1ull << 40; break;                              // unit weights: all equal
                    case 1: q[i] = gen() % 2; break;
                    case 2: q[i] = gen() % 3; break;
                    case 3: q[i] = gen() % 7; break;
                    case 4: q[i] = gen(); break;                                  // all distinct
                    case 5: q[i] = (gen() % 10 < 8) ? (1ull << 40) : (gen() >> 8); break;  // mostly unit, some fills
                    case 6: q[i] = (uint64_t)i / 4; break;                        // ascending runs
                    default: q[i] = (uint64_t)(n - i) / 3; break;                 // descending runs
                }
            }
            if (!check<false>(q) || !check<true>(q)) return 1;
            cases += 2;
        }
    }
    for (int n = 17; n <= 420; n++) {               // every size around the chunk boundaries of the warp version
        for (int rep = 0; rep < 24; rep++) {
            std::vector<uint64_t> q(n);
            const uint64_t span = (rep % 4 == 0) ? 1 : (rep % 4 == 1) ? 2 : (rep % 4 == 2) ? 5 : (1ull << 50);
            for (int i = 0; i < n; i++) q[i] = gen() % span;
            if (!check<false>(q) || !check<true>(q)) return 1;
            cases += 2;
        }
    }
    for (int n : {64, 200, 1000, 4096, 30000}) {
        std::vector<uint64_t> q = killer(n);
        if (!check<false>(q)) return 1;
        std::vector<uint64_t> r(q);
        for (auto& x : r) x = ~x;
        if (!check<true>(r)) return 1;
        cases += 2;
    }
    printf("ok %ld %ld\n", cases, g_heap);
    return 0;
}
