// introsort.cuh — where libstdc++'s std::sort leaves EQUAL keys.
//
// The reference orders the merged neighbours of a star with std::sort on the edge weight
// (preconditioner.cc:295-303 / 334-342), right after it sorted them by row (:275 / :314). Every call site of
// the reference uses unit weights, so ties are the rule, and which of two equal neighbours comes first decides
// which of them can be sampled by the other. std::sort is not stable: for more than 16 elements libstdc++
// (bits/stl_algo.h) runs __introsort_loop - median-of-three pivot moved to the front, __unguarded_partition,
// recursion on the right part, heap sort once 2 * floor(log2 n) levels are used up - until every part has at
// most 16 elements, then ONE stable insertion sort over the whole range (__final_insertion_sort). The result
// is therefore: the stable sort, by key, of the arrangement the partition loop leaves behind - a deterministic
// function of the id-ordered input. This file restates that loop for (key, tag) pairs held in two parallel
// arrays; the caller then sorts by (key, position after the loop). Up to 16 elements the loop does nothing
// and ties keep the id order.
//
// Sequential by nature (every swap depends on the scans before it): one thread runs it per star. It is
// plain C++ - tests/test_introsort.py compiles it for the host and pins it against std::sort itself.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RLAP_HD __host__ __device__
#else
#define RLAP_HD
#endif

namespace rlap {

// comp(x, y) of the reference's lambda on the fixed-point weights: asc `x < y`, desc `x > y`
template <bool DESC, class Key>
RLAP_HD inline bool is_before(Key x, Key y) { return DESC ? (x > y) : (x < y); }

template <class Key, class Tag>
RLAP_HD inline void is_swap(Key* key, Tag* tag, int i, int j) {
    const Key k = key[i]; key[i] = key[j]; key[j] = k;
    const Tag t = tag[i]; tag[i] = tag[j]; tag[j] = t;
}

// std::__adjust_heap followed by std::__push_heap on [first, first + len)
template <bool DESC, class Key, class Tag>
RLAP_HD inline void is_adjust_heap(Key* key, Tag* tag, int first, int hole, int len, Key vk, Tag vt) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (is_before<DESC, Key>(key[first + child], key[first + child - 1])) child--;
        key[first + hole] = key[first + child];
        tag[first + hole] = tag[first + child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        key[first + hole] = key[first + child - 1];
        tag[first + hole] = tag[first + child - 1];
        hole = child - 1;
    }
    int parent = (hole - 1) / 2;
    while (hole > top && is_before<DESC, Key>(key[first + parent], vk)) {
        key[first + hole] = key[first + parent];
        tag[first + hole] = tag[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    key[first + hole] = vk;
    tag[first + hole] = vt;
}

// std::__partial_sort(first, last, last): __make_heap, then __sort_heap (the __heap_select loop is empty)
template <bool DESC, class Key, class Tag>
RLAP_HD inline void is_heap_sort(Key* key, Tag* tag, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            is_adjust_heap<DESC, Key, Tag>(key, tag, first, parent, len, key[first + parent], tag[first + parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    int end = last;
    while (end - first > 1) {
        --end;
        const Key vk = key[end];
        const Tag vt = tag[end];
        key[end] = key[first];
        tag[end] = tag[first];
        is_adjust_heap<DESC, Key, Tag>(key, tag, first, 0, end - first, vk, vt);
    }
}

// std::__unguarded_partition_pivot(first, last): returns the cut
template <bool DESC, class Key, class Tag>
RLAP_HD inline int is_partition_pivot(Key* key, Tag* tag, int first, int last) {
    const int mid = first + (last - first) / 2;
    {   // std::__move_median_to_first(first, first + 1, mid, last - 1)
        const int a = first + 1, b = mid, c = last - 1;
        const Key ka = key[a], kb = key[b], kc = key[c];
        int m;
        if (is_before<DESC, Key>(ka, kb)) {
            if (is_before<DESC, Key>(kb, kc)) m = b;
            else if (is_before<DESC, Key>(ka, kc)) m = c;
            else m = a;
        } else if (is_before<DESC, Key>(ka, kc)) m = a;
        else if (is_before<DESC, Key>(kb, kc)) m = c;
        else m = b;
        is_swap(key, tag, first, m);
    }
    const Key pivot = key[first];   // the pivot stays at `first` during std::__unguarded_partition(first + 1, last, first)
    int lo = first + 1, hi = last;
    while (true) {
        while (is_before<DESC, Key>(key[lo], pivot)) ++lo;
        --hi;
        while (is_before<DESC, Key>(pivot, key[hi])) --hi;
        if (!(lo < hi)) return lo;
        is_swap(key, tag, lo, hi);
        ++lo;
    }
}

constexpr int IS_THRESHOLD = 16;      // libstdc++'s _S_threshold
constexpr int IS_MAX_DEPTH = 64;      // 2 * floor(log2 n) for any int n

RLAP_HD inline int is_lg(int n) {       // std::__lg
    int lg = 0;
    while ((n >> (lg + 1)) != 0) lg++;
    return lg;
}

// std::__introsort_loop(first, last, depth) on (key[i], tag[i]). The recursion on the right part is an explicit stack
// of the left parts still to do (at most one frame per level of the depth limit).
template <bool DESC, class Key, class Tag>
RLAP_HD inline void introsort_loop_range(Key* key, Tag* tag, int first, int last, int depth) {
    if (last - first <= IS_THRESHOLD) return;
    int stk_first[IS_MAX_DEPTH], stk_last[IS_MAX_DEPTH];
    signed char stk_depth[IS_MAX_DEPTH];
    int sp = 0;
    while (true) {
        if (last - first > IS_THRESHOLD) {
            if (depth == 0) {
#ifdef RLAP_IS_HEAP_HOOK
                RLAP_IS_HEAP_HOOK;   // host test: count how often the depth limit is reached
#endif
                is_heap_sort<DESC, Key, Tag>(key, tag, first, last);
            } else {
                --depth;
                const int cut = is_partition_pivot<DESC, Key, Tag>(key, tag, first, last);
                // __introsort_loop(cut, last, depth) now, (first, cut, depth) when it returns
                stk_first[sp] = first; stk_last[sp] = cut; stk_depth[sp] = (signed char)depth;
                sp++;
                first = cut;
                continue;
            }
        }
        if (sp == 0) break;
        --sp;
        first = stk_first[sp]; last = stk_last[sp]; depth = stk_depth[sp];
    }
}

// the whole loop of std::sort(first = 0, last = n): depth limit 2 * std::__lg(n)
template <bool DESC, class Key, class Tag>
RLAP_HD inline void introsort_loop_arrange(Key* key, Tag* tag, int n) {
    if (n <= IS_THRESHOLD) return;
    introsort_loop_range<DESC, Key, Tag>(key, tag, 0, n, 2 * is_lg(n));
}

// ---------------------------------------------------------------------------------------------
// the same loop run by a warp
// ---------------------------------------------------------------------------------------------
// What is sequential in the loop is one partition pass: the two pointers of std::__unguarded_partition meet in the
// middle, and every swap pairs the i-th element from the left that is not before the pivot with the i-th from the
// right that is not after it. The parts the recursion produces are independent of each other. A warp therefore
//   * runs the pass over a long range 32 + 32 elements at a time: the lanes test a chunk at either end, the stops of
//     the two chunks are paired in order and swapped at once, the chunk whose stops are used up moves on - the state
//     (lo, hi) between two steps is exactly the state of the sequential loop after the same swaps;
//   * hands every range of at most `small` elements (and a range whose depth limit is used up: heap sort) to ONE
//     lane, 32 ranges side by side. A long input yields dozens of such ranges and IS_SMALL = 64 keeps the lanes busy;
//     a short one (up to IS_SHORT elements) yields two or three, and taking all passes with the whole warp
//     (small = the 16 of libstdc++, below which the loop stops anyway) is quicker than one lane per range.
// The control flow is warp uniform: every lane keeps its own copy of the stack and of the list of small ranges.
// `WP` supplies the three warp primitives (ballot over a per-lane predicate, per-lane execution, barrier); the host
// twin in tests/native/introsort_check.cc runs the lanes one after the other, so this very code is what is pinned
// against std::sort.
#ifndef RLAP_IS_SMALL
#define RLAP_IS_SMALL 64
#endif
#ifndef RLAP_IS_SHORT
#define RLAP_IS_SHORT 128
#endif
constexpr int IS_SMALL = RLAP_IS_SMALL;      // see introsort_loop_arrange_warp
constexpr int IS_SHORT = RLAP_IS_SHORT;

RLAP_HD inline int is_popc(unsigned m) {
#if defined(__CUDA_ARCH__)
    return __popc(m);
#else
    return __builtin_popcount(m);
#endif
}
// position of the n-th (0-based) set bit of m; n < popcount(m). Five halving steps, no loop over the bits: a lone warp
// issues one dependent instruction every few cycles, and a bit-by-bit loop here was most of a partition step.
RLAP_HD inline int is_nth_bit(unsigned m, int n) {
    int pos = 0, c;
    c = is_popc(m & 0xffffu); if (n >= c) { n -= c; pos += 16; m >>= 16; }
    c = is_popc(m & 0xffu);   if (n >= c) { n -= c; pos += 8;  m >>= 8; }
    c = is_popc(m & 0xfu);    if (n >= c) { n -= c; pos += 4;  m >>= 4; }
    c = is_popc(m & 0x3u);    if (n >= c) { n -= c; pos += 2;  m >>= 2; }
    c = (int)(m & 1u);        if (n >= c) { pos += 1; }
    return pos;
}

// std::__unguarded_partition_pivot(first, last) by a warp; last - first > 16. Returns the cut (uniform).
template <bool DESC, class Key, class Tag, class WP>
RLAP_HD inline int is_partition_pivot_warp(WP& wp, Key* key, Tag* tag, int first, int last) {
    {   // std::__move_median_to_first(first, first + 1, mid, last - 1): every lane decides, lane 0 swaps
        const int a = first + 1, b = first + (last - first) / 2, c = last - 1;
        const Key ka = key[a], kb = key[b], kc = key[c];
        int m;
        if (is_before<DESC, Key>(ka, kb)) {
            if (is_before<DESC, Key>(kb, kc)) m = b;
            else if (is_before<DESC, Key>(ka, kc)) m = c;
            else m = a;
        } else if (is_before<DESC, Key>(ka, kc)) m = a;
        else if (is_before<DESC, Key>(kb, kc)) m = c;
        else m = b;
        wp.sync();                                       // every lane has read the three keys
        wp.each([&](int lane) { if (lane == 0) is_swap(key, tag, first, m); });
        wp.sync();
    }
    const Key pivot = key[first];
    // state of std::__unguarded_partition at the top of its loop: lo = next position the left pointer examines,
    // hi = one past the next position the right pointer examines; [lo, hi) is untouched
    int lo = first + 1, hi = last;
    while (hi - lo > 32) {
        const int cr = (hi - lo - 32 < 32) ? (hi - lo - 32) : 32;    // right chunk: positions hi - 1 down to hi - cr
        const unsigned ML = wp.ballot([&](int lane) { return !is_before<DESC, Key>(key[lo + lane], pivot); });
        const unsigned MR = wp.ballot([&](int lane) { return lane < cr && !is_before<DESC, Key>(pivot, key[hi - 1 - lane]); });
        const int nL = is_popc(ML), nR = is_popc(MR);
        const int t = nL < nR ? nL : nR;
        wp.each([&](int lane) {
            if (lane < t) is_swap(key, tag, lo + is_nth_bit(ML, lane), hi - 1 - is_nth_bit(MR, lane));
        });
        wp.sync();
        // t = min: at least one chunk has used up its stops and moves on; the other side stays behind its last
        // swapped stop (or where it was, with t = 0)
        if (nL == t) lo += 32; else if (t > 0) lo += is_nth_bit(ML, t - 1) + 1;
        if (nR == t) hi -= cr; else if (t > 0) hi -= is_nth_bit(MR, t - 1) + 1;
    }
    // the last chunk: [lo, hi) holds at most 32 elements, the pointers meet inside it
    const int len = hi - lo;
    const unsigned ML = wp.ballot([&](int lane) { return lane < len && !is_before<DESC, Key>(key[lo + lane], pivot); });
    const unsigned MR = wp.ballot([&](int lane) { return lane < len && !is_before<DESC, Key>(pivot, key[hi - 1 - lane]); });
    const int nL = is_popc(ML), nR = is_popc(MR);
    const int tmax = nL < nR ? nL : nR;
    // pairs are swapped while the left stop is below the right stop (l_i rises, r_i falls: a prefix)
    const unsigned OK = wp.ballot([&](int lane) {
        return lane < tmax && lo + is_nth_bit(ML, lane) < hi - 1 - is_nth_bit(MR, lane);
    });
    const int k = is_popc(OK);
    wp.each([&](int lane) {
        if (lane < k) is_swap(key, tag, lo + is_nth_bit(ML, lane), hi - 1 - is_nth_bit(MR, lane));
    });
    wp.sync();
    // where the left pointer stops next: the (k+1)-th left stop if it lies below the k-th right stop, else the k-th
    // right stop (it holds a swapped element that is not before the pivot), else - nothing swapped here and no stop in
    // the chunk - position hi, which holds such an element from an earlier step
    const int rk = (k > 0) ? hi - 1 - is_nth_bit(MR, k - 1) : hi;
    int cut = rk;
    if (nL > k) {
        const int lnext = lo + is_nth_bit(ML, k);
        if (lnext < rk) cut = lnext;
    }
    return cut;
}

// the listed short ranges, one per lane
template <bool DESC, class Key, class Tag, class WP>
RLAP_HD inline void is_run_small(WP& wp, Key* key, Tag* tag, const int* sm_first, const int* sm_last,
                                 const signed char* sm_depth, int ns) {
    wp.sync();
    wp.each([&](int lane) {
        if (lane < ns) introsort_loop_range<DESC, Key, Tag>(key, tag, sm_first[lane], sm_last[lane], (int)sm_depth[lane]);
    });
    wp.sync();
}

// std::__introsort_loop(0, n, 2 * std::__lg(n)) by a warp (all 32 lanes call, uniform arguments)
template <bool DESC, class Key, class Tag, class WP>
RLAP_HD inline void introsort_loop_arrange_warp(WP& wp, Key* key, Tag* tag, int n) {
    if (n <= IS_THRESHOLD) return;
    int stk_first[IS_MAX_DEPTH], stk_last[IS_MAX_DEPTH];
    signed char stk_depth[IS_MAX_DEPTH];
    int sm_first[32], sm_last[32];
    signed char sm_depth[32];
    int sp = 0, ns = 0;
    int first = 0, last = n, depth = 2 * is_lg(n);
    const int small = (n <= IS_SHORT) ? IS_THRESHOLD : IS_SMALL;
    while (true) {
        if (last - first > IS_THRESHOLD) {
            if (last - first <= small || depth == 0) {
                sm_first[ns] = first; sm_last[ns] = last; sm_depth[ns] = (signed char)depth;
                if (++ns == 32) { is_run_small<DESC, Key, Tag, WP>(wp, key, tag, sm_first, sm_last, sm_depth, ns); ns = 0; }
            } else {
                --depth;
                const int cut = is_partition_pivot_warp<DESC, Key, Tag, WP>(wp, key, tag, first, last);
                stk_first[sp] = first; stk_last[sp] = cut; stk_depth[sp] = (signed char)depth;
                sp++;
                first = cut;
                continue;
            }
        }
        if (sp == 0) break;
        --sp;
        first = stk_first[sp]; last = stk_last[sp]; depth = stk_depth[sp];
    }
    if (ns > 0) is_run_small<DESC, Key, Tag, WP>(wp, key, tag, sm_first, sm_last, sm_depth, ns);
}

}  // namespace rlap
