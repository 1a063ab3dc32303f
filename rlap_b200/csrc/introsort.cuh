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
template <bool DESC>
RLAP_HD inline bool is_before(uint64_t x, uint64_t y) { return DESC ? (x > y) : (x < y); }

template <class Tag>
RLAP_HD inline void is_swap(uint64_t* key, Tag* tag, int i, int j) {
    const uint64_t k = key[i]; key[i] = key[j]; key[j] = k;
    const Tag t = tag[i]; tag[i] = tag[j]; tag[j] = t;
}

// std::__adjust_heap followed by std::__push_heap on [first, first + len)
template <bool DESC, class Tag>
RLAP_HD inline void is_adjust_heap(uint64_t* key, Tag* tag, int first, int hole, int len, uint64_t vk, Tag vt) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (is_before<DESC>(key[first + child], key[first + child - 1])) child--;
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
    while (hole > top && is_before<DESC>(key[first + parent], vk)) {
        key[first + hole] = key[first + parent];
        tag[first + hole] = tag[first + parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    key[first + hole] = vk;
    tag[first + hole] = vt;
}

// std::__partial_sort(first, last, last): __make_heap, then __sort_heap (the __heap_select loop is empty)
template <bool DESC, class Tag>
RLAP_HD inline void is_heap_sort(uint64_t* key, Tag* tag, int first, int last) {
    const int len = last - first;
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            is_adjust_heap<DESC, Tag>(key, tag, first, parent, len, key[first + parent], tag[first + parent]);
            if (parent == 0) break;
            parent--;
        }
    }
    int end = last;
    while (end - first > 1) {
        --end;
        const uint64_t vk = key[end];
        const Tag vt = tag[end];
        key[end] = key[first];
        tag[end] = tag[first];
        is_adjust_heap<DESC, Tag>(key, tag, first, 0, end - first, vk, vt);
    }
}

// std::__unguarded_partition_pivot(first, last): returns the cut
template <bool DESC, class Tag>
RLAP_HD inline int is_partition_pivot(uint64_t* key, Tag* tag, int first, int last) {
    const int mid = first + (last - first) / 2;
    {   // std::__move_median_to_first(first, first + 1, mid, last - 1)
        const int a = first + 1, b = mid, c = last - 1;
        const uint64_t ka = key[a], kb = key[b], kc = key[c];
        int m;
        if (is_before<DESC>(ka, kb)) {
            if (is_before<DESC>(kb, kc)) m = b;
            else if (is_before<DESC>(ka, kc)) m = c;
            else m = a;
        } else if (is_before<DESC>(ka, kc)) m = a;
        else if (is_before<DESC>(kb, kc)) m = c;
        else m = b;
        is_swap(key, tag, first, m);
    }
    const uint64_t pivot = key[first];   // the pivot stays at `first` during std::__unguarded_partition(first + 1, last, first)
    int lo = first + 1, hi = last;
    while (true) {
        while (is_before<DESC>(key[lo], pivot)) ++lo;
        --hi;
        while (is_before<DESC>(pivot, key[hi])) --hi;
        if (!(lo < hi)) return lo;
        is_swap(key, tag, lo, hi);
        ++lo;
    }
}

constexpr int IS_THRESHOLD = 16;      // libstdc++'s _S_threshold
constexpr int IS_MAX_DEPTH = 64;      // 2 * floor(log2 n) for any int n

// std::__introsort_loop(first = 0, last = n, 2 * std::__lg(n)) on (key[i], tag[i]). The recursion on the right
// part is an explicit stack of the left parts still to do (at most one frame per level of the depth limit).
template <bool DESC, class Tag>
RLAP_HD inline void introsort_loop_arrange(uint64_t* key, Tag* tag, int n) {
    if (n <= IS_THRESHOLD) return;
    int lg = 0;
    while ((n >> (lg + 1)) != 0) lg++;
    int stk_first[IS_MAX_DEPTH], stk_last[IS_MAX_DEPTH];
    signed char stk_depth[IS_MAX_DEPTH];
    int sp = 0;
    int first = 0, last = n, depth = 2 * lg;
    while (true) {
        if (last - first > IS_THRESHOLD) {
            if (depth == 0) {
#ifdef RLAP_IS_HEAP_HOOK
                RLAP_IS_HEAP_HOOK;   // host test: count how often the depth limit is reached
#endif
                is_heap_sort<DESC, Tag>(key, tag, first, last);
            } else {
                --depth;
                const int cut = is_partition_pivot<DESC, Tag>(key, tag, first, last);
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

}  // namespace rlap
