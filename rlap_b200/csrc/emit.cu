// emit.cu — emission (A.5): surviving vertices, multi-edges merged, rows sorted by (col, row).
// Replaces the output assembly of the reference (rlap/csrc/preconditioner.cc:435-457 / 789-810 /
// 916-934: pop every remaining vertex, walk its list, sort, merge, push rows).
//
// No list walking here: the exact live count of every surviving vertex is known (the `live`
// counters the elimination maintains), so every survivor owns a staging segment of `live` slots:
//   k_emit_prep     live counts of survivors            -> scan -> rawoff
//   k_emit_scatter  streams the fill pool: entries with two alive endpoints go to the TAIL of their owner's segment
//   k_emit_fsort    per survivor with at most 32 fill entries: sorts them in place and looks for a multi-edge (two
//                   fills to the same neighbour, or a fill parallel to a base edge). Without one - 99 % of the
//                   survivors - the row count is the live count and nothing else is staged: "clean".
//   the rest (a multi-edge, or more than 32 fills: the hubs) take the merge path: k_emit_base stages their alive
//   base entries at the front of the segment, k_emit_sort_* sort the segment by neighbour and merge multi-edges
//   with the fixed-point rule, in place                  -> scan -> outoff
//   k_emit_write    clean survivors: the alive base entries (streamed from the CSR, ascending) and the sorted fills
//                   are merged by rank straight into the caller's buffers; the others are copied from staging
#include <stdlib.h>
#include <mutex>
#include "rlap_device.cuh"
#include "schur.cuh"
#include "scan.cuh"
#include "star.cuh"

namespace rlap {

// rawcnt / cursor reuse two per-vertex arrays that are dead once the elimination kernel has returned
__device__ __forceinline__ int* rawcnt_of(const SchurParams& P) { return P.blk; }
__device__ __forceinline__ int* cursor_of(const SchurParams& P) { return P.candround; }

// a failed elimination (pool overflow, scratch overflow) leaves reserved-but-unwritten pool slots and inconsistent
// counters behind: the emission kernels do nothing then, the caller only reads the status
__device__ __forceinline__ bool run_failed(const SchurParams& P) { return ldcg_i32(P.ctr + CTR_STATUS) != 0; }

__global__ void k_emit_prep(SchurParams P) {
    const long long VN = (long long)P.V * P.n;
    long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= VN) return;
    if (run_failed(P)) { rawcnt_of(P)[idx] = 0; cursor_of(P)[idx] = 0; P.outcnt[idx] = 0; return; }
    int c = (P.state[idx] != 2) ? *live_p(P, idx) : 0;
    rawcnt_of(P)[idx] = c;
    cursor_of(P)[idx] = 0;
}

// Base entries of the survivors on the merge path (cursor == MERGE_PATH): an 8-lane tile per vertex walks its
// (neighbour-ascending) CSR segment and writes the entries whose neighbour is alive to the front of the vertex's
// staging segment, in order and without atomics; rows longer than 64 are then served by the whole warp.
constexpr int MERGE_PATH = -1;
__global__ void __launch_bounds__(256) k_emit_base(SchurParams P) {
    if (run_failed(P)) return;
    const long long VN = (long long)P.V * P.n;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31, tl = lane & 7, tile = lane >> 3;
    const unsigned tlt = (1u << tl) - 1u;
    for (long long base = gw * 4; base < VN; base += nw * 4) {
        const long long idx = base + tile;
        int b = 0, nb = 0;
        long long off = 0;
        size_t vb = 0;
        bool big = false;
        if (idx < VN && cursor_of(P)[idx] == MERGE_PATH) {
            const int v = (int)(idx % P.n);
            vb = (size_t)(idx - v);
            b = __ldg(P.ptr + v);
            nb = __ldg(P.ptr + v + 1) - b;
            off = P.rawoff[idx];
            big = nb > 64;
        }
        int cnt = 0;
        const int nbmax = __reduce_max_sync(RLAP_FULL_MASK, big ? 0 : nb);   // the tiles of a warp loop in lock step
        for (int p0 = 0; p0 < nbmax; p0 += 8) {
            const int p = p0 + tl;
            int u = 0;
            bool ok = !big && p < nb;
            if (ok) { u = __ldg(P.col + b + p); ok = P.state[vb + u] != 2; }
            const unsigned m = (__ballot_sync(RLAP_FULL_MASK, ok) >> (tile * 8)) & 0xffu;
            if (ok) P.raw[off + cnt + __popc(m & tlt)] = pack_a((uint32_t)u, __ldg(P.w + b + p));
            cnt += __popc(m);
        }
        __syncwarp();
        unsigned todo = __ballot_sync(RLAP_FULL_MASK, big && tl == 0);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int kb = __shfl_sync(RLAP_FULL_MASK, b, k), knb = __shfl_sync(RLAP_FULL_MASK, nb, k);
            const long long koff = __shfl_sync(RLAP_FULL_MASK, off, k);
            const size_t kvb = (size_t)__shfl_sync(RLAP_FULL_MASK, (unsigned long long)vb, k);
            int c = 0;
            for (int p0 = 0; p0 < knb; p0 += 32) {
                const int p = p0 + lane;
                int u = 0;
                bool ok = p < knb;
                if (ok) { u = __ldg(P.col + kb + p); ok = P.state[kvb + u] != 2; }
                const unsigned m = __ballot_sync(RLAP_FULL_MASK, ok);
                if (ok) P.raw[koff + c + __popc(m & ((1u << lane) - 1u))] = pack_a((uint32_t)u, __ldg(P.w + kb + p));
                c += __popc(m);
            }
        }
    }
}

// Fill entries: grid-stride over the view's pool (blockIdx.y = view), one thread per fill edge = the two entries
// (j <- k), (k <- j) it left in adjacent slots (reservations are even, so pairs never straddle); if both endpoints
// are alive each entry goes to the tail of its owner's staging segment, filled from the end backwards (the front is
// where the alive base entries of a merge-path vertex go; cursor[v] counts the fills of v)
__global__ void __launch_bounds__(256) k_emit_scatter(SchurParams P) {
    if (run_failed(P)) return;
    const int view = blockIdx.y;
    const size_t vb = (size_t)view * (size_t)P.n;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, nt = (long long)gridDim.x * blockDim.x;
    long long used = (long long)P.pool_cursor[view];
    if (used > P.pool_cap) used = P.pool_cap;
    const int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
    for (long long e = 2 * t0; e + 1 < used; e += 2 * nt) {
        const int4 e0 = __ldcs(pool + e);        // {k, w, next, j}
        if (e0.w < 0) continue;                  // tombstones come in pairs as well
        const int j = e0.w, k = e0.x;
        if (P.state[vb + j] == 2 || P.state[vb + k] == 2) continue;
        const long long p0 = P.rawoff[vb + j + 1] - 1 - atomicAdd(cursor_of(P) + vb + j, 1);
        const long long p1 = P.rawoff[vb + k + 1] - 1 - atomicAdd(cursor_of(P) + vb + k, 1);
        if (p0 >= 0 && p1 >= 0 && p0 < P.raw_cap && p1 < P.raw_cap) {
            P.raw[p0] = ((uint64_t)(uint32_t)k << 32) | (uint64_t)(uint32_t)e0.y;
            P.raw[p1] = ((uint64_t)(uint32_t)j << 32) | (uint64_t)(uint32_t)e0.y;
        } else {
            set_status(P, 6);
        }
    }
}

// sort + merge a staged star (shared memory or scratch) and write the merged entries back to the
// start of the vertex's raw segment, neighbours ascending
template <bool CTA>
__device__ void emit_sort_staged(const SchurParams& P, size_t idx, StarBuf sb, CtaScratch* cs) {
    const int gs = g_size<CTA>(), r = g_rank<CTA>(), lane = threadIdx.x & 31;
    const int lraw = rawcnt_of(P)[idx];
    const long long off = P.rawoff[idx];
    if (lraw > sb.cap) {
        if (r == 0) { set_status(P, 6); P.outcnt[idx] = 0; }
        g_sync<CTA>();
        return;
    }
    uint32_t wmaxb = 0;
    const int P2 = next_pow2(lraw);
    for (int i = r; i < P2; i += gs) {
        uint64_t a = (i < lraw) ? P.raw[off + i] : RLAP_PAD_A;
        sb.A[i] = a;
        if (i < lraw) wmaxb = max(wmaxb, (uint32_t)a);
    }
    g_sync<CTA>();
    g_bitonic_sort_keys<CTA>(sb.A, P2);
    // multi-edges are rare among survivors: look for one before paying for the fixed-point merge
    int dup = 0;
    for (int i = r + 1; i < lraw; i += gs) dup |= (a_nbr(sb.A[i]) == a_nbr(sb.A[i - 1]));
    dup = CTA ? __syncthreads_or(dup) : __any_sync(RLAP_FULL_MASK, dup);
    int L = lraw;
    if (dup) {
        wmaxb = g_max_u32<CTA>(wmaxb, cs);
        g_sync<CTA>();
        const int shift = star_shift(__uint_as_float(wmaxb), lraw);
        L = star_merge_sorted<CTA>(sb, lraw, shift, cs);
    }
    int carry = 0;
    for (int base = 0; base < lraw; base += gs) {
        int i = base + r;
        uint64_t a = (i < lraw) ? sb.A[i] : RLAP_PAD_A;
        bool keep = (i < lraw) && !a_dead(a);
        unsigned m = __ballot_sync(RLAP_FULL_MASK, keep);
        int pos = __popc(m & ((1u << lane) - 1u));
        if (CTA) {
            int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
            __syncthreads();
            if (lane == 0) cs->wsum[w] = (unsigned long long)__popc(m);
            __syncthreads();
            int add = 0, tot = 0;
            for (int k = 0; k < nw; k++) { int c = (int)cs->wsum[k]; if (k < w) add += c; tot += c; }
            pos += add + carry;
            carry += tot;
        } else {
            pos += carry;
            carry += __popc(m);
        }
        if (keep) P.raw[off + pos] = a;
    }
    if (r == 0) P.outcnt[idx] = L;
    g_sync<CTA>();
}

// A warp sorts one segment of up to 32 * R entries in registers (element e = r * 32 + lane), merges the runs of
// equal neighbour with the fixed-point rule (segmented suffix sums over the register tile) and writes the
// merged entries back to the start of the segment, neighbours ascending.
template <int R>
__device__ void emit_sort_regs(const SchurParams& P, size_t idx, int lv, long long off) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    uint64_t a[R];
    uint32_t wmaxb = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = r * 32 + lane;
        a[r] = e < lv ? P.raw[off + e] : RLAP_PAD_A;
        if (e < lv) wmaxb = max(wmaxb, (uint32_t)a[r]);
    }
    warp_sort_regs<R>(a);
    // heads of runs
    unsigned headm[R];
    bool anydup = false;
    {
        uint32_t carry = 0xffffffffu;   // neighbour id of element e - 1 for lane 0 (no vertex has this id)
#pragma unroll
        for (int r = 0; r < R; r++) {
            const uint32_t nb = a_nbr(a[r]);
            uint32_t prev = __shfl_up_sync(RLAP_FULL_MASK, nb, 1);
            if (lane == 0) prev = carry;
            const bool valid = a[r] != RLAP_PAD_A;
            const bool head = valid && prev != nb;
            headm[r] = __ballot_sync(RLAP_FULL_MASK, head);
            anydup |= (headm[r] != __ballot_sync(RLAP_FULL_MASK, valid));
            carry = __shfl_sync(RLAP_FULL_MASK, nb, 31);
        }
    }
    if (anydup) {
        wmaxb = warp_max_u32(wmaxb);
        const int shift = star_shift(__uint_as_float(wmaxb), lv);
        unsigned long long qs[R];
#pragma unroll
        for (int r = 0; r < R; r++) qs[r] = (a[r] != RLAP_PAD_A) ? quantize(a_w(a[r]), shift) : 0ull;
        // segmented suffix sums: after the step with distance d, qs(e) covers the run members in [e, e + 2d).
        // Rows are updated in ascending order, in place: a step only reads rows >= r that are still old.
#pragma unroll
        for (int d = 1; d < 32 * R; d <<= 1) {
#pragma unroll
            for (int r = 0; r < R; r++) {
                unsigned long long oq;
                uint32_t onb;
                if (d < 32) {
                    const int src = (lane + d) & 31;
                    unsigned long long q0 = __shfl_sync(RLAP_FULL_MASK, qs[r], src);
                    uint32_t n0 = __shfl_sync(RLAP_FULL_MASK, a_nbr(a[r]), src);
                    unsigned long long q1 = 0;
                    uint32_t n1 = 0xffffffffu;
                    if (r + 1 < R) {
                        q1 = __shfl_sync(RLAP_FULL_MASK, qs[r + 1], src);
                        n1 = __shfl_sync(RLAP_FULL_MASK, a_nbr(a[r + 1]), src);
                    }
                    const bool wrap = lane + d >= 32;
                    oq = wrap ? q1 : q0;
                    onb = wrap ? n1 : n0;
                } else {
                    const int r2 = r + (d >> 5);
                    oq = (r2 < R) ? qs[r2 < R ? r2 : 0] : 0ull;
                    onb = (r2 < R) ? a_nbr(a[r2 < R ? r2 : 0]) : 0xffffffffu;
                }
                if (a[r] != RLAP_PAD_A && onb == a_nbr(a[r])) qs[r] += oq;
            }
        }
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (a[r] == RLAP_PAD_A) continue;
            const bool head = (headm[r] >> lane) & 1u;
            if (!head) {
                a[r] = ((uint64_t)a_nbr(a[r]) << 32) | (uint64_t)RLAP_DEAD_W;
            } else {
                // merged iff the next element (e + 1) exists and is not the head of a run
                const int e1 = r * 32 + lane + 1;
                bool merged = false;
                if (e1 < lv) {
                    const unsigned hm = (lane == 31) ? ((r + 1 < R) ? headm[r + 1 < R ? r + 1 : 0] : 1u) : (headm[r] >> (lane + 1));
                    merged = (hm & 1u) == 0;
                }
                if (merged) a[r] = ((uint64_t)a_nbr(a[r]) << 32) | (uint64_t)__float_as_uint(dequantize_merged(qs[r], shift));
            }
        }
    }
    int rowbase = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const bool keep = a[r] != RLAP_PAD_A && !a_dead(a[r]);
        if (keep) P.raw[off + rowbase + __popc(headm[r] & lt)] = a[r];
        rowbase += __popc(headm[r]);
    }
    if (lane == 0) P.outcnt[idx] = rowbase;
}

// Segments of more than 32 entries are listed by size class in dense lists (the low-list buffers of the elimination
// are free by now): a hub-heavy stretch of vertex ids (the old vertices of a preferential-attachment graph) would
// otherwise hand all its big segments to the few warps that own that stretch.
constexpr int CAP_REGS = 512;   // largest segment sorted in registers
constexpr int N_CLASS = 7;      // 0..3: 64 / 128 / 256 / 512 entries (register sort), 4: <= CAP_CTA, 5: above, 6: <= 32
__device__ __forceinline__ unsigned int* class_list(const SchurParams& P, int c) {
    const size_t VN = (size_t)P.V * (size_t)P.n;
    if (c == 0) return P.wl;
    if (c == 1) return P.wl + VN;
    if (c == 5) return P.dl;
    if (c == 6) return P.low + (size_t)3 * (VN + 16);
    return P.low + (size_t)(c - 2) * (VN + 16);
}
__device__ __forceinline__ int size_class(int lv) {
    return lv <= 32 ? 6 : lv <= 64 ? 0 : lv <= 128 ? 1 : lv <= 256 ? 2 : lv <= CAP_REGS ? 3 : lv <= CAP_CTA ? 4 : 5;
}
__device__ __forceinline__ int* class_tail(const SchurParams& P, int c) { return P.ctr + CTR_EMIT_C0 + c; }

constexpr int CAP_BIG = 12288;   // largest segment sorted in the shared memory of one SM (k_emit_sort_big)

constexpr int FSORT_MAX = 32;    // fills of a clean survivor
constexpr int FS_LANE = 16;      // fills a single lane sorts in its shared-memory slot
constexpr int FS_NB_MAX = 40;    // longest base row a single lane walks

// Every survivor: sort its fill entries (the tail of its staging segment) by neighbour and look for a multi-edge.
// A warp owns 32 consecutive vertices and reads their counts in one go. A vertex with at most FS_LANE fills is
// handled by its lane alone (32 independent chains of loads per warp): insertion sort in a shared-memory slot, then a
// merge walk over the neighbour ids of its base row for a fill parallel to a base edge. 17..32 fills: the warp sorts
// them in registers, one vertex after the other. No multi-edge: the sorted fills go back in place and the survivor is
// "clean" (its row count is its live count; cursor keeps the fill count). Otherwise, or with more than 32 fills, the
// survivor takes the merge path (cursor = MERGE_PATH) and is listed by size class for the sort kernels.
__global__ void __launch_bounds__(256) k_emit_fsort(SchurParams P) {
    __shared__ uint64_t s_slot[8][FS_LANE * 32];
    const long long VN = (long long)P.V * P.n;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const bool failed = run_failed(P);
    uint64_t* sl = s_slot[threadIdx.x >> 5] + lane;    // element e of this lane: sl[e * 32]
    for (long long base = gw * 32; base < VN; base += nw * 32) {
        const long long idx = base + lane;
        int lv = 0, f = 0, b = 0, nb = 0;
        long long end = 0;
        if (idx < VN && !failed) {
            lv = rawcnt_of(P)[idx];
            if (lv > 0) {
                f = cursor_of(P)[idx];
                end = P.rawoff[idx + 1];
                const int v = (int)(idx % P.n);
                b = __ldg(P.ptr + v);
                nb = __ldg(P.ptr + v + 1) - b;
            }
        }
        if (idx < VN && (lv == 0 || f == 0)) P.outcnt[idx] = lv;      // eliminated / isolated, or no fills at all: clean
        int cls = (lv > 0 && f > FSORT_MAX) ? size_class(lv) : -1;
        const bool lanepath = lv > 0 && f > 0 && f <= FS_LANE && nb <= FS_NB_MAX;
        if (lanepath) {
            uint64_t* fp = P.raw + (end - f);
#pragma unroll 4
            for (int i = 0; i < f; i++) sl[i * 32] = fp[i];
            bool dup = false;
            for (int i = 1; i < f; i++) {
                const uint64_t x = sl[i * 32];
                int j = i - 1;
                while (j >= 0) {
                    const uint64_t y = sl[j * 32];
                    if (a_nbr(y) <= a_nbr(x)) { dup |= a_nbr(y) == a_nbr(x); break; }
                    sl[(j + 1) * 32] = y;
                    j--;
                }
                sl[(j + 1) * 32] = x;
            }
            if (!dup) {   // a fill parallel to a base edge of the vertex: both lists ascend, one merge walk
                int i = 0;
                uint32_t fn = a_nbr(sl[0]);
                for (int p0 = 0; p0 < nb && i < f; p0 += 8) {
                    uint32_t c8[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) c8[k] = (p0 + k < nb) ? (uint32_t)__ldg(P.col + b + p0 + k) : 0xffffffffu;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        while (i < f && fn < c8[k]) { i++; fn = i < f ? a_nbr(sl[i * 32]) : 0xffffffffu; }
                        if (i < f && fn == c8[k]) dup = true;
                    }
                }
            }
            if (!dup) {
                for (int i = 0; i < f; i++) fp[i] = sl[i * 32];
                P.outcnt[idx] = lv;
            } else {
                cls = size_class(lv);
            }
        }
        unsigned cand = __ballot_sync(RLAP_FULL_MASK, lv > 0 && f > 0 && f <= FSORT_MAX && !lanepath);
        while (cand) {
            const int k = __ffs(cand) - 1;
            cand &= cand - 1;
            const int kf = __shfl_sync(RLAP_FULL_MASK, f, k), kb = __shfl_sync(RLAP_FULL_MASK, b, k);
            const int knb = __shfl_sync(RLAP_FULL_MASK, nb, k);
            const long long kend = __shfl_sync(RLAP_FULL_MASK, end, k);
            uint64_t a = lane < kf ? P.raw[kend - kf + lane] : RLAP_PAD_A;
            a = warp_sort_u64(a);
            const bool valid = a != RLAP_PAD_A;
            const uint32_t nbr = a_nbr(a);
            const uint32_t prev = __shfl_up_sync(RLAP_FULL_MASK, nbr, 1);
            bool dup = valid && lane > 0 && prev == nbr;               // two fills to the same neighbour
            if (valid) {                                               // a fill parallel to a base edge of the vertex
                int lo = 0, hi = knb;
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((uint32_t)__ldg(P.col + kb + mid) < nbr) lo = mid + 1; else hi = mid;
                }
                dup |= lo < knb && (uint32_t)__ldg(P.col + kb + lo) == nbr;
            }
            if (!__any_sync(RLAP_FULL_MASK, dup)) {
                if (valid) P.raw[kend - kf + lane] = a;
                if (lane == k) P.outcnt[idx] = lv;
            } else if (lane == k) {
                cls = size_class(lv);
            }
        }
        // merge path: dense per-class lists, one tail bump per class and warp pass
        if (cls >= 0) cursor_of(P)[idx] = MERGE_PATH;
        if (__any_sync(RLAP_FULL_MASK, cls >= 0)) {
#pragma unroll
            for (int c = 0; c < N_CLASS; c++) {
                const unsigned m = __ballot_sync(RLAP_FULL_MASK, cls == c);
                if (m == 0) continue;
                int pos0 = 0;
                if (lane == __ffs(m) - 1) pos0 = atomicAdd(class_tail(P, c), __popc(m));
                pos0 = __shfl_sync(RLAP_FULL_MASK, pos0, __ffs(m) - 1);
                if (cls == c) class_list(P, c)[pos0 + __popc(m & lt)] = (unsigned int)idx;
            }
        }
    }
}

// merge path, segments of at most 32 entries: one warp each, one entry per lane
__global__ void __launch_bounds__(256) k_emit_sort_small(SchurParams P) {
    const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nw = (int)(((long long)gridDim.x * blockDim.x) >> 5);
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int end = *class_tail(P, 6);
    const unsigned int* list = class_list(P, 6);
    for (int it = gw; it < end; it += nw) {
        const unsigned int idx = list[it];
        const int lv = rawcnt_of(P)[idx];
        const long long off = P.rawoff[idx];
        uint64_t a = lane < lv ? P.raw[off + lane] : RLAP_PAD_A;
        a = warp_sort_u64(a);
        unsigned long long q;
        int shift;
        const unsigned hmask = warp_merge_sorted(a, q, shift, false);
        if ((hmask >> lane) & 1u) P.raw[off + __popc(hmask & lt)] = a;
        if (lane == 0) P.outcnt[idx] = __popc(hmask);
    }
}

// segments of 33..512 entries: one warp each, registers only. One kernel per size class (64, 128, 256, 512 entries =
// 2, 4, 8, 16 keys per lane) so that the small classes, which hold most segments, run at their own register budget.
// The list entry carries the class in its two top bits (V * n < 2^30); a warp takes 32 entries, reads their
// counts and offsets in one go and sorts the ones of its class.
template <int R>
__global__ void __launch_bounds__(256, (R <= 4 ? 4 : 2)) k_emit_sort_mid(SchurParams P) {
    constexpr int CLS = (R == 2 ? 0 : R == 4 ? 1 : R == 8 ? 2 : 3);
    const int gw = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nw = (int)(((long long)gridDim.x * blockDim.x) >> 5);
    const int end = *class_tail(P, CLS);
    const unsigned int* list = class_list(P, CLS);
    for (int it = gw; it < end; it += nw) {
        const unsigned int idx = list[it];
        emit_sort_regs<R>(P, idx, rawcnt_of(P)[idx], P.rawoff[idx]);
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS, 2) k_emit_sort_block(SchurParams P) {
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ CtaScratch cs;
    const int end = *class_tail(P, 4);
    const unsigned int* list = class_list(P, 4);
    for (int it = (int)blockIdx.x; it < end; it += (int)gridDim.x) {
        emit_sort_staged<true>(P, list[it], cta_buf(smem), &cs);
        __syncthreads();
    }
    // segments beyond the shared memory of an SM: the NSLOT blocks that own a global scratch slot
    if ((int)blockIdx.x < NSLOT) {
        const int end5 = *class_tail(P, 5);
        const unsigned int* list5 = class_list(P, 5);
        for (int it = (int)blockIdx.x; it < end5; it += NSLOT) {
            const unsigned int idx = list5[it];
            if (rawcnt_of(P)[idx] <= CAP_BIG) continue;
            emit_sort_staged<true>(P, idx, scratch_buf(P), &cs);
            __syncthreads();
        }
    }
}

// segments of CAP_CTA + 1 .. CAP_BIG entries (the hubs that survive): one block per SM with most of its shared
// memory as the staging area (keys + fixed-point weights), instead of the global scratch slots
__global__ void __launch_bounds__(BLOCK_THREADS, 1) k_emit_sort_big(SchurParams P) {
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ CtaScratch cs;
    StarBuf sb;
    sb.A = smem; sb.Q = smem + CAP_BIG; sb.K = sb.Q; sb.cap = CAP_BIG;
    const int end = *class_tail(P, 5);
    const unsigned int* list = class_list(P, 5);
    for (int it = (int)blockIdx.x; it < end; it += (int)gridDim.x) {
        const unsigned int idx = list[it];
        if (rawcnt_of(P)[idx] > CAP_BIG) continue;
        emit_sort_staged<true>(P, idx, sb, &cs);
        __syncthreads();
    }
}

// Final rows. A warp owns 32 consecutive vertices; their rows are contiguous in the output.
//  * clean survivor (no multi-edge) of at most WR_LMAX rows: its LANE merges the alive base entries, streamed from
//    the CSR in neighbour order, with its sorted fills into the warp's shared-memory staging area (32 independent
//    chains of loads per warp); the warp then writes the staged rows out together, coalesced.
//  * longer clean survivors: the warp merges by rank (an alive base entry goes to (alive base entries before it) +
//    (fills with a smaller neighbour), fill i to i + (alive base entries with a smaller neighbour), both from
//    shuffle binary searches).
//  * merge-path survivors: copy of the merged staging segment.
constexpr int WR_LMAX = 64;      // longest row a single lane merges
constexpr int WR_NBMAX = 40;     // longest base row a single lane walks
constexpr int WR_CAP = 1024;     // staged rows per pass (a pass may run over by one vertex: + WR_LMAX)
constexpr int WR_WARPS = 4;
__global__ void __launch_bounds__(WR_WARPS * 32) k_emit_write(SchurParams P, int* out_row, int* out_col, float* out_w,
                                                              double* out_f64) {
    __shared__ uint64_t s_stage[WR_WARPS][WR_CAP + WR_LMAX];
    if (run_failed(P)) return;
    const long long VN = (long long)P.V * P.n;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const bool check = (P.flags & 8) != 0;
    uint64_t* stage = s_stage[threadIdx.x >> 5];
    auto put = [&](long long w, uint32_t r, int c, float wt) {
        if (out_row) {
            __stcs(out_row + w, (int)r);
            if (out_col) __stcs(out_col + w, c);      // NULL: the caller rebuilds the columns from the column pointers
            __stcs(out_w + w, wt);
        }
        if (out_f64) {
            out_f64[w * 3 + 0] = (double)r;
            out_f64[w * 3 + 1] = (double)c;
            out_f64[w * 3 + 2] = (double)wt;
        }
    };
    auto mismatch = [&](long long vidx, int got, int want) {   // RLAP_FLAG_CHECK_LIVE: alive base entries + fills must be the live count
        atomicAdd(P.stats + 7, 1ull);
        atomicMax(P.stats + 6, ((unsigned long long)(unsigned)vidx << 32) | ((unsigned long long)(unsigned)(got & 0xffff) << 16) | (unsigned)(want & 0xffff));
    };
    for (long long base = gw * 32; base < VN; base += nw * 32) {
        const long long idx = base + lane;
        int L = 0, f = 0, b = 0, nb = 0, v = 0;
        long long dst = 0, src = 0, end = 0;
        if (idx < VN) {
            L = P.outcnt[idx];
            if (L > 0) {
                f = cursor_of(P)[idx];
                dst = P.outoff[idx];
                src = P.rawoff[idx];
                end = P.rawoff[idx + 1];
                v = (int)(idx % P.n);
                b = __ldg(P.ptr + v);
                nb = __ldg(P.ptr + v + 1) - b;
            }
        }
        // ---- lane tier
        const bool staged = L > 0 && f != MERGE_PATH && L <= WR_LMAX && nb <= WR_NBMAX;
        const int Ls = staged ? L : 0;
        int srel = Ls;   // exclusive prefix of the staged row counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(RLAP_FULL_MASK, srel, d);
            if (lane >= d) srel += t;
        }
        const int stot = __shfl_sync(RLAP_FULL_MASK, srel, 31);
        srel -= Ls;
        int s0 = 0;
        for (int pass = 0; pass * WR_CAP < stot; pass++) {
            const bool mine = staged && srel / WR_CAP == pass;
            if (mine) {
                uint64_t* st = stage + (srel - pass * WR_CAP);
                const size_t vb = (size_t)(idx - v);
                const uint64_t* fp = P.raw + (end - f);   // sorted fills
                // the fills are parked at the end of this lane's staging span and merged forward in place: the output
                // position (base entries emitted + fills emitted) never passes the next unread fill
                uint64_t* sf = st + (L - f);
#pragma unroll 4
                for (int i = 0; i < f; i++) sf[i] = fp[i];
                int r = 0, i = 0;
                uint32_t nfn = f > 0 ? a_nbr(sf[0]) : 0xffffffffu;   // no fill left: the largest id, never "smaller"
                for (int p0 = 0; p0 < nb; p0 += 8) {
                    uint32_t c8[8];
                    float w8[8];
                    uint8_t s8[8];
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const bool in = p0 + k < nb;
                        c8[k] = in ? (uint32_t)__ldg(P.col + b + p0 + k) : 0xffffffffu;
                        w8[k] = in ? __ldg(P.w + b + p0 + k) : 0.f;
                    }
#pragma unroll
                    for (int k = 0; k < 8; k++) s8[k] = (c8[k] != 0xffffffffu) ? P.state[vb + c8[k]] : (uint8_t)2;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        if (s8[k] != 2) {
                            while (nfn < c8[k]) {
                                if (r < L) st[r] = sf[i];
                                r++; i++;
                                nfn = i < f ? a_nbr(sf[i]) : 0xffffffffu;
                            }
                            if (r < L) st[r] = pack_a(c8[k], w8[k]);
                            r++;
                        }
                    }
                }
                r += f - i;   // the remaining fills are already in place
                if (check && r != L) mismatch(idx, r, L);
            }
            __syncwarp();
            // staged rows [s0, s1) of this pass -> output, coalesced; the owner of staged row s is the last lane whose
            // prefix is <= s (lanes that stage nothing share the prefix of their successor)
            const int s1 = __reduce_max_sync(RLAP_FULL_MASK, mine ? srel + Ls : s0);
            for (int sb = s0; sb < s1; sb += 32) {
                const int sidx = sb + lane;
                int lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int r = __shfl_sync(RLAP_FULL_MASK, srel, (lo + step) & 31);
                    if (lo + step < 32 && r <= sidx) lo += step;
                }
                const long long odst = __shfl_sync(RLAP_FULL_MASK, dst, lo);
                const int osrel = __shfl_sync(RLAP_FULL_MASK, srel, lo);
                const int ov = __shfl_sync(RLAP_FULL_MASK, v, lo);
                if (sidx < s1) {
                    const uint64_t a = stage[sidx - pass * WR_CAP];
                    put(odst + (sidx - osrel), a_nbr(a), ov, a_w(a));
                }
            }
            s0 = s1;
            __syncwarp();
        }
        // ---- cooperative tier: the survivors the lanes did not take
        unsigned todo = __ballot_sync(RLAP_FULL_MASK, L > 0 && !staged);
        while (todo) {
            const int k = __ffs(todo) - 1;
            todo &= todo - 1;
            const int kL = __shfl_sync(RLAP_FULL_MASK, L, k), kf = __shfl_sync(RLAP_FULL_MASK, f, k);
            const int kv = __shfl_sync(RLAP_FULL_MASK, v, k);
            const long long kdst = __shfl_sync(RLAP_FULL_MASK, dst, k);
            if (kf == MERGE_PATH) {
                const long long ksrc = __shfl_sync(RLAP_FULL_MASK, src, k);
                for (int o = lane; o < kL; o += 32) {
                    const uint64_t a = __ldcs((const unsigned long long*)P.raw + ksrc + o);
                    put(kdst + o, a_nbr(a), kv, a_w(a));
                }
                continue;
            }
            const int kb = __shfl_sync(RLAP_FULL_MASK, b, k), knb = __shfl_sync(RLAP_FULL_MASK, nb, k);
            const long long kend = __shfl_sync(RLAP_FULL_MASK, end, k);
            const size_t vb = (size_t)(base + k - kv);
            // sorted fills, one per lane; a lane without one holds the largest id, so it never counts as "smaller"
            const uint64_t fa = lane < kf ? __ldcs((const unsigned long long*)P.raw + kend - kf + lane) : RLAP_PAD_A;
            const uint32_t fnbr = a_nbr(fa);
            int cb = 0;     // alive base entries with a smaller neighbour than this lane's fill
            int run = 0;    // alive base entries of the chunks done so far
            for (int p0 = 0; p0 < knb; p0 += 32) {
                const int p = p0 + lane;
                uint32_t c = 0xffffffffu;
                float w = 0.f;
                bool alive = false;
                if (p < knb) {
                    c = (uint32_t)__ldg(P.col + kb + p);
                    w = __ldg(P.w + kb + p);
                    alive = P.state[vb + c] != 2;
                }
                const unsigned mask = __ballot_sync(RLAP_FULL_MASK, alive);
                // fills with a smaller neighbour than this lane's base entry
                int lo = 0, hi = kf;
#pragma unroll
                for (int it = 0; it < 6; it++) {
                    const int mid = (lo + hi) >> 1;
                    const uint32_t fm = __shfl_sync(RLAP_FULL_MASK, fnbr, mid & 31);
                    if (lo < hi) { if (fm < c) lo = mid + 1; else hi = mid; }
                }
                if (alive) put(kdst + run + __popc(mask & lt) + lo, c, kv, w);
                // base entries of this chunk with a smaller neighbour than this lane's fill (the chunk is ascending)
                int lo2 = 0, hi2 = 32;
#pragma unroll
                for (int it = 0; it < 6; it++) {
                    const int mid = (lo2 + hi2) >> 1;
                    const uint32_t cm = __shfl_sync(RLAP_FULL_MASK, c, mid & 31);
                    if (lo2 < hi2) { if (cm < fnbr) lo2 = mid + 1; else hi2 = mid; }
                }
                cb += __popc(mask & (lo2 >= 32 ? 0xffffffffu : ((1u << lo2) - 1u)));
                run += __popc(mask);
            }
            if (lane < kf) put(kdst + lane + cb, fnbr, kv, a_w(fa));
            if (check && lane == 0 && run + kf != kL) mismatch(base + k, run + kf, kL);
        }
    }
}

// column pointers of every view: colptr[view * (n + 1) + v] = rows of `view` that precede column v (v = n: all rows)
__global__ void k_emit_colptr(SchurParams P, int* colptr) {
    const long long VN1 = (long long)P.V * (P.n + 1);
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= VN1) return;
    const long long view = t / (P.n + 1), v = t % (P.n + 1);
    colptr[t] = (int)(P.outoff[view * P.n + v] - P.outoff[view * P.n]);
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
cudaError_t eliminate_grid(int* blocks_out);

cudaError_t launch_emit_count(const SchurParams& P, long long* total_dev, cudaStream_t stream) {
    const size_t smem = (size_t)3 * CAP_CTA * sizeof(uint64_t);
    const size_t smem_big = (size_t)2 * CAP_BIG * sizeof(uint64_t);
    {
        static std::mutex mu;
        static PerDeviceOnce once;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (once.first(dev)) {
            cudaFuncSetAttribute(k_emit_sort_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_emit_sort_big, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_big);
        }
    }
    const long long VN = (long long)P.V * P.n;
    cudaError_t e = cudaMemsetAsync(P.ctr + CTR_EMIT_C0, 0, N_CLASS * sizeof(int), stream);
    if (e != cudaSuccess) return e;
    k_emit_prep<<<(unsigned)((VN + 255) / 256), 256, 0, stream>>>(P);
    e = launch_exclusive_scan<long long>(P.blk, VN, P.rawoff, P.blocksum, nullptr, stream);
    if (e != cudaSuccess) return e;
    {
        long long work = P.pool_cap / 2;
        long long bx = (work + 256 * 4 - 1) / (256 * 4);
        if (bx < 1) bx = 1;
        if (bx > 148 * 8) bx = 148 * 8;
        k_emit_scatter<<<dim3((unsigned)bx, (unsigned)P.V), 256, 0, stream>>>(P);
        bx = (VN + 255) / 256;            // one warp per 32 vertices
        if (bx < 1) bx = 1;
        if (bx > 148 * 16) bx = 148 * 16;
        k_emit_fsort<<<(unsigned)bx, 256, 0, stream>>>(P);
        bx = (VN / 4 * 32 + 255) / 256;   // one 8-lane tile per vertex
        if (bx < 1) bx = 1;
        if (bx > 148 * 16) bx = 148 * 16;
        k_emit_base<<<(unsigned)bx, 256, 0, stream>>>(P);
    }
    int blocks = 0;
    e = eliminate_grid(&blocks);
    if (e != cudaSuccess) return e;
    k_emit_sort_small<<<148 * 8, 256, 0, stream>>>(P);
    k_emit_sort_mid<2><<<148 * 8, 256, 0, stream>>>(P);
    k_emit_sort_mid<4><<<148 * 8, 256, 0, stream>>>(P);
    k_emit_sort_mid<8><<<148 * 4, 256, 0, stream>>>(P);
    k_emit_sort_mid<16><<<148 * 4, 256, 0, stream>>>(P);
    k_emit_sort_block<<<blocks, BLOCK_THREADS, smem, stream>>>(P);
    k_emit_sort_big<<<blocks / 2, BLOCK_THREADS, smem_big, stream>>>(P);
    return launch_exclusive_scan<long long>(P.outcnt, VN, P.outoff, P.blocksum, total_dev, stream);
}

cudaError_t launch_emit_colptr(const SchurParams& P, int* colptr, cudaStream_t stream) {
    const long long VN1 = (long long)P.V * (P.n + 1);
    k_emit_colptr<<<(unsigned)((VN1 + 255) / 256), 256, 0, stream>>>(P, colptr);
    return cudaGetLastError();
}

cudaError_t launch_emit_write(const SchurParams& P, int* out_row, int* out_col, float* out_w, double* out_f64,
                              cudaStream_t stream) {
    const long long VN = (long long)P.V * P.n;
    long long bx = (VN + 32 * WR_WARPS - 1) / (32 * WR_WARPS);  // one warp per 32 vertices
    if (bx < 1) bx = 1;
    if (bx > 148 * 24) bx = 148 * 24;
    k_emit_write<<<(unsigned)bx, WR_WARPS * 32, 0, stream>>>(P, out_row, out_col, out_w, out_f64);
    return cudaGetLastError();
}

}  // namespace rlap
