// schur.cu — ordering, elimination and emission kernels of the rLap randomized Schur complement.
//
// Replaces {Random,Priority,Coarsening}Preconditioner::getSchurComplement
// (rlap/csrc/preconditioner.cc:348-476, 713-825, 835-957) with a round-parallel design:
//   * the graph is an immutable coalesced CSR shared by all views; every view keeps only
//     per-vertex state, a head pointer per vertex and an append-only pool of fill entries
//     (lazy deletion: an entry dies when the vertex it points to is eliminated);
//   * one persistent cooperative kernel runs ALL rounds of the views of a view group (grid.sync between phases);
//     the groups of a call run as concurrent launches (api.cu);
//   * inside a round, vertices that are pairwise non-adjacent are eliminated concurrently, one warp
//     (or one thread block for big stars) per vertex: gather -> fixed-point quantise -> bitonic sort
//     by neighbour -> merge multi-edges -> o_n sort -> warp-shuffle prefix sums -> Philox-driven
//     binary-search sampling -> atomic append of the fill edges to both endpoints.
// The sequential specification this file implements bit for bit is oracle/rlap_oracle.cc (keyed mode)
// and DESIGN.md §3.
#include <cooperative_groups.h>
#include <stdio.h>
#include "rlap_device.cuh"
#include "schur.cuh"
#include "scan.cuh"
#include "star.cuh"

namespace cg = cooperative_groups;

namespace rlap {

__device__ __forceinline__ int graph_of(const SchurParams& P, int v) { return P.gid ? __ldg(P.gid + v) : 0; }

// key of the degree bucket queue (preconditioner.cc:125-246 restated): number of live
// list entries, never below 1 once the vertex had an edge (DegreePQDec is a no-op at key 1), 0 for
// vertices that were isolated from the start; computed inline where it is needed as
// (state == 4) ? 0 : max(live, 1).
// state byte: 0 kept (o_v = random, not eligible), 1 pending, 2 eliminated, 4 pending and isolated from the start

// ---------------------------------------------------------------------------------------------
// star staging
// ---------------------------------------------------------------------------------------------

// Gather the raw live entries of v into sb.A (unordered). Returns their count (group uniform);
// *wmaxb receives the bit pattern of the largest weight. Entries beyond sb.cap are counted, not stored.
template <bool CTA>
__device__ int star_gather(const SchurParams& P, int view, int v, StarBuf sb, CtaScratch* cs, uint32_t* wmaxb_out) {
    const size_t vb = (size_t)view * (size_t)P.n;
    const uint8_t* st = P.state + vb;
    const int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int b = __ldg(P.ptr + v), e = __ldg(P.ptr + v + 1);
    uint32_t wmaxb = 0;
    int cnt = 0;
    if (CTA) {
        if (threadIdx.x == 0) cs->icount = 0;
        __syncthreads();
    }
    // base CSR segment: coalesced
    const int stride = g_size<CTA>();
    for (int p0 = b; p0 < e; p0 += stride) {
        int p = p0 + g_rank<CTA>();
        bool ok = p < e;
        int u = 0;
        float w = 0.f;
        if (ok) {
            u = __ldg(P.col + p);
            ok = ldcg_i32(live_p(P, vb + u)) >= 0;   // an eliminated vertex carries RLAP_LIVE_DEAD
        }
        if (ok) w = __ldg(P.w + p);
        unsigned m = __ballot_sync(RLAP_FULL_MASK, ok);
        int base;
        if (CTA) {
            base = 0;
            if (lane == 0 && m) base = atomicAdd(&cs->icount, __popc(m));
            base = __shfl_sync(RLAP_FULL_MASK, base, 0);
        } else {
            base = cnt;
            cnt += __popc(m);
        }
        if (ok) {
            int pos = base + __popc(m & lt);
            if (pos < sb.cap) sb.A[pos] = pack_a((uint32_t)u, w);
            wmaxb = max(wmaxb, __float_as_uint(w));
        }
    }
    // appended fill entries: a linked list, walked by one warp; 32 hops are collected before the
    // (dependent) state lookups so that those run in parallel
    if (!CTA || (threadIdx.x >> 5) == 0) {
        int p = ldcg_i32(head_p(P, vb + v));
        while (p >= 0) {
            int4 mine = make_int4(0, 0, 0, 0);
            bool have = false;
            for (int k = 0; k < 32 && p >= 0; k++) {
                int4 en = __ldcg(pool + p);
                if (lane == k) { mine = en; have = true; }
                p = en.z;
            }
            bool ok = have && (ldcg_i32(live_p(P, vb + mine.x)) >= 0);
            unsigned m = __ballot_sync(RLAP_FULL_MASK, ok);
            int base;
            if (CTA) {
                base = 0;
                if (lane == 0 && m) base = atomicAdd(&cs->icount, __popc(m));
                base = __shfl_sync(RLAP_FULL_MASK, base, 0);
            } else {
                base = cnt;
                cnt += __popc(m);
            }
            if (ok) {
                int pos = base + __popc(m & lt);
                if (pos < sb.cap) sb.A[pos] = pack_a((uint32_t)mine.x, __int_as_float(mine.y));
                wmaxb = max(wmaxb, (uint32_t)mine.y);
            }
        }
    }
    if (CTA) {
        __syncthreads();
        cnt = cs->icount;
    }
    *wmaxb_out = g_max_u32<CTA>(wmaxb, cs);
    g_sync<CTA>();
    return cnt;
}

// degree / coarsen: appends to the next round's low list through a small per-warp shared buffer, so that the list
// tail is bumped once per few dozen entries. Every member is warp-collective.
constexpr int LOWBUF = 64;
struct LowAppender {
    unsigned int* buf = nullptr;   // [LOWBUF] shared memory, one per warp
    unsigned int* dst = nullptr;   // nullptr: disabled (o_v = random)
    long long cap = 0;
    int* tail = nullptr;
    int* ovf = nullptr;
    int fill = 0;
    __device__ __forceinline__ void flush() {
        if (fill == 0) return;
        const int lane = threadIdx.x & 31;
        int pos0 = 0;
        if (lane == 0) pos0 = atomicAdd(tail, fill);
        pos0 = __shfl_sync(RLAP_FULL_MASK, pos0, 0);
        __syncwarp();
        for (int i = lane; i < fill; i += 32) {
            if ((long long)pos0 + i < cap) dst[pos0 + i] = buf[i]; else *ovf = 1;
        }
        __syncwarp();
        fill = 0;
    }
    __device__ __forceinline__ void push(bool pred, unsigned int val) {
        if (dst == nullptr) return;
        const unsigned m = __ballot_sync(RLAP_FULL_MASK, pred);
        if (m == 0) return;
        if (fill + 32 > LOWBUF) flush();
        if (pred) buf[fill + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = val;
        fill += __popc(m);
    }
};

// Results that are still on their way back when a register tile finishes a star: the `next` fields of the two pool
// entries of its fill edge (from the list-head exchanges) and the old value of the live counter it decremented.
// Consuming them is deferred until the tile has issued the loads of its next star, so the warp does not wait for
// those round trips (degree / coarsen: a grid barrier separates the writers of a list from its readers; o_v =
// random flushes before its fence).
struct PendingPush {
    int4 e0, e1;
    int4* p0 = nullptr;
    int lo_old = 0, lo_lim = 0, lo_M = 0x7fffffff;   // crossing test: lo_M < lo_old <= lo_lim
    unsigned int lo_idx = 0;
    __device__ __forceinline__ void flush(LowAppender& la) {   // warp-collective
        if (p0) { p0[0] = e0; p0[1] = e1; p0 = nullptr; }
        la.push(lo_old > lo_M && lo_old <= lo_lim, lo_idx);
        lo_M = 0x7fffffff;
    }
};

// fill edge (j,k,w): append to both endpoints; o_v = random also records the new dependency.
// LIVE: bump the live counters of both endpoints here (the register tiles apply net deltas instead).
// Returns false for an underflowed fill (weight 0: not created).
template <bool LIVE>
__device__ __forceinline__ bool push_fill(const SchurParams& P, size_t vb, int4* pool, int j, int k, float w,
                                          long long slot, PendingPush* pend = nullptr) {
    if (!(w > 0.f)) {  // underflowed fill: leave two tombstones so that the pool can be read linearly
        pool[slot] = make_int4(-1, 0, -1, -1);
        pool[slot + 1] = make_int4(-1, 0, -1, -1);
        return false;
    }
    {   // both list heads are exchanged before either entry is written: the two round trips overlap
        const int s0 = (int)slot, s1 = (int)slot + 1;
        const int n0 = atomicExch(head_p(P, vb + j), s0);
        const int n1 = atomicExch(head_p(P, vb + k), s1);
        if (pend) {
            pend->e0 = make_int4(k, __float_as_int(w), n0, j);
            pend->e1 = make_int4(j, __float_as_int(w), n1, k);
            pend->p0 = pool + s0;
        } else {
            pool[s0] = make_int4(k, __float_as_int(w), n0, j);
            pool[s1] = make_int4(j, __float_as_int(w), n1, k);
        }
        if (LIVE) { atomicAdd(live_p(P, vb + j), 1); atomicAdd(live_p(P, vb + k), 1); }
    }
    if (P.o_v == 0) {
        if (ldcg_u8(P.state + vb + j) == 1 && ldcg_u8(P.state + vb + k) == 1) {
            int rj = ldcg_i32(P.rank + vb + j), rk = ldcg_i32(P.rank + vb + k);
            if (rj < rk) atomicAdd(P.blk + vb + k, 1); else atomicAdd(P.blk + vb + j, 1);
        }
    }
    return true;
}

// Eliminate vertex v of `view` (A.2 clique sampling / A.4 coarsening / full clique), DESIGN.md §3.3.
struct LocalStats { unsigned long long fills = 0, raw = 0; int maxstar = 0; unsigned nsm = 0; };

template <bool CTA>
__device__ void eliminate_star(const SchurParams& P, const RoundCtx& rc, int view, int v, StarBuf sb, CtaScratch* cs,
                               LocalStats& ls, LowAppender& la) {
    const size_t vb = (size_t)view * (size_t)P.n;
    const int gs = g_size<CTA>(), r = g_rank<CTA>();
    const uint32_t view_id = P.view_base + (uint32_t)view;
    int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
    uint32_t wmaxb;
    int lraw = star_gather<CTA>(P, view, v, sb, cs, &wmaxb);
    if (lraw > sb.cap) {  // cannot happen for the smem tiers (callers check live); scratch tier: report
        if (r == 0) set_status(P, 6);
        lraw = 0;  // leave the vertex in place; the run is invalid anyway
        g_sync<CTA>();
        return;
    }
    int P2 = 0, L = 0;
    if (lraw > 0) {
        const int shift = star_shift(__uint_as_float(wmaxb), lraw);
        L = star_sort_merge<CTA>(sb, lraw, shift, cs, &P2);
        const bool full = (P.flags & 1) != 0;
        const bool coarsen = (P.o_v == 2) && !full;
        const int on = coarsen ? 2 : P.o_n;
        if (!full && (on == 2 || L > 16)) {   // shuffle key; for asc / desc the tie-break of stars with > 16 neighbours
            for (int i = r; i < lraw; i += gs) {
                uint64_t a = sb.A[i];
                if (!a_dead(a)) {
                    uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, a_nbr(a), view_id, TAG_STAR);
                    sb.K[i] = ((uint64_t)x.z << 32) | (uint64_t)x.w;
                }
            }
            g_sync<CTA>();
        }
        if (full || on == 0) g_bitonic_sort<CTA, SORT_ASC>(sb, P2);
        else if (on == 1) g_bitonic_sort<CTA, SORT_DESC>(sb, P2);
        else g_bitonic_sort<CTA, SORT_KEY>(sb, P2);
        g_incl_scan_u64<CTA>(sb.Q, sb.K, L, cs);
        const unsigned long long S = sb.K[L - 1];
        // ---- fill edges
        long long nf = full ? (long long)L * (L - 1) / 2 : (long long)(L - 1);
        long long slot0 = 0;
        bool ovf = false;
        if (nf > 0) {
            if (CTA) {
                if (threadIdx.x == 0) cs->carry = atomicAdd(P.pool_cursor + view, (unsigned long long)(2 * nf));
                __syncthreads();
                slot0 = (long long)cs->carry;
            } else {
                unsigned long long s0 = 0;
                if (r == 0) s0 = atomicAdd(P.pool_cursor + view, (unsigned long long)(2 * nf));
                slot0 = (long long)__shfl_sync(RLAP_FULL_MASK, s0, 0);
            }
            ovf = slot0 + 2 * nf > P.pool_cap;
            if (ovf && r == 0) set_status(P, 5);
        }
        if (nf > 0 && !ovf) {
            if (full) {
                const double Sf = __dmul_rn(__ull2double_rn(S), pow2d(-shift));
                for (int a = 0; a < L - 1; a++) {
                    const uint64_t ea = sb.A[a];
                    const long long off = (long long)a * (2LL * L - a - 1) / 2;
                    for (int b2 = a + 1 + r; b2 < L; b2 += gs) {
                        const uint64_t eb = sb.A[b2];
                        float w = __double2float_rn(__ddiv_rn(__dmul_rn((double)a_w(ea), (double)a_w(eb)), Sf));
                        push_fill<true>(P, vb, pool, (int)a_nbr(ea), (int)a_nbr(eb), w, slot0 + 2 * (off + (b2 - a - 1)));
                    }
                }
            } else if (coarsen) {
                uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, 0xffffffffu, view_id, TAG_PICK);
                unsigned long long u = ((unsigned long long)x.x << 32) | (unsigned long long)x.y;
                unsigned long long rr = __umul64hi(u, S);
                int koff = upper_bound_u64(sb.K, L, rr);
                if (koff >= L) koff = L - 1;
                const uint64_t ek = sb.A[koff];
                const double wk = (double)a_w(ek);
                for (int m = r; m < L; m += gs) {
                    if (m == koff) continue;
                    const uint64_t em = sb.A[m];
                    const double wm = (double)a_w(em);
                    float w = __double2float_rn(__ddiv_rn(__dmul_rn(wk, wm), __dadd_rn(wk, wm)));
                    int sl = m < koff ? m : m - 1;
                    push_fill<true>(P, vb, pool, (int)a_nbr(em), (int)a_nbr(ek), w, slot0 + 2LL * sl);
                }
            } else {
                for (int m = r; m < L - 1; m += gs) {
                    const uint64_t em = sb.A[m];
                    const unsigned long long Cm = sb.K[m];
                    const unsigned long long rem = S - Cm;
                    uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, a_nbr(em), view_id, TAG_STAR);
                    unsigned long long u = ((unsigned long long)x.x << 32) | (unsigned long long)x.y;
                    unsigned long long rr = Cm + __umul64hi(u, rem);
                    int koff = upper_bound_u64(sb.K, L, rr);
                    if (koff >= L) koff = L - 1;
                    float w = __double2float_rn(
                        __ddiv_rn(__dmul_rn((double)a_w(em), __ull2double_rn(rem)), __ull2double_rn(S)));
                    push_fill<true>(P, vb, pool, (int)a_nbr(em), (int)a_nbr(sb.A[koff]), w, slot0 + 2LL * m);
                }
            }
        }
        // every push (and, for o_v = random, every dependency increment) is ordered before the
        // decrements below: a neighbour's counter can only reach zero once all its lower-ranked
        // eventual neighbours are gone
        if (P.o_v == 0) __threadfence();
        g_sync<CTA>();
        // a neighbour whose live counter moves from above the segment's level to the level or below joins the
        // next round's low list (degree / coarsen)
        const int M = (P.o_v != 0) ? ldcg_i32(P.lvl + (size_t)view * P.G + graph_of(P, v)) : -1;
        for (int base = 0; base < P2; base += gs) {
            const int i = base + r;
            bool cross = false;
            int u = 0;
            if (i < P2) {
                const uint64_t a = sb.A[i];
                if (a != RLAP_PAD_A) {
                    u = (int)a_nbr(a);
                    const int old = atomicSub(live_p(P, vb + u), 1);
                    cross = old > M && old - 1 <= M;
                    if (P.o_v == 0 && ldcg_u8(P.state + vb + u) == 1) {
                        int oldb = atomicSub(P.blk + vb + u, 1);
                        if (oldb == 1) {
                            int pos = rc.wl_base + atomicAdd(P.ctr + rc.wslot, 1);
                            P.wl[pos] = (unsigned int)(vb + (size_t)u);
                        }
                    }
                }
            }
            la.push(cross, (unsigned int)(vb + (size_t)u));
        }
        if (r == 0) {
            ls.fills += (unsigned long long)(ovf ? 0 : nf);
            ls.maxstar = max(ls.maxstar, L);
            ls.raw += (unsigned long long)lraw;
        }
    }
    if (r == 0) { P.state[vb + v] = 2; *live_p(P, vb + v) = RLAP_LIVE_DEAD; }
    g_sync<CTA>();
}

// Register-resident elimination: a tile of W lanes (8, 16 or 32) holds one star, one entry per lane, and
// the 32 / W tiles of a warp run in lock step (all shuffles are tile-wide). `idx` is the tile's work
// item (uniform inside the tile) or 0xffffffff for an idle tile. The caller has already read the star's CSR
// bounds (b, nb) and walked its fill list into shared memory (`fills`, nfill entries), and guarantees
// nb + nfill <= W: the raw list always fits, nothing is retried.
// `slot0` / `nslots`: pool slots reserved for this star by the caller (an upper bound, 2 per possible fill).
template <int W>
__device__ void eliminate_star_tile(const SchurParams& P, const RoundCtx& rc, unsigned int idx, int b, int nb,
                                    const uint64_t* fills, int nfill, long long slot0, int nslots, int M,
                                    LocalStats& ls, PendingPush& pend, LowAppender& la) {
    typedef Tile<W> T;
    const int tl = T::tl();
    const bool active = idx != 0xffffffffu;
    const int view = active ? (int)(idx / (unsigned)P.n) : 0, v = active ? (int)(idx % (unsigned)P.n) : 0;
    const size_t vb = (size_t)view * (size_t)P.n;
    const uint32_t view_id = P.view_base + (uint32_t)view;
    int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
    const bool fail = false;
    uint64_t a = RLAP_PAD_A;
    if (active) {
        if (tl < nb) a = pack_a((uint32_t)__ldg(P.col + b + tl), __ldg(P.w + b + tl));
        else if (tl - nb < nfill) a = fills[tl - nb];
    }
    bool dead = false;
    // dead test on the neighbour's (live, head) record: the sector is the one its live counter and list head are
    // updated in further down
    if (a != RLAP_PAD_A) dead = ldcg_i32(live_p(P, vb + a_nbr(a))) < 0;
    // the previous star's pool entries: their `next` fields have arrived by now
    pend.flush(la);
    if (dead) a = RLAP_PAD_A;
    __syncwarp();
    const bool go = active && !fail;
    a = T::sort_u64(a);
    const bool rawvalid = a != RLAP_PAD_A;         // raw live entry (multi-edge duplicates included)
    const int rawnbr = (int)a_nbr(a);
    const int lraw = __popc(T::ballot(rawvalid));
    unsigned long long q;
    int shift;
    int mult;
    const unsigned hmask = T::merge_sorted(a, q, shift, true, mult);
    const int L = __popc(hmask);
    // live counters: a neighbour loses its entries to v (all `mult` of them) and gains one entry per fill it
    // receives. The extra multiplicity goes now, the rest is netted per merged neighbour after sampling.
    bool cross0 = false;
    if (go && ((hmask >> tl) & 1u) && mult > 1) {
        const int old = atomicSub(live_p(P, vb + rawnbr), mult - 1);
        cross0 = old > M && old - (mult - 1) <= M;
    }
    const bool live = (hmask >> tl) & 1u;
    const bool full = (P.flags & 1) != 0;
    const bool coarsen = (P.o_v == 2) && !full;
    const int on = coarsen ? 2 : P.o_n;
    uint64_t key = ~0ull, tie = 0;
    if (live) {
        uint64_t shuf = 0;
        if (!full && (on == 2 || (W == 32 && L > 16))) {
            uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, a_nbr(a), view_id, TAG_STAR);
            shuf = ((uint64_t)x.z << 32) | (uint64_t)x.w;
        }
        if (full || on == 0) { key = q; tie = shuf; }
        else if (on == 1) { key = ~q; tie = shuf; }
        else key = shuf;
    } else {
        a = RLAP_PAD_A;
        q = 0;
        tie = ~0ull;   // padding sorts after every live entry, also after one whose key is ~0 (q = 0 under desc)
    }
    if (W == 32) T::sort_ktaq(key, tie, a, q);   // only a 32-lane tile can hold more than 16 neighbours
    else T::sort_kaq(key, a, q);
    const unsigned long long C = T::incl_scan(q);
    const unsigned long long S = __shfl_sync(RLAP_FULL_MASK, C, (L > 0 ? L - 1 : 0), W);
    const long long nf = (L < 1) ? 0 : (full ? (long long)L * (L - 1) / 2 : (long long)(L - 1));
    const bool ovf = nslots > 0 && slot0 + nslots > P.pool_cap;
    if (go && ovf && tl == 0) set_status(P, 5);
    const bool emit = go && nf > 0 && !ovf;
    // reserved but unused slots (multi-edges merged): tombstones, the pool is read linearly at emission
    if (go && !ovf)
        for (long long u = 2 * nf + tl; u < nslots; u += W) pool[slot0 + u] = make_int4(-1, 0, -1, -1);
    int delta = -1;   // net change of live[neighbour] for the merged neighbour held by this lane
    if (full) {
        const double Sf = __dmul_rn(__ull2double_rn(S), pow2d(-shift));
        for (int b2 = 1; b2 < W; b2++) {
            const uint64_t eb = __shfl_sync(RLAP_FULL_MASK, a, b2, W);
            bool done = false;
            if (emit && b2 < L && tl < b2) {
                float w = __double2float_rn(__ddiv_rn(__dmul_rn((double)a_w(a), (double)a_w(eb)), Sf));
                long long off = (long long)tl * (2LL * L - tl - 1) / 2;
                done = push_fill<false>(P, vb, pool, (int)a_nbr(a), (int)a_nbr(eb), w, slot0 + 2 * (off + (b2 - tl - 1)));
            }
            unsigned dm = T::ballot(done);
            if (done) delta++;
            if (tl == b2) delta += __popc(dm);
        }
    } else if (coarsen) {
        uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, 0xffffffffu, view_id, TAG_PICK);
        unsigned long long u = ((unsigned long long)x.x << 32) | (unsigned long long)x.y;
        int koff = T::upper_bound(C, L, __umul64hi(u, S));
        if (koff >= L) koff = L - 1;
        if (koff < 0) koff = 0;
        const uint64_t ek = __shfl_sync(RLAP_FULL_MASK, a, koff, W);
        bool done = false;
        if (emit && tl < L && tl != koff) {
            const double wk = (double)a_w(ek), wm = (double)a_w(a);
            float w = __double2float_rn(__ddiv_rn(__dmul_rn(wk, wm), __dadd_rn(wk, wm)));
            int sl = tl < koff ? tl : tl - 1;
            done = push_fill<false>(P, vb, pool, (int)a_nbr(a), (int)a_nbr(ek), w, slot0 + 2LL * sl, &pend);
        }
        unsigned dm = T::ballot(done);
        if (done) delta++;
        if (tl == koff) delta += __popc(dm);
    } else {
        const bool act = tl < L - 1;
        unsigned long long rr = 0, rem = 0;
        if (act) {
            rem = S - C;
            uint4 x = philox4x32_10(P.k0, P.k1, (uint32_t)v, a_nbr(a), view_id, TAG_STAR);
            unsigned long long u = ((unsigned long long)x.x << 32) | (unsigned long long)x.y;
            rr = C + __umul64hi(u, rem);
        }
        int koff = T::upper_bound(C, L, rr);
        if (koff >= L) koff = L - 1;
        if (koff < 0) koff = 0;
        const uint64_t ek = __shfl_sync(RLAP_FULL_MASK, a, koff, W);
        if (emit && act) {
            float w = __double2float_rn(__ddiv_rn(__dmul_rn((double)a_w(a), __ull2double_rn(rem)), __ull2double_rn(S)));
            if (push_fill<false>(P, vb, pool, (int)a_nbr(a), (int)a_nbr(ek), w, slot0 + 2LL * tl, &pend)) {
                delta++;
                atomicAdd(live_p(P, vb + (int)a_nbr(ek)), 1);
            }
        }
    }
    // o_v = random: pushes and dependency increments are ordered before the decrements;
    // the other orders separate rounds by grid barriers
    if (P.o_v == 0) { pend.flush(la); __threadfence(); }
    __syncwarp();
    // neighbours whose live counter crossed the segment's level downwards join the next round's low list; the test
    // on the value the atomic returns is left to the next flush of `pend`
    if (go && tl < L && delta != 0) {
        const int old = atomicAdd(live_p(P, vb + (int)a_nbr(a)), delta);
        if (delta < 0) { pend.lo_old = old; pend.lo_M = M; pend.lo_lim = M - delta; pend.lo_idx = (unsigned int)(vb + (size_t)a_nbr(a)); }
    }
    la.push(cross0, (unsigned int)(vb + (size_t)rawnbr));
    if (go && rawvalid) {
        if (P.o_v == 0 && ldcg_u8(P.state + vb + rawnbr) == 1) {
            int old = atomicSub(P.blk + vb + rawnbr, 1);
            if (old == 1) {
                int pos = rc.wl_base + atomicAdd(P.ctr + rc.wslot, 1);
                P.wl[pos] = (unsigned int)(vb + (size_t)rawnbr);
            }
        }
    }
    if (go && tl == 0) {
        ls.fills += (unsigned long long)(ovf ? 0 : nf);
        ls.maxstar = max(ls.maxstar, L);
        ls.raw += (unsigned long long)lraw;
        P.state[vb + v] = 2;
        *live_p(P, vb + v) = RLAP_LIVE_DEAD;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// the persistent elimination kernel
// ---------------------------------------------------------------------------------------------

constexpr int FCAP = 8;   // fill entries per item that the chunk prologue stages in shared memory

// Run one tier: the items whose bit is set in `mask` (lane i holds item i of the warp's chunk) are handed
// to the 32 / W tiles of the warp, 32 / W at a time.
template <int W>
__device__ void run_tier(const SchurParams& P, const RoundCtx& rc, unsigned mask, unsigned int my_idx, int my_b,
                         int my_nb, int my_nfill, const uint64_t* fbuf, long long my_slot0, int my_nslots, int my_M,
                         LocalStats& ls, PendingPush& pend, LowAppender& la) {
    constexpr int TPW = 32 / W;
    const int lane = threadIdx.x & 31;
    const int tile = lane / W;
    while (mask) {
        // tile t takes the t-th pending item (lowest set bits first)
        unsigned rest = mask;
        unsigned src = 0xffffffffu;
#pragma unroll
        for (int t = 0; t < TPW; t++) {
            const unsigned took = rest ? (unsigned)(__ffs(rest) - 1) : 0xffffffffu;
            if (t == tile) src = took;
            rest &= rest - 1;          // 0 & anything stays 0
        }
        const int sl = (int)(src & 31);
        unsigned int idx = __shfl_sync(RLAP_FULL_MASK, my_idx, sl);
        const int b = __shfl_sync(RLAP_FULL_MASK, my_b, sl), nb = __shfl_sync(RLAP_FULL_MASK, my_nb, sl);
        const int nfill = __shfl_sync(RLAP_FULL_MASK, my_nfill, sl);
        const long long sl0 = __shfl_sync(RLAP_FULL_MASK, my_slot0, sl);
        const int nsl = __shfl_sync(RLAP_FULL_MASK, my_nslots, sl);
        const int M = __shfl_sync(RLAP_FULL_MASK, my_M, sl);
        if (src == 0xffffffffu) idx = 0xffffffffu;
        eliminate_star_tile<W>(P, rc, idx, b, nb, fbuf + sl * FCAP, nfill, sl0, nsl, M, ls, pend, la);
        mask = rest;
    }
}

// process work-list items [start, end): a warp takes a chunk of up to 32 items and serves them tier by
// tier (8-, 16-, 32-lane register tiles, then the shared-memory path); big stars go to the block phase
__device__ void run_warp_items(const SchurParams& P, const RoundCtx& rc, uint64_t* smem, CtaScratch* cs, int* next,
                               int start, int end, LocalStats& ls, LowAppender& la) {
    const int gw = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nw = (int)((gridDim.x * blockDim.x) >> 5);
    const int lane = threadIdx.x & 31;
    StarBuf sb = warp_buf(smem);
    uint64_t* fbuf = sb.A;   // 32 x FCAP staged fill entries; the shared-memory path reuses the area afterwards
    const int count = end - start;
    if (count <= 0) return;
    // All warps of the launch (a view group: ~150 warps) fetch chunks of the round's work list from one global cursor,
    // with a guided size: 1 / (2 x warps) of what is left, at least 4 and at most 32 items. Blocks share their SMs
    // with the blocks of other groups and run at different speeds; per-block slices left a third of the phase's warp
    // time waiting for the slowest block (wait timers, profiles/README.md). A stale read of the cursor only changes a
    // chunk size. flags & 256: static stride, for comparison.
    const bool dynamic = (P.flags & 256) == 0;
    int chunk = 32, c0 = 0, cstep = 0, chunk_now = 0;
    const int my_end = end;
    if (!dynamic) {
        chunk = (count + nw - 1) / nw;       // spread small rounds over all warps
        if (chunk > 32) chunk = 32;
        c0 = start + gw * chunk; cstep = nw * chunk;
    }
    (void)next;
    for (;; c0 += cstep) {
        if (dynamic) {
            if (lane == 0) {
                const int left = count - ldcg_i32(P.ctr + rc.sslot);
                int c = left / (2 * nw);
                c = (c + 3) & ~3;
                c = c < 4 ? 4 : (c > 32 ? 32 : c);
                c0 = start + atomicAdd(P.ctr + rc.sslot, c);
                chunk_now = c;
            }
            c0 = __shfl_sync(RLAP_FULL_MASK, c0, 0);
            chunk_now = __shfl_sync(RLAP_FULL_MASK, chunk_now, 0);
        }
        if (c0 >= my_end) break;
        const int end = my_end;
        const int it = c0 + lane;
        unsigned int idx = 0xffffffffu;
        int lv = -1, cls = -1, b = 0, nb = 0, nfill = 0, M = -1;
        __syncwarp();   // the previous chunk is done with the staging buffer
        if (lane < (dynamic ? chunk_now : chunk) && it < end) {
            idx = __ldcg(P.wl + it);
            const int view = (int)(idx / (unsigned)P.n), v = (int)(idx % (unsigned)P.n);
            bool skip = false;
            if (P.o_v != 0) {  // truncated final round of a graph: only the highest ids go
                size_t seg = (size_t)view * P.G + graph_of(P, v);
                skip = ldcg_i32(P.ovfseg + seg) && idx < __ldcg(P.thresh + seg);
                M = ldcg_i32(P.lvl + seg);
            }
            if (!skip) {
                lv = ldcg_i32(live_p(P, idx));
                b = __ldg(P.ptr + v);
                nb = __ldg(P.ptr + v + 1) - b;
                if (lv <= CAP_WARP) {
                    // every lane walks the fill list of its own item (32 chains in flight) into the staging buffer;
                    // with the CSR bounds this gives the exact raw length, i.e. the tile width that holds the star
                    const int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
                    int p = ldcg_i32(head_p(P, idx));
                    while (p >= 0 && nfill < FCAP) {
                        const int4 en = __ldcg(pool + p);
                        fbuf[lane * FCAP + nfill] = pack_a((uint32_t)en.x, __int_as_float(en.y));
                        nfill++;
                        p = en.z;
                    }
                    cls = (p >= 0 || nb + nfill > 32) ? 33 : nb + nfill;   // 33: the shared-memory path gathers it itself
                    if (P.flags & 64) cls = 33;   // debug: no register tiles
                } else {
                    cls = lv;
                }
            }
        }
        __syncwarp();
        if (lv > CAP_WARP) {
            int pos = rc.dl_base + atomicAdd(P.ctr + rc.dslot, 1);
            P.dl[pos] = idx;
        }
        unsigned m8 = __ballot_sync(RLAP_FULL_MASK, cls >= 0 && cls <= 8);
        unsigned m16 = __ballot_sync(RLAP_FULL_MASK, cls > 8 && cls <= 16);
        unsigned m32 = __ballot_sync(RLAP_FULL_MASK, cls > 16 && cls <= 32);
        unsigned msm = __ballot_sync(RLAP_FULL_MASK, cls > 32 && lv <= CAP_WARP);
        // pool slots for the register tiers are reserved once per chunk: 2 per possible fill (L <= live), one
        // atomic per view present in the chunk instead of one per star
        const bool full = (P.flags & 1) != 0;
        int nslots = 0;
        if (lv >= 2 && cls <= 32) nslots = full ? lv * (lv - 1) : 2 * (lv - 1);
        long long slot0 = 0;
        {
            const int view = (idx == 0xffffffffu) ? -1 : (int)(idx / (unsigned)P.n);
            int incl = nslots;  // inclusive prefix over the lanes of the chunk
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(RLAP_FULL_MASK, incl, d);
                if (lane >= d) incl += t;
            }
            const unsigned need = __ballot_sync(RLAP_FULL_MASK, nslots > 0);
            if (need) {
                const int first = __ffs(need) - 1, last = 31 - __clz(need);
                const int v0 = __shfl_sync(RLAP_FULL_MASK, view, first);
                const bool same = __all_sync(RLAP_FULL_MASK, nslots == 0 || view == v0);
                if (same) {
                    const int total = __shfl_sync(RLAP_FULL_MASK, incl, last);
                    unsigned long long b0 = 0;
                    if (lane == first) b0 = atomicAdd(P.pool_cursor + v0, (unsigned long long)total);
                    b0 = __shfl_sync(RLAP_FULL_MASK, b0, first);
                    slot0 = (long long)b0 + incl - nslots;
                } else if (nslots > 0) {
                    slot0 = (long long)atomicAdd(P.pool_cursor + view, (unsigned long long)nslots);
                }
            }
        }
        PendingPush pend;
        run_tier<8>(P, rc, m8, idx, b, nb, nfill, fbuf, slot0, nslots, M, ls, pend, la);
        run_tier<16>(P, rc, m16, idx, b, nb, nfill, fbuf, slot0, nslots, M, ls, pend, la);
        run_tier<32>(P, rc, m32, idx, b, nb, nfill, fbuf, slot0, nslots, M, ls, pend, la);
        pend.flush(la);
        __syncwarp();
        if (lane == 0) ls.nsm += (unsigned)__popc(msm);
        while (msm) {
            int k = __ffs(msm) - 1;
            msm &= msm - 1;
            unsigned int kidx = __shfl_sync(RLAP_FULL_MASK, idx, k);
            eliminate_star<false>(P, rc, (int)(kidx / (unsigned)P.n), (int)(kidx % (unsigned)P.n), sb, cs, ls, la);
        }
    }
}

// deferred items [start, end): one block per item in shared memory; stars beyond CAP_CTA go to the
// NSLOT blocks that own a global scratch slot
__device__ void run_block_items(const SchurParams& P, const RoundCtx& rc, uint64_t* smem, CtaScratch* cs, int start,
                                int end, LocalStats& ls, LowAppender& la) {
    for (int it = start + (int)blockIdx.x; it < end; it += (int)gridDim.x) {
        unsigned int idx = __ldcg(P.dl + it);
        int view = (int)(idx / (unsigned)P.n), v = (int)(idx % (unsigned)P.n);
        if (ldcg_i32(live_p(P, idx)) <= CAP_CTA) eliminate_star<true>(P, rc, view, v, cta_buf(smem), cs, ls, la);
        __syncthreads();
    }
    const int nslot = min(NSLOT, (int)gridDim.x);   // a view group may run on fewer blocks than there are slots
    if ((int)blockIdx.x < nslot) {
        int j = 0;
        for (int it = start; it < end; it++) {
            unsigned int idx = __ldcg(P.dl + it);
            if (ldcg_i32(live_p(P, idx)) <= CAP_CTA) continue;
            if ((j++ % nslot) != (int)blockIdx.x) continue;
            int view = (int)(idx / (unsigned)P.n), v = (int)(idx % (unsigned)P.n);
            eliminate_star<true>(P, rc, view, v, scratch_buf(P), cs, ls, la);
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS, 2) k_eliminate(SchurParams P) {
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ CtaScratch cs;
    __shared__ int s_next, s_nsel;
    __shared__ unsigned int s_lowbuf[WARPS_PER_BLOCK][LOWBUF];
    cg::grid_group grid = cg::this_grid();
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nthr = (long long)gridDim.x * blockDim.x;
    const long long VN = (long long)P.V * P.n;
    const long long VG = (long long)P.V * P.G;
    const bool random_order = (P.o_v == 0);

    // ---- init: per-vertex state (ordering kernel for o_v = random: keyed Feistel rank)
    for (long long idx = tid; idx < VN; idx += nthr) {
        int view = (int)(idx / P.n), v = (int)(idx % P.n);
        *live_p(P, idx) = __ldg(P.ptr + v + 1) - __ldg(P.ptr + v);
        *head_p(P, idx) = -1;
        if (random_order) {
            int g = graph_of(P, v);
            int gb = __ldg(P.gptr + g), ng = __ldg(P.gptr + g + 1) - gb;
            RankPerm rp;
            uint32_t oview = (P.flags & 2) ? 0u : (P.view_base + (uint32_t)view);
            rp.init(P.k0, P.k1, (uint32_t)g, oview, (uint32_t)ng);
            int rk = (int)rp.rank((uint32_t)(v - gb));
            P.rank[idx] = rk;
            P.state[idx] = rk < __ldg(P.teff + g) ? 1 : 0;
        } else {
            P.state[idx] = (__ldg(P.ptr + v + 1) == __ldg(P.ptr + v)) ? 4 : 1;
            P.candround[idx] = -1;
            P.rank[idx] = -1;   // degree / coarsen: round stamp of the low list
            P.outoff[idx] = 0;  // (round, key) snapshot of phase A (rounds are stored + 1: 0 = never visited)
        }
    }
    if (!random_order) {
        for (long long s = tid; s < VG; s += nthr) {
            int g = (int)(s % P.G);
            P.rem[s] = __ldg(P.teff + g);
            P.lvl[s] = -1;
            P.minkey[s] = 0x7fffffff;
            P.minkey[VG + s] = 0x7fffffff;
            P.cntI[s] = 0;
            P.ovfseg[s] = 0;
        }
    }
    for (long long s = tid; s < P.V; s += nthr) P.pool_cursor[s] = 0ull;
    grid.sync();

    // phase timing (block 0, thread 0; nanoseconds between grid barriers, waits included)
    unsigned long long tmark = 0;
    auto lap = [&](int slot) {
        if (tid == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (slot >= 0) P.stats[slot] += now - tmark;
            tmark = now;
        }
    };
    lap(-1);
    // grid barrier that ends a phase; with flags & 128 every warp also records how long it waited there
    const bool wait_timers = (P.flags & 128) != 0;
    auto gsync = [&](int slot) {
        unsigned long long t0 = 0;
        if (wait_timers && (threadIdx.x & 31) == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        grid.sync();
        if (wait_timers && (threadIdx.x & 31) == 0) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            atomicAdd(P.stats + ST_W_INIT + (slot - ST_T_INIT), t1 - t0);
        }
        lap(slot);
    };
    // debug (flags & 512): thread 0 of block 0 times the parts of phase B
    unsigned long long dmark = 0;
    auto dlap = [&](int slot) {
        if ((P.flags & 512) && tid == 0) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (slot >= 0) P.stats[ST_DBG + slot] += now - dmark;
            dmark = now;
        }
    };
    int wl_start = 0;   // first unconsumed work-list item
    int dl_start = 0;
    int rounds = 0;
    RoundCtx rc;
    LocalStats ls;
    LowAppender la;
    la.buf = s_lowbuf[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const unsigned uVN = (unsigned)VN, un = (unsigned)P.n;

    if (random_order) {
        // dependency counters: pending lower-ranked eligible neighbours; roots seed the work list
        // (they count as appended in round -1, i.e. on counter slot 2)
        for (unsigned base = (unsigned)(tid - lane); base < uVN; base += (unsigned)nthr) {
            unsigned idx = base + lane;
            bool elig = idx < uVN && ldcg_u8(P.state + idx) == 1;
            int view = 0, v = 0, rv = 0, b = 0, e = 0;
            if (elig) {
                view = (int)(idx / un); v = (int)(idx % un);
                rv = ldcg_i32(P.rank + idx);
                b = __ldg(P.ptr + v); e = __ldg(P.ptr + v + 1);
            }
            const bool longrow = elig && (e - b) > 64;
            int c = 0;
            if (elig && !longrow) {
                size_t vb = (size_t)view * P.n;
                for (int p = b; p < e; p++) {
                    int u = __ldg(P.col + p);
                    if (ldcg_u8(P.state + vb + u) == 1 && ldcg_i32(P.rank + vb + u) < rv) c++;
                }
            }
            unsigned todo = __ballot_sync(RLAP_FULL_MASK, longrow);
            while (todo) {  // hubs: the warp walks the row together
                int k = __ffs(todo) - 1;
                todo &= todo - 1;
                int kb = __shfl_sync(RLAP_FULL_MASK, b, k), ke = __shfl_sync(RLAP_FULL_MASK, e, k);
                int krv = __shfl_sync(RLAP_FULL_MASK, rv, k);
                size_t kvb = (size_t)__shfl_sync(RLAP_FULL_MASK, view, k) * P.n;
                int cc = 0;
                for (int p = kb + lane; p < ke; p += 32) {
                    int u = __ldg(P.col + p);
                    if (ldcg_u8(P.state + kvb + u) == 1 && ldcg_i32(P.rank + kvb + u) < krv) cc++;
                }
                cc = __reduce_add_sync(RLAP_FULL_MASK, cc);
                if (lane == k) c = cc;
            }
            if (elig) P.blk[idx] = c;
            bool root = elig && c == 0;
            unsigned rm = __ballot_sync(RLAP_FULL_MASK, root);
            if (rm) {
                int pos0 = 0;
                if (lane == 0) pos0 = atomicAdd(P.ctr + CTR_WCNT0 + 2, __popc(rm));
                pos0 = __shfl_sync(RLAP_FULL_MASK, pos0, 0);
                if (root) P.wl[pos0 + __popc(rm & lt)] = idx;
            }
        }
        gsync(ST_T_INIT);
        while (true) {
            // items to consume were appended in the previous round
            int wl_end = wl_start + ldcg_i32(P.ctr + CTR_WCNT0 + (rounds + 2) % 3);
            if (wl_end == wl_start) break;
            if (tid == 0) { P.ctr[CTR_WCNT0 + (rounds + 1) % 3] = 0; P.ctr[CTR_DCNT0 + (rounds + 1) % 3] = 0; P.ctr[CTR_STEAL0 + (rounds + 1) % 3] = 0; }
            rc.sslot = CTR_STEAL0 + rounds % 3;
            rc.wl_base = wl_end; rc.wslot = CTR_WCNT0 + rounds % 3;
            rc.dl_base = dl_start; rc.dslot = CTR_DCNT0 + rounds % 3;
            run_warp_items(P, rc, smem, &cs, &s_next, wl_start, wl_end, ls, la);
            wl_start = wl_end;
            gsync(ST_T_D1);
            int dl_end = dl_start + ldcg_i32(P.ctr + rc.dslot);
            if (dl_end != dl_start) {
                run_block_items(P, rc, smem, &cs, dl_start, dl_end, ls, la);
                dl_start = dl_end;
                gsync(ST_T_D2);
            }
            rounds++;
        }
    } else {
        lap(ST_T_INIT);
        // degree / coarsen: rounds over the minimum-key bucket of every (view, graph) segment.
        // Full scans of the vertices are only made when a segment's level advances: lvl[seg] is the minimum key
        // the last full scan found, and every alive vertex whose key is <= lvl[seg] is on the round's low list
        // (bucket members that were blocked, listed vertices that were not in the bucket, and every vertex whose
        // live counter was seen crossing from above lvl to lvl or below by the elimination phase). While the low
        // list of a segment has an entry in play the bucket is taken from the list alone.
        constexpr int SEG_SM = 1024, WBUF = 224, MBUF = 160;
        const int INF = 0x7fffffff;
        // (round, key) snapshots written by phase A, read by the bucket test of phase B; the emission's row offsets
        // are not needed before the elimination is over
        unsigned long long* mark = (unsigned long long*)P.outoff;
        auto mark_of = [](int round, int key) { return ((unsigned long long)(unsigned)(round + 1) << 32) | (unsigned)key; };
        int* mkl = P.minkey;          // minimum over the low list
        int* mks = P.minkey + VG;     // minimum over a full scan
        const bool bsm = VG <= (long long)SEG_SM;
        while (true) {
            const int par = rounds & 1;
            const unsigned int* low_in = P.low + (size_t)par * (size_t)P.low_cap;
            la.dst = P.low + (size_t)(par ^ 1) * (size_t)P.low_cap;
            la.cap = P.low_cap;
            la.tail = P.ctr + CTR_LOW0 + (par ^ 1);
            la.ovf = P.ctr + CTR_LOWOVF0 + (par ^ 1);
            long long n_in = ldcg_i32(P.ctr + CTR_LOW0 + par);
            if (n_in > P.low_cap) n_in = P.low_cap;
            if (ldcg_i32(P.ctr + CTR_LOWOVF0 + par)) n_in = 0;   // entries were lost: every segment rescans
            rc.wl_base = wl_start; rc.wslot = CTR_WCNT0 + rounds % 3;
            rc.dl_base = dl_start; rc.dslot = CTR_DCNT0 + rounds % 3;
            rc.sslot = CTR_STEAL0 + rounds % 3;
            if (tid == 0) {
                P.ctr[CTR_ACTIVE0 + (par ^ 1)] = 0; P.ctr[CTR_OVF0 + (par ^ 1)] = 0;
                P.ctr[CTR_WCNT0 + (rounds + 1) % 3] = 0; P.ctr[CTR_DCNT0 + (rounds + 1) % 3] = 0;
                P.ctr[CTR_STEAL0 + (rounds + 1) % 3] = 0;
                P.ctr[CTR_LOW0 + (par ^ 1)] = 0; P.ctr[CTR_LOWOVF0 + (par ^ 1)] = 0;
            }
            for (long long s = tid; s < VG; s += nthr) { P.cntI[s] = 0; P.ovfseg[s] = 0; }
            // ---- phase A1: minimum key over the low-list entries in play
            {
                bool any = false;
                for (long long i0 = tid - lane; i0 < n_in; i0 += nthr) {
                    const long long i = i0 + lane;
                    bool valid = false;
                    int seg = 0, key = INF;
                    if (i < n_in) {
                        const unsigned idx = __ldcg(low_in + i);
                        const uint8_t st = ldcg_u8(P.state + idx);
                        if (st != 2) {
                            seg = (int)(idx / un) * P.G + graph_of(P, (int)(idx % un));
                            key = (st == 4) ? 0 : max(ldcg_i32(live_p(P, idx)), 1);
                            valid = ldcg_i32(P.rem + seg) > 0 && key <= ldcg_i32(P.lvl + seg);
                            if (valid) mark[idx] = mark_of(rounds, key);
                        }
                    }
                    const unsigned vm = __ballot_sync(RLAP_FULL_MASK, valid);
                    if (vm == 0) continue;
                    any = true;
                    const int leader = __ffs(vm) - 1;
                    const int seg0 = __shfl_sync(RLAP_FULL_MASK, seg, leader);
                    if (__all_sync(RLAP_FULL_MASK, !valid || seg == seg0)) {
                        const int k = __reduce_min_sync(RLAP_FULL_MASK, valid ? key : INF);
                        if (lane == leader) atomicMin(mkl + seg0, k);
                    } else if (valid) {
                        atomicMin(mkl + seg, key);
                    }
                }
                if (any && lane == 0) P.ctr[CTR_ACTIVE0 + par] = 1;
            }
            gsync(ST_T_A);
            if ((P.flags & 512) && tid == 0) {   // debug: list size and number of rescanning segments per round
                int ns = 0, mn = INF;
                for (long long q = 0; q < VG; q++) {
                    if (ldcg_i32(P.rem + q) > 0 && ldcg_i32(mkl + q) == INF) ns++;
                    mn = min(mn, ldcg_i32(mkl + q));
                }
                printf("round %d: low list %lld entries, %d of %lld segments rescan, min list key %d, lvl[0] %d rem[0] %d\n",
                       rounds, n_in, ns, VG, mn, P.lvl[0], P.rem[0]);
            }
            // ---- phase A2: segments with nothing in play on the list scan all their vertices for the minimum key
            {
                int* smin = (int*)smem;         // [VG] block-level minima
                int* sneed = smin + SEG_SM;     // [VG] 1 = the segment rescans
                int* sviews = sneed + SEG_SM;   // views with a rescanning segment (the scan visits only their vertices)
                int* sflag = sviews + SEG_SM;
                unsigned nsel = (unsigned)P.V;
                if (bsm) {
                    for (int q = threadIdx.x; q < P.V; q += blockDim.x) sflag[q] = 0;
                    if (threadIdx.x == 0) s_nsel = 0;
                    __syncthreads();
                    for (int q = threadIdx.x; q < (int)VG; q += blockDim.x) {
                        smin[q] = INF;
                        const int need = (ldcg_i32(P.rem + q) > 0 && ldcg_i32(mkl + q) == INF) ? 1 : 0;
                        sneed[q] = need;
                        if (need) sflag[q / P.G] = 1;
                    }
                    __syncthreads();
                    // ascending view order: every block must enumerate the selected vertices identically
                    for (int q = threadIdx.x; q < P.V; q += blockDim.x) {
                        if (sflag[q]) {
                            int pos = 0;
                            for (int j = 0; j < q; j++) pos += sflag[j];
                            sviews[pos] = q;
                            atomicAdd(&s_nsel, 1);
                        }
                    }
                    __syncthreads();
                    nsel = (unsigned)s_nsel;
                }
                bool any = false;
                // four independent elements per thread and iteration; every load is issued before the first use;
                // (selected view, vertex) of the running index are advanced incrementally: no division in the loop
                const unsigned total = nsel * un;
                const unsigned step_q = (unsigned)nthr / un, step_r = (unsigned)nthr % un;
                unsigned cview = (unsigned)tid / un, cv = (unsigned)tid % un;
                for (unsigned base = (unsigned)(tid - lane); base < total; base += 4u * (unsigned)nthr) {
                    uint8_t st4[4];
                    int lv4[4], sg4[4];
                    unsigned ix4[4];
                    bool nd4[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        const unsigned e = base + (unsigned)q4 * (unsigned)nthr + lane;
                        st4[q4] = 2; lv4[q4] = 0; sg4[q4] = 0; nd4[q4] = false;
                        const unsigned sv = cview;
                        const int v = (int)cv;
                        cview += step_q; cv += step_r;
                        if (cv >= un) { cv -= un; cview++; }
                        if (e < total) {
                            const int view = bsm ? sviews[sv] : (int)sv;
                            const unsigned idx = (unsigned)view * un + (unsigned)v;
                            ix4[q4] = idx;
                            sg4[q4] = view * P.G + graph_of(P, v);
                            nd4[q4] = bsm ? (sneed[sg4[q4]] != 0)
                                          : (ldcg_i32(P.rem + sg4[q4]) > 0 && ldcg_i32(mkl + sg4[q4]) == INF);
                            if (nd4[q4]) {
                                st4[q4] = ldcg_u8(P.state + idx);
                                lv4[q4] = ldcg_i32(live_p(P, idx));
                            }
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        const bool valid = nd4[q4] && st4[q4] != 2;
                        const int seg = sg4[q4];
                        const int key = valid ? ((st4[q4] == 4) ? 0 : max(lv4[q4], 1)) : INF;
                        if (valid) mark[ix4[q4]] = mark_of(rounds, key);
                        const unsigned vm = __ballot_sync(RLAP_FULL_MASK, valid);
                        if (vm == 0) continue;
                        any = true;
                        const int leader = __ffs(vm) - 1;
                        const int seg0 = __shfl_sync(RLAP_FULL_MASK, seg, leader);
                        if (__all_sync(RLAP_FULL_MASK, !valid || seg == seg0)) {
                            const int k = __reduce_min_sync(RLAP_FULL_MASK, key);
                            if (lane == leader) { if (bsm) atomicMin(smin + seg0, k); else atomicMin(mks + seg0, k); }
                        } else if (valid) {
                            if (bsm) atomicMin(smin + seg, key); else atomicMin(mks + seg, key);
                        }
                    }
                }
                if (bsm) {
                    __syncthreads();
                    for (int q = threadIdx.x; q < (int)VG; q += blockDim.x) {
                        const int k = smin[q];
                        if (k != INF) atomicMin(mks + q, k);
                    }
                    __syncthreads();
                }
                if (any && lane == 0) P.ctr[CTR_ACTIVE0 + par] = 1;
            }
            gsync(ST_T_A);
            if (ldcg_i32(P.ctr + CTR_ACTIVE0 + par) == 0) break;
            if (rounds > P.n + 8) { if (tid == 0) set_status(P, 10); break; }   // every round removes a vertex of every active segment
            // ---- phase B: members of the minimum bucket with no bucket neighbour of higher id are selected, the
            // other members (and the listed vertices below the level that are not in the bucket) go to the next list.
            // Shared memory (the star buffers are idle): per-segment copies if there are few enough segments, a
            // candidate buffer per warp so that the work-list tail is bumped once per few hundred candidates.
            dlap(-1);
            int* srem = (int*)smem;                 // remaining removals (0: the segment is done)
            int* smk = srem + SEG_SM;               // bucket key of the round (INF: nothing to do)
            int* scnt = smk + SEG_SM;               // candidates selected by this block
            int* slv = scnt + SEG_SM;               // level; -1 marks a segment that rescans in this round
            int* sviews = slv + SEG_SM;             // views with a segment that rescans in this round
            int* sflag = sviews + SEG_SM;
            unsigned int* wbuf = (unsigned int*)((int*)smem + 6 * SEG_SM) + (size_t)(threadIdx.x >> 5) * WBUF;
            int wfill = 0;
            unsigned nselB = (unsigned)P.V;
            // per-segment view of the round: bucket key m, scan flag, level
            auto seg_round = [&](int seg, int& rm, int& m, bool& scan, int& lv) {
                if (bsm) {
                    rm = srem[seg]; m = smk[seg]; lv = slv[seg]; scan = lv < 0;
                } else {
                    rm = ldcg_i32(P.rem + seg);
                    const int ml = ldcg_i32(mkl + seg);
                    scan = ml == INF;
                    m = scan ? ldcg_i32(mks + seg) : ml;
                    lv = scan ? -1 : ldcg_i32(P.lvl + seg);
                }
            };
            if (bsm) {
                for (int q = threadIdx.x; q < P.V; q += blockDim.x) sflag[q] = 0;
                if (threadIdx.x == 0) s_nsel = 0;
                __syncthreads();
                for (int q = threadIdx.x; q < (int)VG; q += blockDim.x) {
                    const int ml = ldcg_i32(mkl + q);
                    const bool scan = ml == INF;
                    srem[q] = ldcg_i32(P.rem + q);
                    smk[q] = scan ? ldcg_i32(mks + q) : ml;
                    slv[q] = scan ? -1 : ldcg_i32(P.lvl + q);
                    scnt[q] = 0;
                    if (scan && srem[q] > 0 && smk[q] != INF) sflag[q / P.G] = 1;
                }
                __syncthreads();
                for (int q = threadIdx.x; q < P.V; q += blockDim.x) {   // ascending view order, identical in every block
                    if (sflag[q]) {
                        int pos = 0;
                        for (int j = 0; j < q; j++) pos += sflag[j];
                        sviews[pos] = q;
                        atomicAdd(&s_nsel, 1);
                    }
                }
                __syncthreads();
                nselB = (unsigned)s_nsel;
            }
            auto flush = [&]() {
                if (wfill == 0) return;
                int pos0 = 0;
                if (lane == 0) pos0 = atomicAdd(P.ctr + rc.wslot, wfill);
                pos0 = __shfl_sync(RLAP_FULL_MASK, pos0, 0);
                __syncwarp();
                for (int i = lane; i < wfill; i += 32) P.wl[rc.wl_base + pos0 + i] = wbuf[i];
                __syncwarp();
                wfill = 0;
            };
            // bucket member v of `view` (key m): true if no alive neighbour of higher id has key m. Phase A left the
            // (round, key) snapshot of every vertex it visited in `mark`: all alive vertices of a rescanning segment,
            // all vertices in play otherwise, i.e. every alive vertex whose key can equal m. One 8-byte load per
            // neighbour answers "alive bucket member?" (an eliminated or unvisited neighbour carries an older round).
            auto member_free = [&](unsigned idx, int view, int v, int m) -> bool {
                const size_t vb = (size_t)view * P.n;
                const unsigned long long want = mark_of(rounds, m);
                bool ok = true;
                // base neighbours, four at a time: ids first, then their snapshots together
                const int pb = __ldg(P.ptr + v), pe = __ldg(P.ptr + v + 1);
                for (int p = pb; p < pe && ok; p += 4) {
                    int u4[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) u4[k] = (p + k < pe) ? __ldg(P.col + p + k) : -1;
                    unsigned long long k4[4];
#pragma unroll
                    for (int k = 0; k < 4; k++) k4[k] = (u4[k] > v) ? __ldcg(mark + vb + u4[k]) : 0ull;
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (u4[k] > v && k4[k] == want) ok = false;
                }
                const int4* pool = P.pool + (size_t)view * (size_t)P.pool_cap;
                for (int p = ldcg_i32(head_p(P, idx)); p >= 0 && ok;) {
                    int4 en = __ldcg(pool + p);
                    if (en.x > v && __ldcg(mark + vb + en.x) == want) ok = false;
                    p = en.z;
                }
                return ok;
            };
            // warp-collective: record the selected members (`cand`) of one pass over 32 vertices
            auto select = [&](bool cand, unsigned idx, int seg, int rm) {
                const unsigned cm = __ballot_sync(RLAP_FULL_MASK, cand);
                if (cm == 0) return;
                if (wfill + 32 > WBUF) flush();
                if (cand) {
                    P.candround[idx] = rounds;
                    wbuf[wfill + __popc(cm & lt)] = idx;
                    if (bsm) {
                        atomicAdd(scnt + seg, 1);
                    } else {
                        int c = atomicAdd(P.cntI + seg, 1);
                        if (c + 1 > rm) P.ctr[CTR_OVF0 + par] = 1;
                    }
                }
                wfill += __popc(cm);
            };
            // member buffer of the scan part: up to 32 left over + 128 new per iteration
            unsigned int* mbuf = (unsigned int*)((int*)smem + 6 * SEG_SM + WARPS_PER_BLOCK * WBUF) + (size_t)(threadIdx.x >> 5) * MBUF;
            int mfill = 0;
            auto test_members = [&](int first, int count) {      // warp-collective: lane i tests member first + i
                bool cand = false, keep = false;
                unsigned idx = 0;
                int seg = 0, rm = 0;
                if (lane < count) {
                    idx = mbuf[first + lane];
                    const int view = (int)(idx / un), v = (int)(idx % un);
                    seg = view * P.G + graph_of(P, v);
                    int m, lvq;
                    bool scan;
                    seg_round(seg, rm, m, scan, lvq);
                    if (member_free(idx, view, v, m)) cand = true; else keep = true;
                }
                __syncwarp();
                select(cand, idx, seg, rm);
                la.push(keep, idx);
            };
            dlap(0);
            // B1: the low list (segments that do not rescan)
            for (long long i0 = tid - lane; i0 < n_in; i0 += nthr) {
                const long long i = i0 + lane;
                bool cand = false, keep = false;
                unsigned idx = 0;
                int seg = 0, rm = 0;
                if (i < n_in) {
                    idx = __ldcg(low_in + i);
                    const uint8_t st = ldcg_u8(P.state + idx);
                    if (st != 2) {
                        const int view = (int)(idx / un), v = (int)(idx % un);
                        seg = view * P.G + graph_of(P, v);
                        int m, lv;
                        bool scan;
                        seg_round(seg, rm, m, scan, lv);
                        const int key = (st == 4) ? 0 : max(ldcg_i32(live_p(P, idx)), 1);
                        // in play, and the first copy of this vertex on the list (P.rank holds the round stamp)
                        if (rm > 0 && !scan && key <= lv && atomicExch(P.rank + idx, rounds) != rounds) {
                            if (key == m && member_free(idx, view, v, m)) cand = true; else keep = true;
                        }
                    }
                }
                select(cand, idx, seg, rm);
                la.push(keep, idx);
            }
            dlap(1);
            // B2: full scan of the segments that rescan; the level moves to the bucket key found
            for (long long q = tid; q < VG; q += nthr) {
                if (ldcg_i32(mkl + q) == INF) { const int m2 = ldcg_i32(mks + q); if (m2 != INF) P.lvl[q] = m2; }
            }
            {
                const unsigned total = nselB * un;
                const unsigned step_q = (unsigned)nthr / un, step_r = (unsigned)nthr % un;
                unsigned cview = (unsigned)tid / un, cv = (unsigned)tid % un;
                for (unsigned base = (unsigned)(tid - lane); base < total; base += 4u * (unsigned)nthr) {
                    uint8_t st4[4];
                    int lv4[4], rm4[4], sg4[4], mk4[4], vw4[4], vx4[4];
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        const unsigned e = base + (unsigned)q4 * (unsigned)nthr + lane;
                        st4[q4] = 2; lv4[q4] = 0; rm4[q4] = 0; sg4[q4] = 0; mk4[q4] = -1;
                        const unsigned sv = cview;
                        const int v = (int)cv;
                        vw4[q4] = 0; vx4[q4] = v;
                        cview += step_q; cv += step_r;
                        if (cv >= un) { cv -= un; cview++; }
                        if (e < total) {
                            const int view = bsm ? sviews[sv] : (int)sv;
                            vw4[q4] = view;
                            const unsigned idx = (unsigned)view * un + (unsigned)v;
                            sg4[q4] = view * P.G + graph_of(P, v);
                            int lvq;
                            bool scan;
                            seg_round(sg4[q4], rm4[q4], mk4[q4], scan, lvq);
                            if (scan && rm4[q4] > 0 && mk4[q4] != INF) {
                                st4[q4] = ldcg_u8(P.state + idx);
                                lv4[q4] = ldcg_i32(live_p(P, idx));
                            } else {
                                mk4[q4] = -1;
                            }
                        }
                    }
                    // the bucket members among these 128 vertices go to the warp's member buffer; the tests run on full
                    // batches of 32 members, one per lane (testing in place left three quarters of the lanes idle while
                    // the warp waited for the longest list of each of the four passes)
#pragma unroll
                    for (int q4 = 0; q4 < 4; q4++) {
                        const unsigned idx = (unsigned)vw4[q4] * un + (unsigned)vx4[q4];
                        const int m = mk4[q4];
                        const bool member = m >= 0 && st4[q4] != 2 && ((st4[q4] == 4) ? 0 : max(lv4[q4], 1)) == m;
                        const unsigned mm = __ballot_sync(RLAP_FULL_MASK, member);
                        if (member) mbuf[mfill + __popc(mm & lt)] = idx;
                        mfill += __popc(mm);
                    }
                    __syncwarp();
                    while (mfill >= 32) { mfill -= 32; test_members(mfill, 32); }
                }
                if (mfill > 0) { test_members(0, mfill); mfill = 0; }
            }
            dlap(2);
            flush();
            la.flush();
            dlap(3);
            if (bsm) {
                __syncthreads();
                for (int q = threadIdx.x; q < (int)VG; q += blockDim.x) {
                    int c = scnt[q];
                    if (c > 0) {
                        int c0 = atomicAdd(P.cntI + q, c);
                        if (c0 + c > srem[q]) P.ctr[CTR_OVF0 + par] = 1;
                    }
                }
            }
            __syncthreads();   // the shared copies are star buffers again from here on
            dlap(4);
            gsync(ST_T_B);
            dlap(5);
            // phase C: a graph that selected more than it may still remove keeps its highest ids
            if (ldcg_i32(P.ctr + CTR_OVF0 + par)) {
                const int gw = (int)(tid >> 5), nw = (int)(nthr >> 5), lane = threadIdx.x & 31;
                const long long nblk = (VN + SEL_BLOCK - 1) / SEL_BLOCK;
                for (long long bk = gw; bk < nblk; bk += nw) {
                    int c = 0;
                    for (int j = 0; j < SEL_BLOCK / 32; j++) {
                        long long idx = bk * SEL_BLOCK + j * 32 + lane;
                        bool f = idx < VN && ldcg_i32(P.candround + idx) == rounds;
                        c += __popc(__ballot_sync(RLAP_FULL_MASK, f));
                    }
                    if (lane == 0) P.blockcnt[bk] = c;
                }
                grid.sync();
                for (long long s = gw; s < VG; s += nw) {
                    int need = ldcg_i32(P.rem + s);
                    if (ldcg_i32(P.cntI + s) <= need) continue;
                    int view = (int)(s / P.G), g = (int)(s % P.G);
                    long long lo = (long long)view * P.n + __ldg(P.gptr + g);
                    long long hi = (long long)view * P.n + __ldg(P.gptr + g + 1);  // exclusive
                    // walk down from hi in SEL_BLOCK-aligned pieces until `need` candidates are covered
                    long long cur = hi;
                    int acc = 0;
                    long long T = lo;
                    while (cur > lo) {
                        long long pb = ((cur - 1) / SEL_BLOCK) * SEL_BLOCK;  // aligned block holding cur-1
                        long long pstart = pb > lo ? pb : lo;
                        bool whole = (pstart == pb) && (cur == pb + SEL_BLOCK);
                        int c;
                        if (whole) {
                            c = ldcg_i32(P.blockcnt + pb / SEL_BLOCK);
                        } else {
                            c = 0;
                            for (long long q0 = pstart; q0 < cur; q0 += 32) {
                                long long q = q0 + lane;
                                bool f = q < cur && ldcg_i32(P.candround + q) == rounds;
                                c += __popc(__ballot_sync(RLAP_FULL_MASK, f));
                            }
                        }
                        if (acc + c >= need) {
                            // the threshold lies inside [pstart, cur): scan it from the top, 32 ids at a time
                            long long q1 = cur;
                            while (q1 > pstart) {
                                long long q0 = q1 - 32 > pstart ? q1 - 32 : pstart;
                                long long q = q0 + lane;
                                bool f = q < q1 && ldcg_i32(P.candround + q) == rounds;
                                unsigned mb = __ballot_sync(RLAP_FULL_MASK, f);
                                int cc = __popc(mb);
                                if (acc + cc >= need) {
                                    int want = need - acc;  // keep the `want` highest set bits of mb
                                    int bit = 31;
                                    for (;; bit--) {
                                        if (mb & (1u << bit)) { want--; if (want == 0) break; }
                                    }
                                    T = q0 + bit;
                                    acc = need;
                                    break;
                                }
                                acc += cc;
                                q1 = q0;
                            }
                            break;
                        }
                        acc += c;
                        cur = pstart;
                    }
                    if (lane == 0) { P.thresh[s] = (unsigned int)T; P.ovfseg[s] = 1; }
                }
                gsync(ST_T_C);
            }
            // the truncated count is what phase D will eliminate; the per-round minima are reset for the next round
            for (long long q = tid; q < VG; q += nthr) {
                int rm = ldcg_i32(P.rem + q), c = ldcg_i32(P.cntI + q);
                if (c > 0) P.rem[q] = rm - min(rm, c);
                mkl[q] = INF; mks[q] = INF;
            }
            // phase D: eliminate
            int wl_end = wl_start + ldcg_i32(P.ctr + rc.wslot);
            run_warp_items(P, rc, smem, &cs, &s_next, wl_start, wl_end, ls, la);
            wl_start = wl_end;
            la.flush();
            gsync(ST_T_D1);
            int dl_end = dl_start + ldcg_i32(P.ctr + rc.dslot);
            if (dl_end != dl_start) {
                run_block_items(P, rc, smem, &cs, dl_start, dl_end, ls, la);
                dl_start = dl_end;
                la.flush();
                gsync(ST_T_D2);
            }
            rounds++;
        }
    }
    if (tid == 0) P.ctr[CTR_ROUNDS] = rounds;
    if (ls.raw | ls.fills) {
        atomicAdd(P.stats + ST_FILLS, ls.fills);
        atomicAdd(P.stats + ST_RAW, ls.raw);
        atomicMax(P.stats + ST_MAXSTAR, (unsigned long long)ls.maxstar);
        if (ls.nsm) atomicAdd(P.stats + ST_DEFERRED, (unsigned long long)ls.nsm);
    }
}

// gid[v] = graph of vertex v, teff[g] = min(max(num_remove, 0), n_g - 1)
__global__ void k_setup_graphs(int n, int G, const int* gptr, const long long* num_remove, int* gid, int* teff) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < G) {
        long long ng = gptr[tid + 1] - gptr[tid];
        long long t = num_remove[tid];
        if (t < 0) t = 0;
        if (t > ng - 1) t = ng - 1;
        if (t < 0) t = 0;
        teff[tid] = (int)t;
    }
    if (gid && tid < n) {
        int lo = 0, hi = G;  // last g with gptr[g] <= v
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (gptr[mid] <= (int)tid) lo = mid; else hi = mid;
        }
        gid[tid] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from api.cu)
// ---------------------------------------------------------------------------------------------
static const size_t kSmemBytes = (size_t)3 * CAP_CTA * sizeof(uint64_t);

cudaError_t launch_setup_graphs(int n, int G, const int* gptr, const long long* num_remove, int* gid, int* teff,
                                cudaStream_t stream) {
    long long work = n > G ? n : G;
    int blocks = (int)((work + 255) / 256);
    if (blocks < 1) blocks = 1;
    k_setup_graphs<<<blocks, 256, 0, stream>>>(n, G, gptr, num_remove, gid, teff);
    return cudaGetLastError();
}

cudaError_t eliminate_grid(int* blocks_out) {
    static int cached = 0;
    if (!cached) {
        cudaError_t e = cudaFuncSetAttribute(k_eliminate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_eliminate, BLOCK_THREADS, kSmemBytes);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        cached = sms * occ;
    }
    *blocks_out = cached;
    return cudaSuccess;
}

cudaError_t launch_eliminate(const SchurParams& P, cudaStream_t stream, int blocks_req) {
    int blocks = 0;
    cudaError_t e = eliminate_grid(&blocks);
    if (e != cudaSuccess) return e;
    if (blocks_req > 0 && blocks_req < blocks) blocks = blocks_req;
    SchurParams Pc = P;
    void* args[] = {(void*)&Pc};
    return cudaLaunchCooperativeKernel((void*)k_eliminate, dim3(blocks), dim3(BLOCK_THREADS), args, kSmemBytes, stream);
}

// fold the control blocks of the view groups into the caller-visible one: first error, largest round count,
// summed counters, longest phase times
__global__ void k_combine_groups(int K, const int* gctr, const unsigned long long* gstats, int* ctr,
                                 unsigned long long* stats) {
    const int t = threadIdx.x;
    if (t == 0) {
        int status = 0, rounds = 0;
        for (int g = 0; g < K; g++) {
            const int* c = gctr + (size_t)g * CTR_COUNT;
            if (status == 0) status = c[CTR_STATUS];
            rounds = max(rounds, c[CTR_ROUNDS]);
        }
        ctr[CTR_STATUS] = status;
        ctr[CTR_ROUNDS] = rounds;
    }
    if (t < ST_COUNT) {
        unsigned long long acc = 0;
        for (int g = 0; g < K; g++) {
            const unsigned long long v = gstats[(size_t)g * ST_COUNT + t];
            if (t == ST_MAXSTAR || t >= ST_T_INIT) acc = v > acc ? v : acc; else acc += v;
        }
        stats[t] = acc;
    }
}

cudaError_t launch_combine_groups(int K, const int* gctr, const unsigned long long* gstats, int* ctr,
                                  unsigned long long* stats, cudaStream_t stream) {
    k_combine_groups<<<1, 32, 0, stream>>>(K, gctr, gstats, ctr, stats);
    return cudaGetLastError();
}

}  // namespace rlap
