// ingest.cu — COO -> coalesced CSR on the device.
//
// Replaces EdgeInfoMatrixReader::Read (rlap/csrc/reader.cc:42-61: drop zero weights, sum duplicate
// (row, col) in order of appearance, rows ascending inside a column) and the validation half of
// Factorizer::computeLaplacian (rlap/csrc/factorizers.cc:18-22: the adjacency must be symmetric).
// Entry (src -> dst, w) is stored in dst's segment, like the reference's column-compressed matrix.
// Pipeline: count per owner -> exclusive scan -> scatter (owner-major, unordered inside a segment) ->
// per-segment bitonic sort by (neighbour, input position) + duplicate merge -> scan -> compact copy.
#include "rlap_device.cuh"
#include "schur.cuh"
#include "scan.cuh"
#include <mutex>
#include "ingest.cuh"

namespace rlap {

__global__ void k_ingest_count(IngestParams P) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long nthr = (long long)gridDim.x * blockDim.x;
    for (long long p = tid; p < P.e; p += nthr) {
        long long s = P.src[p], d = P.dst[p];
        if (s < 0 || s >= P.n || d < 0 || d >= P.n) { atomicCAS(P.status, 0, 2); continue; }
        float w = P.w ? P.w[p] : 1.0f;
        if (w == 0.0f) continue;
        if (!(w > 0.0f) || isinf(w)) { atomicCAS(P.status, 0, 9); continue; }
        if (s == d) { atomicCAS(P.status, 0, 3); continue; }
        atomicAdd(P.cnt + d, 1);
    }
}

__global__ void k_ingest_scatter(IngestParams P) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long nthr = (long long)gridDim.x * blockDim.x;
    for (long long p = tid; p < P.e; p += nthr) {
        long long s = P.src[p], d = P.dst[p];
        if (s < 0 || s >= P.n || d < 0 || d >= P.n || s == d) continue;
        float w = P.w ? P.w[p] : 1.0f;
        if (!(w > 0.0f) || isinf(w)) continue;
        int pos = P.rawptr[d] + atomicAdd(P.cursor + d, 1);
        P.rkey[pos] = ((uint64_t)(uint32_t)s << 32) | (uint64_t)(uint32_t)p;
        P.rw[pos] = w;
    }
}

// sort one owner's segment by (neighbour, input position), merge duplicates (fp32 sum in input order),
// write the merged entries to tcol/tw at the segment's raw offset and the merged count to cnt2
template <bool CTA>
__device__ void ingest_row(const IngestParams& P, int v, StarBuf sb, CtaScratch* cs) {
    const int gs = g_size<CTA>(), r = g_rank<CTA>();
    const int b = P.rawptr[v], len = P.rawptr[v + 1] - b;
    if (len > sb.cap) {
        if (r == 0) { atomicCAS(P.status, 0, 6); P.cnt2[v] = 0; }
        return;
    }
    const int P2 = next_pow2(len);
    for (int i = r; i < P2; i += gs) {
        if (i < len) { sb.A[i] = P.rkey[b + i]; sb.Q[i] = (uint64_t)__float_as_uint(P.rw[b + i]); }
        else { sb.A[i] = RLAP_PAD_A; sb.Q[i] = 0; }
    }
    g_sync<CTA>();
    g_bitonic_sort<CTA, SORT_BY_A>(sb, P2);
    // heads of runs of equal neighbour: sequential fp32 sum, ordered output position
    const int lane = threadIdx.x & 31;
    int carry = 0;
    for (int base = 0; base < len; base += gs) {
        int i = base + r;
        bool act = i < len;
        uint32_t nb = act ? a_nbr(sb.A[i]) : 0xffffffffu;
        bool headf = act && (i == 0 || a_nbr(sb.A[i - 1]) != nb);
        float sum = 0.f;
        if (headf) {
            int j = i;
            do { sum += __uint_as_float((uint32_t)sb.Q[j]); j++; } while (j < len && a_nbr(sb.A[j]) == nb);
        }
        unsigned m = __ballot_sync(RLAP_FULL_MASK, headf);
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
        if (headf) { P.tcol[b + pos] = (int)nb; P.tw[b + pos] = sum; }
    }
    if (r == 0) P.cnt2[v] = carry;
    g_sync<CTA>();
}

__global__ void __launch_bounds__(BLOCK_THREADS, 2) k_ingest_rows_warp(IngestParams P) {
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ CtaScratch cs;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    int w = threadIdx.x >> 5;
    StarBuf sb;
    sb.A = smem + (size_t)w * 3 * CAP_WARP; sb.Q = sb.A + CAP_WARP; sb.K = sb.Q + CAP_WARP; sb.cap = CAP_WARP;
    for (long long base = gw * 32; base < P.n; base += nw * 32) {
        long long v = base + lane;
        int len = 0;
        if (v < P.n) {
            len = P.rawptr[v + 1] - P.rawptr[v];
            if (len == 0) P.cnt2[v] = 0;
            if (len > CAP_WARP) {
                int pos = atomicAdd(P.dl_tail, 1);
                P.dl[pos] = (unsigned int)v;
            }
        }
        unsigned todo = __ballot_sync(RLAP_FULL_MASK, len > 0 && len <= CAP_WARP);
        while (todo) {
            int k = __ffs(todo) - 1;
            todo &= todo - 1;
            ingest_row<false>(P, (int)(base + k), sb, &cs);
        }
    }
}

__global__ void __launch_bounds__(BLOCK_THREADS, 2) k_ingest_rows_block(IngestParams P) {
    extern __shared__ __align__(16) uint64_t smem[];
    __shared__ CtaScratch cs;
    StarBuf sb;
    sb.A = smem; sb.Q = smem + CAP_CTA; sb.K = smem + 2 * CAP_CTA; sb.cap = CAP_CTA;
    int end = *P.dl_tail;
    for (int it = (int)blockIdx.x; it < end; it += (int)gridDim.x) {
        int v = (int)P.dl[it];
        if (P.rawptr[v + 1] - P.rawptr[v] <= CAP_CTA) ingest_row<true>(P, v, sb, &cs);
        __syncthreads();
    }
    if ((int)blockIdx.x < NSLOT) {
        StarBuf gb;
        gb.A = P.scratch + (size_t)blockIdx.x * 3 * (size_t)P.scratch_cap;
        gb.Q = gb.A + P.scratch_cap; gb.K = gb.Q + P.scratch_cap; gb.cap = P.scratch_cap;
        int j = 0;
        for (int it = 0; it < end; it++) {
            int v = (int)P.dl[it];
            if (P.rawptr[v + 1] - P.rawptr[v] <= CAP_CTA) continue;
            if ((j++ % NSLOT) != (int)blockIdx.x) continue;
            ingest_row<true>(P, v, gb, &cs);
            __syncthreads();
        }
    }
}

// copy the merged entries of every owner to their final (gap-free) position
__global__ void k_ingest_compact(IngestParams P) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (long long v = gw; v < P.n; v += nw) {
        int sb = P.rawptr[v], db = P.ptr[v], len = P.ptr[v + 1] - db;
        for (int i = lane; i < len; i += 32) { P.col[db + i] = P.tcol[sb + i]; P.wout[db + i] = P.tw[sb + i]; }
    }
}

// symmetry: every entry (v <- u, w) needs a twin (u <- v, w'); accumulates sum (w - w')^2 and sum w^2
__global__ void k_ingest_symmetry(IngestParams P) {
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    double d2 = 0, n2 = 0;
    for (long long v = gw; v < P.n; v += nw) {
        int b = P.ptr[v], e = P.ptr[v + 1];
        for (int p = b + lane; p < e; p += 32) {
            int u = P.col[p];
            float w = P.wout[p];
            int lo = P.ptr[u], hi = P.ptr[u + 1];
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (P.col[mid] < (int)v) lo = mid + 1; else hi = mid;
            }
            if (lo < P.ptr[u + 1] && P.col[lo] == (int)v) {
                double d = (double)w - (double)P.wout[lo];
                d2 += d * d;
            } else {
                atomicCAS(P.status, 0, 4);
            }
            n2 += (double)w * (double)w;
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        d2 += __shfl_xor_sync(RLAP_FULL_MASK, d2, d);
        n2 += __shfl_xor_sync(RLAP_FULL_MASK, n2, d);
    }
    if (lane == 0 && (d2 != 0 || n2 != 0)) { atomicAdd(P.sym_acc, d2); atomicAdd(P.sym_acc + 1, n2); }
}

static int grid_for(long long work, int threads, int cap_blocks) {
    long long b = (work + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > cap_blocks) b = cap_blocks;
    return (int)b;
}

cudaError_t launch_ingest_stage1(const IngestParams& P, cudaStream_t stream) {
    k_ingest_count<<<grid_for(P.e, 256, 148 * 16), 256, 0, stream>>>(P);
    cudaError_t e = launch_exclusive_scan<int>(P.cnt, P.n, P.rawptr, P.blocksum, nullptr, stream);
    if (e != cudaSuccess) return e;
    k_ingest_scatter<<<grid_for(P.e, 256, 148 * 16), 256, 0, stream>>>(P);
    const size_t smem = (size_t)3 * CAP_CTA * sizeof(uint64_t);
    {
        static std::mutex mu;
        static PerDeviceOnce once;
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lk(mu);
        if (once.first(dev)) {
            cudaFuncSetAttribute(k_ingest_rows_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            cudaFuncSetAttribute(k_ingest_rows_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        }
    }
    int blocks = grid_for((P.n + 31) / 32 * 32, BLOCK_THREADS / 32 * 32, 148 * 2);
    k_ingest_rows_warp<<<blocks, BLOCK_THREADS, smem, stream>>>(P);
    k_ingest_rows_block<<<148 * 2, BLOCK_THREADS, smem, stream>>>(P);
    e = launch_exclusive_scan<int>(P.cnt2, P.n, P.ptr, P.blocksum, P.total_dev, stream);
    if (e != cudaSuccess) return e;
    k_ingest_compact<<<grid_for(P.n * 32, 256, 148 * 16), 256, 0, stream>>>(P);
    if (P.validate) k_ingest_symmetry<<<grid_for(P.n * 32, 256, 148 * 16), 256, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace rlap
