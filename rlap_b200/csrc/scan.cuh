// scan.cuh — exclusive prefix sum of int32 counts (three small kernels: block sums, top scan, final).
// out receives n + 1 values (out[n] = total); blocksum needs ceil((n+1)/SCAN_ITEMS) int64 entries.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rlap {

constexpr int SCAN_ITEMS = 4096;  // per block
constexpr int SCAN_THREADS = 512;

template <typename Dummy>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_blocksum(const int* in, long long n, long long* blocksum) {
    __shared__ long long ws[SCAN_THREADS / 32];
    long long base = (long long)blockIdx.x * SCAN_ITEMS;
    long long s = 0;
    for (int i = threadIdx.x; i < SCAN_ITEMS; i += blockDim.x) {
        long long idx = base + i;
        if (idx < n) s += in[idx];
    }
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += ws[k];
        blocksum[blockIdx.x] = t;
    }
}

template <typename Dummy>
__global__ void k_scan_top(long long* blocksum, long long nb, long long* total_out) {
    int lane = threadIdx.x;  // a single warp
    long long carry = 0;
    for (long long base = 0; base < nb; base += 32) {
        long long i = base + lane;
        long long v = i < nb ? blocksum[i] : 0;
        long long s = v;
        for (int d = 1; d < 32; d <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (i < nb) blocksum[i] = carry + s - v;
        carry += __shfl_sync(0xffffffffu, s, 31);
    }
    if (lane == 0 && total_out) *total_out = carry;
}

template <typename OutT>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const int* in, long long n, const long long* blocksum,
                                                             OutT* out) {
    __shared__ long long ws[SCAN_THREADS / 32];
    __shared__ long long carry_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    long long base = (long long)blockIdx.x * SCAN_ITEMS;
    if (threadIdx.x == 0) carry_s = blocksum[blockIdx.x];
    __syncthreads();
    for (int c0 = 0; c0 < SCAN_ITEMS; c0 += blockDim.x) {
        long long idx = base + c0 + threadIdx.x;
        long long v = idx < n ? in[idx] : 0;
        long long s = v;
        for (int d = 1; d < 32; d <<= 1) {
            long long t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) ws[w] = s;
        __syncthreads();
        long long add = 0, tot = 0;
        for (int k = 0; k < nw; k++) { if (k < w) add += ws[k]; tot += ws[k]; }
        long long carry = carry_s;
        if (idx <= n) out[idx] = (OutT)(carry + add + s - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

inline long long scan_blocks(long long n) { return (n + 1 + SCAN_ITEMS - 1) / SCAN_ITEMS; }

template <typename OutT>
inline cudaError_t launch_exclusive_scan(const int* in, long long n, OutT* out, long long* blocksum, long long* total_dev,
                                         cudaStream_t stream) {
    long long nb = scan_blocks(n);
    k_scan_blocksum<int><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, blocksum);
    k_scan_top<int><<<1, 32, 0, stream>>>(blocksum, nb, total_dev);
    k_scan_final<OutT><<<(unsigned)nb, SCAN_THREADS, 0, stream>>>(in, n, blocksum, out);
    return cudaGetLastError();
}

}  // namespace rlap
