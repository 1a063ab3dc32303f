// ingest.cuh — parameter block of the ingest kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rlap {

struct IngestParams {
    const long long* src;
    const long long* dst;
    const float* w;      // may be null (unit weights)
    long long e;
    long long n;
    int validate;
    // outputs
    int* ptr;            // [n+1]
    int* col;
    float* wout;
    // workspace
    int* cnt;            // [n]   raw entries per owner
    int* cursor;         // [n]
    int* rawptr;         // [n+1]
    int* cnt2;           // [n]   merged entries per owner
    uint64_t* rkey;      // [e]   (neighbour << 32) | input position
    float* rw;           // [e]
    int* tcol;           // [e]
    float* tw;           // [e]
    unsigned int* dl;    // [n]   owners with more than CAP_WARP raw entries
    int* dl_tail;
    int* status;
    long long* total_dev;
    double* sym_acc;     // [2]
    long long* blocksum;
    uint64_t* scratch;   // NSLOT * 3 * scratch_cap
    int scratch_cap;
};

cudaError_t launch_ingest_stage1(const IngestParams& P, cudaStream_t stream);

}  // namespace rlap
