// C-ABI driver around the UNMODIFIED reference classes (rlap/csrc/factorizers.h:24-48),
// built into oracle/_ref/libref_rlap.so by oracle/Makefile. TEST INFRASTRUCTURE ONLY:
// used by tests/ to pin oracle/rlap_oracle.cc, to generate tests/golden/, and by
// bench.py's cpu_baseline / --impl reference leg. It plays the role of
// rlap/csrc/py_api_binder.cc:54-69 (approximate_cholesky_cpu) without torch.
#include <cstdint>
#include <cstring>
#include <string>
#include <Eigen/Core>
#include "factorizers.h"

extern "C" {
uint64_t rlap_ref_sample_seed = 5489ull;  // std::mt19937_64::default_seed
int rlap_ref_rd_mode = 0;
uint64_t rlap_ref_rd_state = 0;

void ref_set_seeds(uint64_t sample_seed, int rd_mode, uint64_t rd_state) {
    rlap_ref_sample_seed = sample_seed;
    rlap_ref_rd_mode = rd_mode;
    rlap_ref_rd_state = rd_state;
}

// edge_info: row-major [E,3] doubles (row, col, weight) exactly as rlap/ops.py:47 builds it.
// Returns the number of output rows; *out receives a malloc'ed row-major [rows,3] buffer
// (caller frees with ref_free). Mirrors the tensor->Eigen->tensor copies of the binder.
int64_t ref_approximate_cholesky(const double* edge_info, int64_t E, int64_t num_nodes, int64_t num_remove,
                                 const char* o_v, const char* o_n, double** out) {
    Eigen::MatrixXd m(E, 3);
    for (int64_t p = 0; p < E; p++)
        for (int j = 0; j < 3; j++) m(p, j) = edge_info[p * 3 + j];
    ApproximateCholesky ac = ApproximateCholesky();
    ac.setup(m, num_nodes, num_nodes, std::string(o_v), std::string(o_n));
    Eigen::MatrixXd r = ac.getSchurComplement(num_remove);
    int64_t rows = r.rows();
    double* buf = (double*)malloc(sizeof(double) * (size_t)(rows > 0 ? rows : 1) * 3);
    for (int64_t p = 0; p < rows; p++)
        for (int j = 0; j < 3; j++) buf[p * 3 + j] = r(p, j);
    *out = buf;
    return rows;
}

void ref_free(double* p) { free(p); }
}
