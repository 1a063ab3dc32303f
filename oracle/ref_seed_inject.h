// Force-included (-include) in front of the UNMODIFIED reference sources when they
// are compiled into oracle/_ref/ (see oracle/Makefile). TEST INFRASTRUCTURE ONLY.
//
// The reference has no seed API: sampling uses a default-constructed
// std::mt19937_64 (rlap/csrc/preconditioner.cc:356,721,844) and every permutation /
// o_n="random" shuffle uses std::mt19937(std::random_device()()) (:594,304,340,671,706).
// This header renames both types so the test driver can (a) leave behaviour
// untouched (mode 0: real random_device, default seed 5489) or (b) inject
// reproducible values (mode 1) without touching a line of the reference.
#ifndef RLAP_ORACLE_REF_SEED_INJECT_H
#define RLAP_ORACLE_REF_SEED_INJECT_H
#include <random>
#include <algorithm>
#include <cstdint>

extern "C" {
extern uint64_t rlap_ref_sample_seed;   // seed of the mt19937_64 sampling stream (default 5489)
extern int rlap_ref_rd_mode;            // 0: real std::random_device; 1: injected stream
extern uint64_t rlap_ref_rd_state;      // state of the injected stream
}

namespace std {
struct rlap_inj_mt64 : public mt19937_64 {
    rlap_inj_mt64() : mt19937_64(rlap_ref_sample_seed) {}
};
struct rlap_inj_rd {
    typedef unsigned int result_type;
    static constexpr result_type min() { return 0; }
    static constexpr result_type max() { return 0xffffffffu; }
    result_type operator()() {
        if (rlap_ref_rd_mode == 0) { random_device rd; return rd(); }
        // splitmix64 step; the k-th call of a run returns the k-th value for a given start state
        uint64_t z = (rlap_ref_rd_state += 0x9e3779b97f4a7c15ull);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z = z ^ (z >> 31);
        return (result_type)(z >> 32);
    }
};
}  // namespace std
#define mt19937_64 rlap_inj_mt64
#define random_device rlap_inj_rd
#endif
