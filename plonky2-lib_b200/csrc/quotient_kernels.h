// Internal interface of quotient_kernels.cu: compute_quotient_polys' per-point evaluation.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/gl_b200.h"
extern std::atomic<unsigned long long> g_gl_launches;

#define QUOTIENT_MAX_GATES 16
#define QUOTIENT_MAX_CHALLENGES 4
#define QUOTIENT_MAX_RATE 32   // 2^quotient_degree_bits

struct quotient_args {
    // column-major LDE leaves of the three oracles (leaf order along the fast axis) and their leading dimensions
    const uint64_t *cs, *wires, *zs;
    uint64_t cs_ld, w_ld, z_ld;
    uint32_t lg_lde;            // log2(n << quotient_degree_bits): points evaluated
    uint32_t qdb;               // quotient_degree_bits
    uint32_t nch, R, deg, chunks, num_prods, num_constants, num_selectors, num_gates, nterms, gate_term0;
    gl_gate gates[QUOTIENT_MAX_GATES];
    const uint64_t* k_is;       // [R]
    const uint64_t* apow;       // [nch][nterms]: alpha_c^t
    const uint64_t* xtab;       // 3 x 1024 power table of w_{lde}
    uint64_t betas[QUOTIENT_MAX_CHALLENGES], gammas[QUOTIENT_MAX_CHALLENGES], pih[4];
    uint64_t zh[QUOTIENT_MAX_RATE], zh_inv[QUOTIENT_MAX_RATE];   // ZeroPolyOnCoset evals / inverses
    uint64_t n_field;           // n mod p
    uint64_t* out;              // [nch][lde_size], natural order
};
unsigned quotient_gate_constraints(const gl_gate& g);
void launch_quotient(const quotient_args& a, cudaStream_t st);
