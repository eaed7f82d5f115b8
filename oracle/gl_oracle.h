/*
 * gl_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A literal plain-C restatement of the plonky2 v0.1.4 / plonky2_field 0.1.1 algorithms that sit
 * underneath every `data.prove(pw)` / `builder.build::<C>()` / `PoseidonHash::*` call of
 * Orbiter-Finance/Plonky2-lib.  Those algorithms live in the
 * UN-VENDORED dependency `plonky2 0.1.4` (ZeroKPunk fork, Cargo.toml:10-11, Cargo.lock:952-1013),
 * whose source is not under /root/reference; the published upstream algorithm is restated here and
 * parity is anchored on the reference's own call sites and known-answer tests.
 *
 * PARITY STATUS
 *   pinned   : Poseidon permutation / two_to_one / hash_pad-vs-hash_no_pad sponge
 *              (src/zkdsa/circuits/mod.rs:77-106,136-153; src/smt/gadgets/common.rs:28-101;
 *               src/smt/goldilocks_poseidon/mod.rs:158-184).
 *   UNPINNED : NTT/LDE ordering, Merkle digest layout, caps, FRI layer commits, proof-of-work.
 *              "parity unpinned" -- the reference holds no golden vector for them (every circuit
 *              test is prove->verify self-consistency); they follow upstream semantics and are
 *              cross-checked by verifier-style properties only.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (plonky2-lib_b200) never links or calls it.
 */
#ifndef GL_ORACLE_H
#define GL_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define GLO_P 0xFFFFFFFF00000001ULL

/* ---- P0: field ---- */
uint64_t glo_add(uint64_t a, uint64_t b);
uint64_t glo_sub(uint64_t a, uint64_t b);
uint64_t glo_mul(uint64_t a, uint64_t b);
uint64_t glo_pow(uint64_t a, uint64_t e);
uint64_t glo_inv(uint64_t a);
uint64_t glo_primitive_root_of_unity(unsigned lg_n);
void glo_ext_mul(const uint64_t a[2], const uint64_t b[2], uint64_t out[2]);

/* ---- P5: Poseidon ---- */
void glo_poseidon_round_constants(uint64_t out[360]);
void glo_poseidon_permute(uint64_t state[12]);
void glo_poseidon_permute_naive(uint64_t state[12]); /* literal mds_layer; cross-check only */
void glo_poseidon_permute_slow(uint64_t state[12]);  /* literal rounds, MDS on 32-bit halves; cross-check only */
/* the derived fast-partial-round tables of glo_poseidon_permute (upstream's schedule), for inspection */
void glo_poseidon_fast_tables(uint64_t *first12, uint64_t *init121, uint64_t *post11, uint64_t *alpha22,
                              uint64_t *v242, uint64_t *w242);
void glo_hash_no_pad(const uint64_t *in, size_t len, uint64_t out[4]);
void glo_hash_pad(const uint64_t *in, size_t len, uint64_t out[4]);
void glo_hash_or_noop(const uint64_t *in, size_t len, uint64_t out[4]);
void glo_two_to_one(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
/* batch forms (OpenMP) used as CPU baseline */
void glo_permute_batch(uint64_t *states, size_t m);
void glo_two_to_one_batch(const uint64_t *l, const uint64_t *r, uint64_t *out, size_t m);
void glo_hash_no_pad_batch(const uint64_t *in, size_t len_each, size_t m, uint64_t *out);

/* ---- P1/P2/P9: FFT family, all natural order in and out, in place ---- */
void glo_fft(uint64_t *a, unsigned lg_n, unsigned zero_factor);
void glo_ifft(uint64_t *a, unsigned lg_n);
void glo_coset_fft(uint64_t *a, unsigned lg_n, uint64_t shift, unsigned zero_factor);
void glo_coset_ifft(uint64_t *a, unsigned lg_n, uint64_t shift);

/* ---- P3/P4/P11: Merkle tree ---- */
size_t glo_reverse_bits(size_t x, unsigned bits);
/* leaves [num_leaves][leaf_len] row-major; digests [2*(num_leaves-2^h)][4]; cap [2^h][4] */
int glo_merkle_tree(const uint64_t *leaves, size_t num_leaves, size_t leaf_len, unsigned cap_height,
                    uint64_t *digests, uint64_t *cap);
/* siblings [lg(num_leaves)-cap_height][4] */
void glo_merkle_prove(const uint64_t *digests, size_t num_leaves, unsigned cap_height, size_t leaf_index,
                      uint64_t *siblings);
int glo_merkle_verify(const uint64_t *leaf, size_t leaf_len, size_t leaf_index, const uint64_t *siblings,
                      unsigned num_siblings, const uint64_t *cap, unsigned cap_height);

/* ---- P*: PolynomialBatch::from_values / from_coeffs ---- */
/* values/coeffs [c][n] (one contiguous column after another); leaves [N][c]; any output may be NULL */
int glo_commit_from_values(const uint64_t *values, unsigned lg_n, unsigned c, unsigned rate_bits,
                           unsigned cap_height, uint64_t *coeffs_out, uint64_t *leaves_out,
                           uint64_t *digests_out, uint64_t *cap_out);
int glo_commit_from_coeffs(const uint64_t *coeffs, unsigned lg_n, unsigned c, unsigned rate_bits,
                           unsigned cap_height, uint64_t *leaves_out, uint64_t *digests_out,
                           uint64_t *cap_out);

/* ---- P6/P7: SMT node hashing + process-proof verification ---- */
void glo_smt_leaf_hash(const uint64_t key[4], const uint64_t value[4], uint64_t out[4]);
void glo_smt_internal_hash(const uint64_t l[4], const uint64_t r[4], uint64_t out[4]);
typedef struct {
    uint64_t old_root[4], old_key[4], old_value[4];
    uint64_t new_root[4], new_key[4], new_value[4];
    uint64_t siblings[256][4]; /* zero padded beyond num_siblings */
    uint32_t num_siblings;
    uint32_t is_old0;
    uint32_t fnc; /* 0 noop, 1 update, 2 insert, 3 delete */
    uint32_t pad_;
} glo_smt_process_proof;
/* 0 = the reference's asserts all hold; otherwise a code naming the first assert that would fire */
int glo_smt_verify_process_proof(const glo_smt_process_proof *proof);
void glo_smt_verify_process_batch(const glo_smt_process_proof *proofs, size_t m, int32_t *status);

/* in-memory sparse Merkle tree (restates src/smt/tree.rs) -- generates process proofs for tests */
typedef struct glo_smt glo_smt;
glo_smt *glo_smt_new(void);
void glo_smt_free(glo_smt *t);
void glo_smt_root(const glo_smt *t, uint64_t out[4]);
int glo_smt_set(glo_smt *t, const uint64_t key[4], const uint64_t value[4], glo_smt_process_proof *proof);
/* find: returns 1 if found; fills siblings, found key/value */
int glo_smt_find(const glo_smt *t, const uint64_t key[4], uint64_t *siblings, uint32_t *num_siblings,
                 uint64_t not_found_key[4], uint64_t value[4], uint32_t *is_old0);

/* ---- P8/P10: FRI layer commit + proof of work ---- */
/* values: ext elems [len][2] in natural order of the coset shift*<w>, coeffs same length.
 * One step of fri_committed_trees: builds the layer tree over bit-reversed values chunked by arity,
 * then (given beta) folds coeffs and re-evaluates on shift^arity.  See gl_oracle.c. */
int glo_fri_layer_tree(const uint64_t *values_ext, size_t len, unsigned arity_bits, unsigned cap_height,
                       uint64_t *leaves_out, uint64_t *digests_out, uint64_t *cap_out);
void glo_fri_fold(const uint64_t *coeffs_ext, size_t len, unsigned arity_bits, const uint64_t beta[2],
                  uint64_t *folded_coeffs_ext);
void glo_ext_coset_fft(uint64_t *a_ext, unsigned lg_n, uint64_t shift);
/* smallest witness w such that permute(state with state[pos]=w)[out_pos] has >= min_lz leading zeros */
uint64_t glo_pow_grind(const uint64_t state[12], unsigned pos, unsigned out_pos, unsigned min_lz,
                       uint64_t start, uint64_t count);

/* ---- N1: prove_openings before fri_proof ---- */
void glo_reduce_polys_base(const uint64_t *const *polys, size_t k, size_t n, const uint64_t alpha[2], uint64_t *out_ext);
void glo_divide_by_linear(const uint64_t *poly_ext, size_t n, const uint64_t z[2], uint64_t *quot_ext);
void glo_ext_poly_scale_add(uint64_t *acc_ext, size_t n, const uint64_t scalar[2], const uint64_t *add_ext);
void glo_eval_base_poly_at_ext(const uint64_t *coeffs, size_t n, const uint64_t point[2], uint64_t out[2]);

/* ---- N3: permutation argument (Z, partial products) and compute_quotient_polys ---- */
enum { GLO_GATE_NOOP = 0, GLO_GATE_CONSTANT = 1, GLO_GATE_PUBLIC_INPUT = 2, GLO_GATE_U32_INTERLEAVE = 3,
       GLO_GATE_UNINTERLEAVE_TO_U32 = 4, GLO_GATE_UNINTERLEAVE_TO_B32 = 5 };
typedef struct {
    uint32_t kind, num_ops;      /* num_ops: ConstantGate's num_consts / the custom gates' num_ops */
    uint32_t selector_index;     /* column of its selector polynomial among the constants */
    uint32_t group_start, group_end; /* selectors_info.groups[selector_index]; the gate's own index is its position */
    uint32_t reserved;
} glo_gate;
typedef struct {
    uint32_t degree_bits, num_wires, num_routed_wires;
    uint32_t num_constants;      /* constants columns INCLUDING the num_selectors selector columns that come first */
    uint32_t num_selectors, num_challenges, quotient_degree_factor, num_gates;
} glo_circuit;
unsigned glo_num_gate_constraints(const glo_gate *gates, unsigned num_gates);
int glo_permutation_zs(const glo_circuit *cd, const uint64_t *k_is, const uint64_t *wires, const uint64_t *sigmas,
                       const uint64_t *betas, const uint64_t *gammas, uint64_t *out);
int glo_quotient_polys(const glo_circuit *cd, const glo_gate *gates, const uint64_t *k_is, const uint64_t *cs_lde,
                       const uint64_t *wires_lde, const uint64_t *zs_lde, unsigned rate_bits, const uint64_t pih[4],
                       const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, uint64_t *out);
void glo_vanishing_at_point(const glo_circuit *cd, const glo_gate *gates, const uint64_t *k_is, uint64_t x, const uint64_t *lcs,
                            const uint64_t *lw, const uint64_t *lz, const uint64_t *nz, const uint64_t pih[4],
                            const uint64_t *betas, const uint64_t *gammas, const uint64_t *alphas, uint64_t *out);

int glo_num_threads(void);
void glo_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
