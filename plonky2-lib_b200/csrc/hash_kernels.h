// Internal launcher interface of hash_kernels.cu (host side; streams are plain cudaStream_t).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/gl_b200.h"

// kernels launched by this process (gl_ctx_kernel_launches); bumped at every launch site
extern std::atomic<unsigned long long> g_gl_launches;

int gl_poseidon_upload_constants(const uint64_t* rc360);
void launch_permute_batch(uint64_t* states, uint64_t m, cudaStream_t st);
void launch_duplex_chain(uint64_t* state, const uint64_t* chunks, uint64_t m, cudaStream_t st);
void launch_two_to_one_batch(const uint64_t* l, const uint64_t* r, uint64_t* out, uint64_t m, cudaStream_t st);
void launch_hash_no_pad_rows(const uint64_t* in, uint32_t len, uint64_t m, uint64_t* out, cudaStream_t st);
void launch_smt_leaf_hash_batch(const uint64_t* k, const uint64_t* v, uint64_t* out, uint64_t m, cudaStream_t st);
void launch_smt_verify_process(const gl_smt_proof_hdr* p, const uint64_t* sib_pool, const uint64_t* sib_off,
                               uint64_t m, int* status, cudaStream_t st);
void launch_merkle_verify_batch(const uint64_t* leaves, uint32_t leaf_len, const uint64_t* idx, const uint64_t* paths,
                                uint32_t path_len, const uint64_t* cap, uint32_t cap_height, uint64_t k, int* ok, cudaStream_t st);
void launch_pow_grind(const uint64_t* state12, unsigned pos, unsigned out_pos, unsigned min_lz, uint64_t start,
                      uint64_t count, unsigned long long* best, cudaStream_t st);
void launch_leaf_hash_cols(const uint64_t* lde, uint64_t ld, uint32_t c, unsigned lg_leaves, unsigned cap_height,
                           uint64_t* digests, uint64_t* cap, cudaStream_t st);
void launch_leaf_absorb_cols(const uint64_t* lde, uint64_t ld, uint32_t col_begin, uint32_t col_end, unsigned lg_leaves,
                             unsigned cap_height, uint64_t* state, bool first, bool last, uint64_t* digests, uint64_t* cap,
                             cudaStream_t st);
void launch_merkle_levels(unsigned lg_leaves, unsigned cap_height, uint64_t* digests, uint64_t* cap, cudaStream_t st);
void launch_merkle_cols(const uint64_t* lde, uint64_t ld, uint32_t c, unsigned lg_leaves, unsigned cap_height,
                        uint64_t* digests, uint64_t* cap, cudaStream_t st);
void launch_merkle_rows(const uint64_t* leaves, uint32_t leaf_len, unsigned lg_leaves, unsigned cap_height,
                        uint64_t* digests, uint64_t* cap, cudaStream_t st);
// Challenger on the device (one thread): observe n_obs elements from `observe`, then pop n_squeeze challenges into
// `squeeze`; pow_state (or null) receives the 12-lane duplex state fri_proof_of_work grinds on.
void launch_challenger_step(gl_challenger* ch, const uint64_t* observe, uint32_t n_obs, uint64_t* squeeze, uint32_t n_squeeze,
                            uint64_t* pow_state, cudaStream_t st);

// blinding: `count` uniform field elements from (seed, stream) in counter mode
void launch_salt_fill(uint64_t* out, uint64_t count, uint64_t seed, uint64_t stream, cudaStream_t st);
