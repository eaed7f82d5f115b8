// Internal launcher interface of ntt_kernels.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
extern std::atomic<unsigned long long> g_gl_launches;

#define NTT_THREADS 512

struct ntt_pass_args {
    const uint64_t* in;        // column 0 of the input
    uint64_t in_ld;            // elements between columns
    uint64_t in_coset_stride;  // elements between coset blocks of the input (0: every coset reads the same data)
    uint64_t* out;
    uint64_t out_ld;
    uint64_t out_coset_stride;
    const uint64_t* pre_tab;   // [cosets][3][1024] power tables of the coset shifts, or null
    const uint64_t* post_tab;  // [3][1024] power table of w_{2^(s+m)} (direction already applied), or null
    const uint64_t* small_tab; // w_{2^m}^x, x < 2^(m-1)
    uint64_t rows;             // rows per CTA (multiple of 2^m)
    uint64_t final_scale;      // multiply every output (1 = none)
    unsigned m, s, T;
    int canonical_out;
};

void launch_ntt_pass(const ntt_pass_args& a, uint64_t n, uint32_t columns, uint32_t cosets, cudaStream_t st);
void launch_bitrev_permute(const uint64_t* in, uint64_t in_ld, uint64_t* out, uint64_t out_ld, unsigned L,
                           uint32_t columns, uint64_t scale, cudaStream_t st);
void launch_scale_powers(uint64_t* data, uint64_t ld, uint64_t n, uint32_t columns, const uint64_t* tab,
                         cudaStream_t st);
void launch_transpose_to_rows(const uint64_t* cols, uint64_t ld, uint32_t c, uint64_t r0, uint64_t nrows,
                              uint64_t* rows, cudaStream_t st);
void launch_gather_rows(const uint64_t* cols, uint64_t ld, uint32_t c, const uint64_t* idx, uint32_t k,
                        uint64_t* rows, cudaStream_t st);
void launch_gather_paths(const uint64_t* digests, unsigned sub_bits, const uint64_t* idx, uint32_t k,
                         uint64_t* paths, cudaStream_t st);
void launch_fri_leaves(const uint64_t* values_ext, unsigned lg_len, unsigned arity_bits, uint64_t* cols,
                       cudaStream_t st);
void launch_fri_fold(const uint64_t* coeffs_ext, uint64_t out_len, unsigned arity_bits, uint64_t b0, uint64_t b1,
                     uint64_t* out_cols, uint64_t out_ld, cudaStream_t st);
void launch_interleave2(const uint64_t* cols, uint64_t ld, uint64_t n, uint64_t* ext, cudaStream_t st);
void launch_deinterleave2(const uint64_t* ext, uint64_t n, uint64_t* cols, uint64_t ld, cudaStream_t st);

// ---- fast pass: 2^M-point DFTs (M = 8, 9, 10) as two radix-16 register rounds + a radix-2^(M-8) round, one exchange
// through shared memory between rounds (sizes and shifts the TMA path below does not take) ----
struct ntt16_args {
    const uint64_t* in;
    uint64_t in_ld, in_coset_stride;
    uint64_t* out;
    uint64_t out_ld, out_coset_stride;
    const uint64_t* pre_tab;   // [cosets][3][1024] or null
    const uint64_t* post_tab;  // [3][1024] power table of w_{2^(s+M)} (direction applied) or null (s == 0)
    const uint64_t* wtab;      // w_{2^M}^e, e < 2^M (direction applied)
    const uint64_t* pre_direct;   // [cosets][n]: shift_coset^i, replaces pre_tab's two lookups + multiply (or null)
    const uint64_t* post_direct;  // [2^(s+M)]: w_{2^(s+M)}^e, replaces post_tab's lookups + multiply (or null)
    uint64_t n;
    unsigned s, logW;
    int canonical_out;
};
// returns false when (M, s, n) is not handled by the fast kernel
// out[i] = base^i for i < len, from a 3 x 1024 power table (fills the direct twiddle tables once per geometry)
void launch_fill_powers(uint64_t* out, uint64_t len, const uint64_t* powtab, cudaStream_t st);
bool launch_ntt16(const ntt16_args& a, unsigned M, bool inverse, uint64_t n, uint32_t columns, uint32_t cosets,
                  cudaStream_t st);
// gl_add / gl_sub / gl_mul against exact host arithmetic on wrap-around operands; 0 = ok, > 0 = first failing pair + 1,
// < 0 = -cudaError.  Run once per context (gl_ctx_create), next to the Poseidon constants fingerprint.
int gl_field_selftest(cudaStream_t st);
void launch_gather_open(const uint64_t* cols, uint64_t ld, uint32_t c, const uint64_t* digests, unsigned sub_bits,
                        const uint64_t* loc, const uint64_t* slot, uint32_t mine, uint64_t* out, uint64_t per, cudaStream_t st);
void launch_fri_fold_dev(const uint64_t* coeffs_ext, uint64_t out_len, unsigned arity_bits, const uint64_t* beta_dev,
                         uint64_t* out_cols, uint64_t out_ld, cudaStream_t st);
void launch_fri_query_indices(const uint64_t* challenges, uint32_t rounds, unsigned lg_N, const uint32_t* cum_bits, uint32_t layers,
                              uint64_t* idx, uint64_t* proof_x, uint64_t query_stride, cudaStream_t st);
void launch_gather_proof(const uint64_t* cols, uint64_t ld, uint32_t c, const uint64_t* digests, unsigned sub_bits, const uint64_t* idx,
                         uint32_t k, uint64_t* out, uint64_t stride, cudaStream_t st);

// ---- TMA path (ntt_tma.cu): 2^16 .. 2^20 points as a strided 256-point pass + a contiguous 2^(L-8)-point pass ----
struct ntt_tma_args {
    const uint64_t* wt1;      // w_256^e (direction applied), e < 256
    const uint64_t* rowfac;   // [cosets][256]: shift_coset^(row << s), or null (no coset shift)
    const uint64_t* post3;    // [cosets][256][2^s]: (shift_coset * w_{2^L}^brev8(row))^j
    uint64_t* out;            // pass 2 works in place on the output
    uint64_t out_ld, out_coset_stride;
    const uint64_t* wt2;      // w_{2^s}^e (direction applied), e < 2^s
    uint32_t s;               // L - 8
    uint32_t columns, cosets;
    int canonical_out;
};
struct ntt_tma_job {
    const uint64_t* in;
    uint64_t in_ld, in_coset_stride;
    uint64_t* out;
    uint64_t out_ld, out_coset_stride;
    unsigned L;
    uint32_t columns, cosets;
    bool inverse;
    const uint64_t *wt1, *wt2, *rowfac, *post3;
    int canonical_out;
};
bool ntt_tma_supported(unsigned L);
// T = post3, rowfac as above, from the 3 x 1024 power tables of w_{2^L} (post_tab) and of the coset shifts (pre_tabs, may be null)
void launch_ntt_tma_tables(uint64_t* T, uint64_t* rowfac, const uint64_t* post_tab, const uint64_t* pre_tabs, unsigned s,
                           uint32_t cosets, cudaStream_t st);
bool launch_ntt_tma(const ntt_tma_job& j, cudaStream_t st);
