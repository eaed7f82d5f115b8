// Internal launcher interface of fri_kernels.cu (prove_openings before fri_proof: SURVEY 8f N1).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
extern std::atomic<unsigned long long> g_gl_launches;

// comp[i] = sum_j alpha^j * polys[j][i]  (ReducingFactor::reduce_polys_base); polys: k device pointers,
// alpha_pows [k][2]; comp [n][2]
void launch_fri_reduce_polys(const uint64_t* const* polys, uint32_t k, uint64_t n, const uint64_t* alpha_pows,
                             uint64_t* comp_ext, cudaStream_t st);
// final = final * shift + (comp / (X - z) padded with a zero)   (divide_by_linear, shift_poly, +=)
// seg_h, seg_b: scratch [n / DIV_SEG][2] each
#define FRI_DIV_SEG 128
void launch_fri_divide_accumulate(const uint64_t* comp_ext, uint64_t n, const uint64_t z[2], const uint64_t z_seg[2],
                                  const uint64_t shift[2], uint64_t* seg_h, uint64_t* seg_b, uint64_t* final_ext,
                                  cudaStream_t st);
// [n][2] extension coefficients -> two zero-padded base columns cols[0][0..N), cols[1][0..N)
void launch_ext_to_padded_cols(const uint64_t* ext, uint64_t n, uint64_t N, uint64_t* cols, uint64_t* padded_ext,
                               int times_x, cudaStream_t st);
// out[j] = polynomial j of coeffs [c][n] evaluated at the extension point z (OpeningSet::new); partial: scratch
// [c][ceil(n / 4096)][2]; z256 = z^256, zc = z^4096
#define FRI_EVAL_CHUNK 4096
void launch_eval_at(const uint64_t* coeffs, uint64_t n, uint32_t c, const uint64_t z[2], const uint64_t z256[2],
                    const uint64_t zc[2], uint64_t* partial, uint64_t* out, cudaStream_t st);
