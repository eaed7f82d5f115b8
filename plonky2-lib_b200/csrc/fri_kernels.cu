// PolynomialBatch::prove_openings up to the FRI polynomial (plonky2::fri::oracle): the
// alpha-combination of every opened polynomial, the division by (X - z) and the accumulation into final_poly.
// All of it is streaming work over the resident coefficients (HBM / L2 bound, a few ms at 2^20 x 171).
#include <cuda_runtime.h>

#include "fri_kernels.h"
#include "gl_field.cuh"

// one thread per coefficient index: coalesced across the k polynomials
__global__ void __launch_bounds__(256)
k_fri_reduce_polys(const u64* const* __restrict__ polys, u32 k, u64 n, const u64* __restrict__ pw, u64* __restrict__ comp) {
    extern __shared__ u64 spw[];   // alpha powers [k][2]
    for (u32 j = threadIdx.x; j < 2 * k; j += blockDim.x) spw[j] = pw[j];
    __syncthreads();
    u64 i = blockIdx.x * (u64)256 + threadIdx.x;
    if (i >= n) return;
    u64 a0 = 0, a1 = 0;
    for (u32 j = 0; j < k; j++) {
        u64 c = __ldg(polys[j] + i);
        a0 = gl_add(a0, gl_mul(c, spw[2 * j]));
        a1 = gl_add(a1, gl_mul(c, spw[2 * j + 1]));
    }
    comp[2 * i] = gl_canon(a0);
    comp[2 * i + 1] = gl_canon(a1);
}

// divide_by_linear is the Horner recurrence b_i = b_{i+1} z + c_i (quotient coefficient q_{i-1} = b_i).  Three steps:
// per segment s the local Horner value H_s = sum_j c_{sT+j} z^j; a serial carry pass B_{s-1} = B_s z^T + H_s over
// the n / T segments; then every segment replays its recurrence from its incoming carry.
__global__ void __launch_bounds__(128)
k_fri_div_local(const u64* __restrict__ comp, u64 nseg, u64 z0, u64 z1, u64* __restrict__ seg_h) {
    u64 s = blockIdx.x * (u64)128 + threadIdx.x;
    if (s >= nseg) return;
    const u64* c = comp + 2 * s * FRI_DIV_SEG;
    gl_ext acc = {0, 0}, z = {z0, z1};
    for (int j = FRI_DIV_SEG - 1; j >= 0; j--) {
        gl_ext t = gl_ext_mul(acc, z);
        acc.a = gl_add(t.a, c[2 * j]);
        acc.b = gl_add(t.b, c[2 * j + 1]);
    }
    seg_h[2 * s] = acc.a;
    seg_h[2 * s + 1] = acc.b;
}
__global__ void k_fri_div_carry(const u64* __restrict__ seg_h, u64 nseg, u64 zt0, u64 zt1, u64* __restrict__ seg_b) {
    if (threadIdx.x | blockIdx.x) return;
    gl_ext acc = {0, 0}, zt = {zt0, zt1};
    for (u64 s = nseg; s-- > 0;) {
        seg_b[2 * s] = acc.a;       // carry entering segment s from above
        seg_b[2 * s + 1] = acc.b;
        gl_ext t = gl_ext_mul(acc, zt);
        acc.a = gl_add(t.a, seg_h[2 * s]);
        acc.b = gl_add(t.b, seg_h[2 * s + 1]);
    }
}
__global__ void __launch_bounds__(128)
k_fri_div_final(const u64* __restrict__ comp, u64 nseg, u64 z0, u64 z1, const u64* __restrict__ seg_b, u64 sh0, u64 sh1,
                u64* __restrict__ fin) {
    u64 s = blockIdx.x * (u64)128 + threadIdx.x;
    if (s >= nseg) return;
    const u64 n = nseg * FRI_DIV_SEG;
    const u64* c = comp + 2 * s * FRI_DIV_SEG;
    gl_ext acc = {seg_b[2 * s], seg_b[2 * s + 1]}, z = {z0, z1}, shift = {sh0, sh1};
    for (int j = FRI_DIV_SEG - 1; j >= 0; j--) {
        const u64 i = s * FRI_DIV_SEG + j;
        if (i == n - 1) {   // quotient padded back to a power of two with a zero
            gl_ext f = {fin[2 * i], fin[2 * i + 1]};
            f = gl_ext_mul(f, shift);
            fin[2 * i] = gl_canon(f.a);
            fin[2 * i + 1] = gl_canon(f.b);
        }
        gl_ext t = gl_ext_mul(acc, z);
        acc.a = gl_add(t.a, c[2 * j]);
        acc.b = gl_add(t.b, c[2 * j + 1]);      // acc = b_i
        if (i >= 1) {                          // q_{i-1} = b_i ; final[i-1] = final[i-1] * shift + q_{i-1}
            gl_ext f = {fin[2 * (i - 1)], fin[2 * (i - 1) + 1]};
            f = gl_ext_mul(f, shift);
            fin[2 * (i - 1)] = gl_canon(gl_add(f.a, acc.a));
            fin[2 * (i - 1) + 1] = gl_canon(gl_add(f.b, acc.b));
        }
    }
}
// small polynomials (n < FRI_DIV_SEG): one thread
__global__ void k_fri_div_small(const u64* __restrict__ comp, u64 n, u64 z0, u64 z1, u64 sh0, u64 sh1, u64* __restrict__ fin) {
    if (threadIdx.x | blockIdx.x) return;
    gl_ext acc = {0, 0}, z = {z0, z1}, shift = {sh0, sh1};
    {
        gl_ext f = {fin[2 * (n - 1)], fin[2 * (n - 1) + 1]};
        f = gl_ext_mul(f, shift);
        fin[2 * (n - 1)] = gl_canon(f.a);
        fin[2 * (n - 1) + 1] = gl_canon(f.b);
    }
    for (u64 i = n; i-- > 0;) {
        gl_ext t = gl_ext_mul(acc, z);
        acc.a = gl_add(t.a, comp[2 * i]);
        acc.b = gl_add(t.b, comp[2 * i + 1]);
        if (i >= 1) {
            gl_ext f = {fin[2 * (i - 1)], fin[2 * (i - 1) + 1]};
            f = gl_ext_mul(f, shift);
            fin[2 * (i - 1)] = gl_canon(gl_add(f.a, acc.a));
            fin[2 * (i - 1) + 1] = gl_canon(gl_add(f.b, acc.b));
        }
    }
}

// times_x: the polynomial is multiplied by X first (coefficients move up one place; the top one is the zero
// prove_openings padded the quotient with) -- the form forks before upstream's "remove the multiplication by X" carry
__global__ void __launch_bounds__(256)
k_ext_to_padded_cols(const u64* __restrict__ ext, u64 n, u64 N, u64* __restrict__ cols, u64* __restrict__ padded, int times_x) {
    u64 i = blockIdx.x * (u64)256 + threadIdx.x;
    if (i >= N) return;
    u64 a = 0, b = 0;
    if (times_x) {
        if (i >= 1 && i < n) { a = ext[2 * (i - 1)]; b = ext[2 * (i - 1) + 1]; }
    } else if (i < n) {
        a = ext[2 * i]; b = ext[2 * i + 1];
    }
    cols[i] = a;
    cols[N + i] = b;
    padded[2 * i] = a;
    padded[2 * i + 1] = b;
}

// OpeningSet::new / PolynomialCoeffs::eval at an extension point for every polynomial of a resident commit:
// out[j] = sum_i coeffs[j][i] * z^i.  A block takes EVAL_CHUNK consecutive coefficients of one polynomial: thread t sums
// c[base + t + 256 k] * (z^256)^k by Horner in k, scales by z^t (binary powering, 8 steps) and by z^base (one thread,
// powering of z^EVAL_CHUNK), and the block adds up; a second pass adds the chunks of each polynomial.
#define EVAL_CHUNK 4096
GL_D gl_ext gl_ext_pow_dev(gl_ext x, u64 e) {
    gl_ext acc = {1, 0};
    while (e) {
        if (e & 1) acc = gl_ext_mul(acc, x);
        x = gl_ext_mul(x, x);
        e >>= 1;
    }
    return acc;
}
__global__ void __launch_bounds__(256)
k_eval_chunks(const u64* __restrict__ coeffs, u64 n, u64 za, u64 zb, u64 z256a, u64 z256b, u64 zca, u64 zcb, u64* __restrict__ partial) {
    __shared__ u64 red[2][256];
    __shared__ u64 zbase[2];
    const u64 chunk = blockIdx.x, poly = blockIdx.y, chunks = gridDim.x;
    const u64 base = chunk * EVAL_CHUNK;
    const unsigned t = threadIdx.x;
    const u64* c = coeffs + poly * n + base;
    if (t == 0) {
        gl_ext zb_ = gl_ext_pow_dev({zca, zcb}, chunk);
        zbase[0] = zb_.a;
        zbase[1] = zb_.b;
    }
    const gl_ext z256 = {z256a, z256b};
    gl_ext acc = {0, 0};
#pragma unroll 1
    for (int k = EVAL_CHUNK / 256 - 1; k >= 0; k--) {
        const u64 i = base + t + 256ull * k;
        acc = gl_ext_mul(acc, z256);
        if (i < n) acc.a = gl_add(acc.a, c[t + 256ull * k]);
    }
    acc = gl_ext_mul(acc, gl_ext_pow_dev({za, zb}, t));
    __syncthreads();
    acc = gl_ext_mul(acc, {zbase[0], zbase[1]});
    red[0][t] = gl_canon(acc.a);
    red[1][t] = gl_canon(acc.b);
    __syncthreads();
    for (unsigned s_ = 128; s_ > 0; s_ >>= 1) {
        if (t < s_) {
            red[0][t] = gl_canon(gl_add(red[0][t], red[0][t + s_]));
            red[1][t] = gl_canon(gl_add(red[1][t], red[1][t + s_]));
        }
        __syncthreads();
    }
    if (t == 0) {
        partial[2 * (poly * chunks + chunk)] = red[0][0];
        partial[2 * (poly * chunks + chunk) + 1] = red[1][0];
    }
}
__global__ void __launch_bounds__(128) k_eval_sum(const u64* __restrict__ partial, u64 chunks, u32 c, u64* __restrict__ out) {
    const u32 poly = blockIdx.x * 128 + threadIdx.x;
    if (poly >= c) return;
    u64 a = 0, b = 0;
    for (u64 k = 0; k < chunks; k++) {
        a = gl_add(a, partial[2 * (poly * chunks + k)]);
        b = gl_add(b, partial[2 * (poly * chunks + k) + 1]);
    }
    out[2 * poly] = gl_canon(a);
    out[2 * poly + 1] = gl_canon(b);
}
// partial: scratch [c][ceil(n / EVAL_CHUNK)][2]; z256 = z^256, zc = z^EVAL_CHUNK (host, exact)
void launch_eval_at(const u64* coeffs, u64 n, u32 c, const u64 z[2], const u64 z256[2], const u64 zc[2], u64* partial, u64* out,
                    cudaStream_t st) {
    const u64 chunks = (n + EVAL_CHUNK - 1) / EVAL_CHUNK;
    k_eval_chunks<<<dim3((unsigned)chunks, c), 256, 0, st>>>(coeffs, n, z[0], z[1], z256[0], z256[1], zc[0], zc[1], partial);
    k_eval_sum<<<(c + 127) / 128, 128, 0, st>>>(partial, chunks, c, out);
    g_gl_launches += 2;
}

void launch_fri_reduce_polys(const u64* const* polys, u32 k, u64 n, const u64* alpha_pows, u64* comp_ext, cudaStream_t st) {
    k_fri_reduce_polys<<<(unsigned)((n + 255) / 256), 256, 2 * k * sizeof(u64), st>>>(polys, k, n, alpha_pows, comp_ext);
    ++g_gl_launches;
}
void launch_fri_divide_accumulate(const u64* comp_ext, u64 n, const u64 z[2], const u64 z_seg[2], const u64 shift[2],
                                  u64* seg_h, u64* seg_b, u64* final_ext, cudaStream_t st) {
    if (n < FRI_DIV_SEG) {
        k_fri_div_small<<<1, 1, 0, st>>>(comp_ext, n, z[0], z[1], shift[0], shift[1], final_ext);
        ++g_gl_launches;
        return;
    }
    u64 nseg = n / FRI_DIV_SEG;
    unsigned blocks = (unsigned)((nseg + 127) / 128);
    k_fri_div_local<<<blocks, 128, 0, st>>>(comp_ext, nseg, z[0], z[1], seg_h);
    k_fri_div_carry<<<1, 32, 0, st>>>(seg_h, nseg, z_seg[0], z_seg[1], seg_b);
    k_fri_div_final<<<blocks, 128, 0, st>>>(comp_ext, nseg, z[0], z[1], seg_b, shift[0], shift[1], final_ext);
    g_gl_launches += 3;
}
void launch_ext_to_padded_cols(const u64* ext, u64 n, u64 N, u64* cols, u64* padded_ext, int times_x, cudaStream_t st) {
    k_ext_to_padded_cols<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(ext, n, N, cols, padded_ext, times_x);
    ++g_gl_launches;
}
