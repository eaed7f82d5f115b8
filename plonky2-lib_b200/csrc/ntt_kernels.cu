// Batched Goldilocks NTT passes (8a rows P1, P2, P3, P9).
//
// A size-2^L transform of one column is done as up to three decimation-in-frequency passes over
// disjoint groups of index bits, each pass = 2^m-point DFTs followed by the "four-step" twiddle.
// Passes of m = 8..10 bits run in k_ntt16 (radix-16 rounds in registers, second half of this file);
// k_ntt_pass (shared-memory radix-2 stages, any m <= 10) takes the small leftover pass.
// Natural order in, bit-reversed order out, in place per index bits:
//
//     pass (m, s):  for every (outer, j_lo):  x[outer][.][j_lo]  <-  DFT_{2^m}(x[outer][.][j_lo])
//                   stored at bit-reversed position, then times w_{2^(s+m)}^(k * j_lo)
//
// A strided pass (s > 0) gives one CTA a tile of T adjacent j_lo so that every global access is a
// T*8-byte segment; the last pass (s = 0) gives one CTA consecutive 2^m-blocks (fully coalesced).
// Bit-reversed output is exactly the leaf order PolynomialBatch needs (leaves[i] = lde[bitrev(i)]),
// so the LDE never runs a separate "transpose + reverse_index_bits" over the 9 GB of leaves.
#include <cuda_runtime.h>

#include "gl_field.cuh"
#include "ntt_kernels.h"
#include "ntt_radix16.cuh"

// Start-up self-test of the carry-chain arithmetic (gl_add / gl_sub / gl_mul / gl_reduce128) on wrap-around operands:
// every pair of the `count` probes, results canonical: out[(i * count + j) * 3 + {0, 1, 2}] = a+b, a-b, a*b.
__global__ void k_field_selftest(const u64* __restrict__ probes, unsigned count, u64* __restrict__ out) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count * count) return;
    u64 a = probes[t / count], b = probes[t % count];
    out[3 * t] = gl_canon(gl_add(a, b));
    out[3 * t + 1] = gl_canon(gl_sub(a, b));
    out[3 * t + 2] = gl_canon(gl_mul(a, b));
}
int gl_field_selftest(cudaStream_t st) {
    static const u64 probes[] = {0, 1, 2, GL_P - 1, GL_P, GL_P + 1, 0xFFFFFFFFULL, 0x100000000ULL, 0xFFFFFFFFFFFFFFFFULL,
                                 0xFFFFFFFF00000000ULL, 0xFFFFFFFEFFFFFFFFULL, 0x8000000000000000ULL, 0x7FFFFFFF80000001ULL,
                                 0x0123456789ABCDEFULL, 0xFEDCBA9876543210ULL, 0xFFFFFFFF7FFFFFFFULL};
    const unsigned count = sizeof(probes) / sizeof(probes[0]);
    u64 *d_in = nullptr, *d_out = nullptr;
    static u64 got[3 * 16 * 16];
    cudaError_t e = cudaMalloc(&d_in, sizeof(probes));
    if (e == cudaSuccess) e = cudaMalloc(&d_out, sizeof(got));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_in, probes, sizeof(probes), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_field_selftest<<<(count * count + 127) / 128, 128, 0, st>>>(d_in, count, d_out);
        ++g_gl_launches;
        e = cudaMemcpyAsync(got, d_out, sizeof(got), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d_in);
    cudaFree(d_out);
    if (e != cudaSuccess) return -(int)e;
    for (unsigned i = 0; i < count; i++)
        for (unsigned j = 0; j < count; j++) {
            const u64* g = got + 3 * (i * count + j);
            if (g[0] != glh::add(probes[i], probes[j]) || g[1] != glh::sub(probes[i], probes[j]) ||
                g[2] != glh::mul(probes[i], probes[j]))
                return 1 + (int)(i * count + j);
        }
    return 0;
}

// base^e from a 3 x 1024 table (e < 2^30)
GL_D u64 powtab_eval(const u64* __restrict__ tab, u64 e) {
    u64 r = __ldg(tab + (e & 1023));
    if (e >> 10) {
        r = gl_mul(r, __ldg(tab + 1024 + ((e >> 10) & 1023)));
        if (e >> 20) r = gl_mul(r, __ldg(tab + 2048 + ((e >> 20) & 1023)));
    }
    return r;
}

// One pass.  grid = (tiles per column, columns, cosets).  Dynamic smem: rows * T elements + 2^(m-1) twiddles.
__global__ void __launch_bounds__(NTT_THREADS)
k_ntt_pass(ntt_pass_args a) {
    extern __shared__ u64 sm[];
    const unsigned m = a.m, s = a.s, T = a.T;
    const u64 rows = a.rows;                   // rows per CTA (multiple of 2^m)
    const u64 E = rows * T;
    u64* tw = sm + E;                          // w_{2^m}^x, x < 2^(m-1)
    const unsigned coset = blockIdx.z;
    const u64 col = blockIdx.y;
    // position of this CTA's tile inside the column
    u64 tile = blockIdx.x, base;
    if (s == 0) {
        base = tile * rows;                    // consecutive 2^m blocks
    } else {
        u64 tiles_per_block = ((u64)1 << s) / T;        // tiles along j_lo
        u64 outer = tile / tiles_per_block, jt = tile % tiles_per_block;
        base = (outer << (s + m)) + jt * T;
    }
    const u64 rowstride = (u64)1 << s;
    const u64* in = a.in + col * a.in_ld + (u64)coset * a.in_coset_stride;
    u64* out = a.out + col * a.out_ld + (u64)coset * a.out_coset_stride;
    const u64* pre = a.pre_tab ? a.pre_tab + (u64)coset * 3072 : nullptr;

    for (u64 x = threadIdx.x; x < ((u64)1 << m) / 2; x += NTT_THREADS) tw[x] = __ldg(a.small_tab + x);
    for (u64 e = threadIdx.x; e < E; e += NTT_THREADS) {
        u64 row = e / T, t = e % T;
        u64 pos = base + row * rowstride + t;
        u64 v = in[pos];
        if (pre) v = gl_mul(v, powtab_eval(pre, pos));   // coset_fft: coeffs[i] * shift^i
        sm[e] = v;
    }
    __syncthreads();
    // DIF stages over the low m bits of the row index
    for (unsigned st = 0; st < m; st++) {
        const unsigned lh = m - 1 - st;                  // log2(half)
        const u64 half = (u64)1 << lh;
        for (u64 b = threadIdx.x; b < E / 2; b += NTT_THREADS) {
            u64 t = b % T, bb = b / T;
            u64 j = bb & (half - 1), blk = bb >> lh;
            u64 i0 = ((blk << (lh + 1)) + j) * T + t, i1 = i0 + half * T;
            u64 u = sm[i0], v = sm[i1];
            sm[i0] = gl_add(u, v);
            u64 d = gl_sub(u, v);
            sm[i1] = st == m - 1 ? d : gl_mul(d, tw[j << st]);
        }
        __syncthreads();
    }
    // store (in-place positions) with the four-step twiddle w_{2^(s+m)}^(k * j_lo), k = bitrev_m(row)
    for (u64 e = threadIdx.x; e < E; e += NTT_THREADS) {
        u64 row = e / T, t = e % T;
        u64 pos = base + row * rowstride + t;
        u64 v = sm[e];
        if (s != 0 && a.post_tab) {
            u64 k = m ? (__brevll(row & (((u64)1 << m) - 1)) >> (64 - m)) : 0;
            u64 jlo = pos & (rowstride - 1);
            v = gl_mul(v, powtab_eval(a.post_tab, k * jlo));
        }
        if (a.final_scale != 1) v = gl_mul(v, a.final_scale);
        if (a.canonical_out) v = gl_canon(v);
        out[pos] = v;
    }
}

void launch_ntt_pass(const ntt_pass_args& a, u64 n, u32 columns, u32 cosets, cudaStream_t st) {
    u64 E = a.rows * a.T;
    size_t smem = (E + (((u64)1 << a.m) / 2)) * sizeof(u64);
    // per launch (cheap) rather than once per process: the attribute is per device
    if (smem > 48 * 1024) cudaFuncSetAttribute(k_ntt_pass, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    dim3 grid((unsigned)(n / E), columns, cosets);
    { k_ntt_pass<<<grid, NTT_THREADS, smem, st>>>(a); ++g_gl_launches; }
}

// out[col][bitrev_L(i)] = in[col][i] * scale  (canonical).  Tiled through shared memory so both sides
// move 32-element (256 B) runs: i = (hi | mid | lo) with |hi| = |lo| = 5 bits swaps hi and lo blocks.
__global__ void __launch_bounds__(256)
k_bitrev_permute(const u64* __restrict__ in, u64 in_ld, u64* __restrict__ out, u64 out_ld, unsigned L, u64 scale) {
    const u64 col = blockIdx.y;
    const u64* src = in + col * in_ld;
    u64* dst = out + col * out_ld;
    if (L < 10) {  // small: direct
        u64 n = (u64)1 << L;
        for (u64 i = blockIdx.x * (u64)256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
            u64 j = L ? (__brevll(i) >> (64 - L)) : 0;
            u64 v = src[i];
            if (scale != 1) v = gl_mul(v, scale);
            dst[j] = gl_canon(v);
        }
        return;
    }
    __shared__ u64 tile[32][33];
    const unsigned midbits = L - 10;
    for (u64 mid = blockIdx.x; mid < ((u64)1 << midbits); mid += gridDim.x) {
        u64 rmid = midbits ? (__brevll(mid) >> (64 - midbits)) : 0;
        unsigned lo = threadIdx.x & 31, hi0 = threadIdx.x >> 5;
        for (unsigned hi = hi0; hi < 32; hi += 8) {
            u64 v = src[((u64)hi << (L - 5)) | (mid << 5) | lo];
            if (scale != 1) v = gl_mul(v, scale);
            tile[hi][lo] = gl_canon(v);
        }
        __syncthreads();
        // destination index: (rev(lo) | rev(mid) | rev(hi)); let thread's fast index run over rev(hi)
        unsigned a = threadIdx.x & 31, b0 = threadIdx.x >> 5;   // a = rev5(hi) position, b = rev5(lo)
        for (unsigned b = b0; b < 32; b += 8) {
            unsigned hi = __brev(a) >> 27, lo2 = __brev(b) >> 27;
            dst[((u64)b << (L - 5)) | (rmid << 5) | a] = tile[hi][lo2];
        }
        __syncthreads();
    }
}

void launch_bitrev_permute(const u64* in, u64 in_ld, u64* out, u64 out_ld, unsigned L, u32 columns, u64 scale,
                           cudaStream_t st) {
    u64 blocks = L < 10 ? 1 : ((u64)1 << (L - 10));
    if (blocks > 4096) blocks = 4096;
    dim3 grid((unsigned)blocks, columns);
    { k_bitrev_permute<<<grid, 256, 0, st>>>(in, in_ld, out, out_ld, L, scale); ++g_gl_launches; }
}

// ifft_with_options epilogue done directly: plonky2 runs the forward FFT and then maps
// out[i] <-> out[n-i] * n^-1; that is the inverse DFT, which the passes above compute with inverse roots.

// Elementwise helpers ------------------------------------------------------------------------------
// data[col][i] *= base^i (coset scaling for natural-order coset_fft / coset_ifft)
__global__ void __launch_bounds__(256)
k_scale_powers(u64* __restrict__ data, u64 ld, u64 n, const u64* __restrict__ tab) {
    u64 col = blockIdx.y;
    for (u64 i = blockIdx.x * (u64)256 + threadIdx.x; i < n; i += (u64)gridDim.x * 256) {
        u64* p = data + col * ld + i;
        *p = gl_canon(gl_mul(*p, powtab_eval(tab, i)));
    }
}
void launch_scale_powers(u64* data, u64 ld, u64 n, u32 columns, const u64* tab, cudaStream_t st) {
    u64 blocks = (n + 255) / 256;
    if (blocks > 2048) blocks = 2048;
    dim3 grid((unsigned)blocks, columns);
    { k_scale_powers<<<grid, 256, 0, st>>>(data, ld, n, tab); ++g_gl_launches; }
}

// Column-major [c][ld] (rows r0 .. r0+nrows) -> row-major [nrows][c]  ("transpose LDEs" for mirror mode)
__global__ void __launch_bounds__(256)
k_transpose_to_rows(const u64* __restrict__ cols, u64 ld, u32 c, u64 r0, u64 nrows, u64* __restrict__ rows) {
    __shared__ u64 tile[32][33];
    u64 rb = (u64)blockIdx.x * 32, cb = (u64)blockIdx.y * 32;
    unsigned tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (unsigned j = ty; j < 32; j += 8) {
        u64 col = cb + j, row = rb + tx;
        tile[j][tx] = (col < c && row < nrows) ? cols[col * ld + r0 + row] : 0;
    }
    __syncthreads();
    for (unsigned j = ty; j < 32; j += 8) {
        u64 row = rb + j, col = cb + tx;
        if (row < nrows && col < c) rows[row * c + col] = tile[tx][j];
    }
}
void launch_transpose_to_rows(const u64* cols, u64 ld, u32 c, u64 r0, u64 nrows, u64* rows, cudaStream_t st) {
    dim3 grid((unsigned)((nrows + 31) / 32), (c + 31) / 32);
    { k_transpose_to_rows<<<grid, 256, 0, st>>>(cols, ld, c, r0, nrows, rows); ++g_gl_launches; }
}

// Gather k rows (leaf indices local to the buffer) into [k][c]
__global__ void k_gather_rows(const u64* __restrict__ cols, u64 ld, u32 c, const u64* __restrict__ idx, u32 k,
                              u64* __restrict__ rows) {
    u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (g >= (u64)k * c) return;
    u64 q = g / c, col = g % c;
    rows[g] = cols[col * ld + idx[q]];
}
void launch_gather_rows(const u64* cols, u64 ld, u32 c, const u64* idx, u32 k, u64* rows, cudaStream_t st) {
    u64 total = (u64)k * c;
    if (total) { k_gather_rows<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cols, ld, c, idx, k, rows); ++g_gl_launches; }
}

// MerkleTree::prove for k local leaf indices: paths [k][sub_bits][4]
__global__ void k_gather_paths(const u64* __restrict__ digests, unsigned sub_bits, const u64* __restrict__ idx, u32 k,
                               u64* __restrict__ paths) {
    u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (g >= (u64)k * sub_bits) return;
    u64 q = g / sub_bits;
    unsigned layer = (unsigned)(g % sub_bits);
    u64 leaf = idx[q];
    u64 subtree = leaf >> sub_bits;
    u64 pair = (leaf & (((u64)1 << sub_bits) - 1)) >> layer;   // node index at `layer`
    u64 parity = pair & 1;
    pair >>= 1;
    u64 sib = 2 * ((pair << (layer + 1)) + ((u64)1 << layer) - 1) + (1 - parity);
    u64 per_subtree = 2 * (((u64)1 << sub_bits) - 1);
    const u64* src = digests + 4 * (subtree * per_subtree + sib);
    u64* dst = paths + 4 * g;
    dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
}
void launch_gather_paths(const u64* digests, unsigned sub_bits, const u64* idx, u32 k, u64* paths, cudaStream_t st) {
    u64 total = (u64)k * sub_bits;
    if (total) { k_gather_paths<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(digests, sub_bits, idx, k, paths); ++g_gl_launches; }
}

// gl_group_commit_open: rows and paths of the `mine` leaves this shard owns, written at their slots of the exchange
// buffer: out[slot * per + col] = row, out[slot * per + c + 4 * layer + w] = sibling digest of `layer`.
__global__ void k_gather_open(const u64* __restrict__ cols, u64 ld, u32 c, const u64* __restrict__ digests, unsigned sub_bits,
                              const u64* __restrict__ loc, const u64* __restrict__ slot, u32 mine, u64* __restrict__ out,
                              u64 per) {
    u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (g >= (u64)mine * per) return;
    u64 q = g / per, e = g % per;
    u64 leaf = loc[q];
    u64* dst = out + (slot ? slot[q] : q) * per + e;
    if (e < c) {
        *dst = cols[e * ld + leaf];
        return;
    }
    unsigned layer = (unsigned)((e - c) >> 2), w = (unsigned)((e - c) & 3);
    u64 subtree = leaf >> sub_bits;
    u64 pair = (leaf & (((u64)1 << sub_bits) - 1)) >> layer;
    u64 parity = pair & 1;
    pair >>= 1;
    u64 sib = 2 * ((pair << (layer + 1)) + ((u64)1 << layer) - 1) + (1 - parity);
    u64 per_subtree = 2 * (((u64)1 << sub_bits) - 1);
    *dst = digests[4 * (subtree * per_subtree + sib) + w];
}
void launch_gather_open(const u64* cols, u64 ld, u32 c, const u64* digests, unsigned sub_bits, const u64* loc, const u64* slot,
                        u32 mine, u64* out, u64 per, cudaStream_t st) {
    u64 total = (u64)mine * per;
    if (total) { k_gather_open<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cols, ld, c, digests, sub_bits, loc, slot, mine, out, per); ++g_gl_launches; }
}

// gl_fri_prove: (row, path) of leaf idx[q] for every query round q, written into the proof buffer at out[q * stride + e],
// e < c + 4 * sub_bits (row first, then the siblings from the leaf level up)
__global__ void k_gather_proof(const u64* __restrict__ cols, u64 ld, u32 c, const u64* __restrict__ digests, unsigned sub_bits,
                               const u64* __restrict__ idx, u32 k, u64* __restrict__ out, u64 stride) {
    const u64 width = c + 4 * (u64)sub_bits;
    u64 g = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (g >= (u64)k * width) return;
    u64 q = g / width, e = g % width;
    u64 leaf = idx[q];
    u64* dst = out + q * stride + e;
    if (e < c) {
        *dst = cols[e * ld + leaf];
        return;
    }
    unsigned layer = (unsigned)((e - c) >> 2), w = (unsigned)((e - c) & 3);
    u64 subtree = leaf >> sub_bits;
    u64 pair = (leaf & (((u64)1 << sub_bits) - 1)) >> layer;
    u64 parity = pair & 1;
    pair >>= 1;
    u64 sib = 2 * ((pair << (layer + 1)) + ((u64)1 << layer) - 1) + (1 - parity);
    u64 per_subtree = 2 * (((u64)1 << sub_bits) - 1);
    *dst = digests[4 * (subtree * per_subtree + sib) + w];
}
void launch_gather_proof(const u64* cols, u64 ld, u32 c, const u64* digests, unsigned sub_bits, const u64* idx, u32 k, u64* out,
                         u64 stride, cudaStream_t st) {
    u64 total = (u64)k * (c + 4 * (u64)sub_bits);
    if (total) { k_gather_proof<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(cols, ld, c, digests, sub_bits, idx, k, out, stride); ++g_gl_launches; }
}

// FRI layer leaves, column-major: element (leaf j, column 2a+e) = values_ext[bitrev(j * arity + a)][e]
__global__ void __launch_bounds__(256)
k_fri_leaves(const u64* __restrict__ values_ext, unsigned lg_len, unsigned arity_bits, u64* __restrict__ cols) {
    u64 len = (u64)1 << lg_len, nl = len >> arity_bits;
    u64 g = blockIdx.x * (u64)256 + threadIdx.x;
    if (g >= len) return;
    u64 a = g / nl, j = g % nl;           // consecutive threads -> consecutive leaves of one column pair
    u64 src = lg_len ? (__brevll((j << arity_bits) + a) >> (64 - lg_len)) : 0;
    cols[(2 * a) * nl + j] = gl_canon(values_ext[2 * src]);
    cols[(2 * a + 1) * nl + j] = gl_canon(values_ext[2 * src + 1]);
}
void launch_fri_leaves(const u64* values_ext, unsigned lg_len, unsigned arity_bits, u64* cols, cudaStream_t st) {
    u64 len = (u64)1 << lg_len;
    { k_fri_leaves<<<(unsigned)((len + 255) / 256), 256, 0, st>>>(values_ext, lg_len, arity_bits, cols); ++g_gl_launches; }
}

// reduce_with_powers over chunks of 2^arity_bits extension coefficients; output split into two base
// columns out[0][k], out[1][k] (ready for the two base-field NTTs of the extension coset_fft)
__global__ void __launch_bounds__(256)
k_fri_fold(const u64* __restrict__ coeffs_ext, u64 out_len, unsigned arity_bits, u64 b0, u64 b1,
           u64* __restrict__ out_cols, u64 out_ld) {
    u64 k = blockIdx.x * (u64)256 + threadIdx.x;
    if (k >= out_len) return;
    unsigned arity = 1u << arity_bits;
    gl_ext beta = {b0, b1}, acc = {0, 0};
    const u64* p = coeffs_ext + 2 * (k << arity_bits);
    for (int j = (int)arity - 1; j >= 0; j--) {
        gl_ext cj = {p[2 * j], p[2 * j + 1]};
        acc = gl_ext_add(gl_ext_mul(acc, beta), cj);
    }
    out_cols[k] = gl_canon(acc.a);
    out_cols[out_ld + k] = gl_canon(acc.b);
}
// the same fold with beta read from device memory (gl_fri_prove: the challenge never visits the host)
__global__ void __launch_bounds__(256)
k_fri_fold_dev(const u64* __restrict__ coeffs_ext, u64 out_len, unsigned arity_bits, const u64* __restrict__ beta_dev,
               u64* __restrict__ out_cols, u64 out_ld) {
    u64 k = blockIdx.x * (u64)256 + threadIdx.x;
    if (k >= out_len) return;
    unsigned arity = 1u << arity_bits;
    gl_ext beta = {__ldg(beta_dev), __ldg(beta_dev + 1)}, acc = {0, 0};
    const u64* p = coeffs_ext + 2 * (k << arity_bits);
    for (int j = (int)arity - 1; j >= 0; j--) {
        gl_ext cj = {p[2 * j], p[2 * j + 1]};
        acc = gl_ext_add(gl_ext_mul(acc, beta), cj);
    }
    out_cols[k] = gl_canon(acc.a);
    out_cols[out_ld + k] = gl_canon(acc.b);
}
void launch_fri_fold_dev(const u64* coeffs_ext, u64 out_len, unsigned arity_bits, const u64* beta_dev, u64* out_cols, u64 out_ld,
                         cudaStream_t st) {
    { k_fri_fold_dev<<<(unsigned)((out_len + 255) / 256), 256, 0, st>>>(coeffs_ext, out_len, arity_bits, beta_dev, out_cols, out_ld); ++g_gl_launches; }
}
// fri_prover_query_rounds: x_index = challenge % N for every round; idx[0][q] = x_index, idx[l + 1][q] = x_index >> (arity bits
// folded up to and including layer l); the x_index also goes to its slot of the proof buffer
__global__ void k_fri_query_indices(const u64* __restrict__ challenges, u32 rounds, unsigned lg_N, const u32* __restrict__ cum_bits,
                                    u32 layers, u64* __restrict__ idx, u64* __restrict__ proof_x, u64 query_stride) {
    u32 q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= rounds) return;
    u64 x = challenges[q] & (((u64)1 << lg_N) - 1);     // canonical challenge mod N, N a power of two
    idx[q] = x;
    proof_x[(u64)q * query_stride] = x;
    for (u32 l = 0; l < layers; l++) idx[(u64)(l + 1) * rounds + q] = x >> cum_bits[l];
}
void launch_fri_query_indices(const u64* challenges, u32 rounds, unsigned lg_N, const u32* cum_bits, u32 layers, u64* idx,
                              u64* proof_x, u64 query_stride, cudaStream_t st) {
    k_fri_query_indices<<<(rounds + 63) / 64, 64, 0, st>>>(challenges, rounds, lg_N, cum_bits, layers, idx, proof_x, query_stride);
    ++g_gl_launches;
}

void launch_fri_fold(const u64* coeffs_ext, u64 out_len, unsigned arity_bits, u64 b0, u64 b1, u64* out_cols,
                     u64 out_ld, cudaStream_t st) {
    { k_fri_fold<<<(unsigned)((out_len + 255) / 256), 256, 0, st>>>(coeffs_ext, out_len, arity_bits, b0, b1, out_cols, out_ld); ++g_gl_launches; }
}

// [2][n] base columns <-> [n][2] interleaved extension elements
__global__ void k_interleave2(const u64* __restrict__ cols, u64 ld, u64 n, u64* __restrict__ ext) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    ext[2 * i] = cols[i];
    ext[2 * i + 1] = cols[ld + i];
}
__global__ void k_deinterleave2(const u64* __restrict__ ext, u64 n, u64* __restrict__ cols, u64 ld) {
    u64 i = blockIdx.x * (u64)blockDim.x + threadIdx.x;
    if (i >= n) return;
    cols[i] = ext[2 * i];
    cols[ld + i] = ext[2 * i + 1];
}
void launch_interleave2(const u64* cols, u64 ld, u64 n, u64* ext, cudaStream_t st) {
    if (n) { k_interleave2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cols, ld, n, ext); ++g_gl_launches; }
}
void launch_deinterleave2(const u64* ext, u64 n, u64* cols, u64 ld, cudaStream_t st) {
    if (n) { k_deinterleave2<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ext, n, cols, ld); ++g_gl_launches; }
}

// ================================================================================================
// Fast pass.  One CTA transforms W independent 2^M-point DFTs (W adjacent j_lo for a strided pass, W
// consecutive blocks for the last pass), 16 elements per thread:
//   round 1: radix-16 over the top 4 index bits, in registers (loaded straight from HBM), then the
//            twiddle w_{2^M}^(k1 * low) from a shared-memory table;
//   round 2: after one exchange through shared memory, radix-16 over the next 4 bits, twiddle
//            w_{2^(M-4)}^(k2 * low);
//   round 3: the last M-8 <= 2 bits after a second exchange through shared memory (its own swizzle), as 2^(4-LB)
//            independent radix-2^LB blocks per thread; results go straight to HBM with
//            the four-step twiddle w_{2^(s+M)}^(k * j_lo).
// 2^192 = 1 (mod p): every root of unity of order <= 64 is a power of two (plonky2's w_16 is 2^156),
// so the 17 twiddles inside a radix-16 block are compile-time constants with one or two set bits per
// 32-bit half and their products compile to shifts.
// ================================================================================================
template <int M, bool INV, bool STRIDED>
__global__ void __launch_bounds__(512, 2) k_ntt16(ntt16_args a) {
    extern __shared__ u64 sm[];
    constexpr unsigned Q = 1u << (M - 4);          // threads per DFT
    const unsigned logW = a.logW, W = 1u << logW;
    u64* wt = sm + ((size_t)W << M);
    const unsigned tid = threadIdx.x;
    const unsigned d = STRIDED ? (tid & (W - 1)) : (tid >> (M - 4));
    const unsigned q = STRIDED ? (tid >> logW) : (tid & (Q - 1));
    const unsigned s = a.s;
    const unsigned coset = blockIdx.z;
    const u64 col = blockIdx.y;
    u64 base;
    if (STRIDED) {
        u64 tiles_per_block = ((u64)1 << s) >> logW;
        u64 outer = blockIdx.x / tiles_per_block, jt = blockIdx.x % tiles_per_block;
        base = (outer << (s + M)) + (jt << logW) + d;
    } else {
        base = ((u64)blockIdx.x << (M + logW)) + ((u64)d << M);
    }
    const u64* in = a.in + col * a.in_ld + (u64)coset * a.in_coset_stride + base;
    u64* out = a.out + col * a.out_ld + (u64)coset * a.out_coset_stride + base;
    const unsigned rs = STRIDED ? s : 0;          // log2 of the row stride in HBM

    for (unsigned e = tid; e < (1u << M); e += blockDim.x) wt[e] = __ldg(a.wtab + e);

    u64 x[16];
    // ---- round 1: rows a * Q + q -----------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = in[(u64)(i * Q + q) << rs];
    if (a.pre_direct) {
        const u64* pre = a.pre_direct + (u64)coset * a.n + base;
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = gl_mul(x[i], __ldg(pre + ((u64)(i * Q + q) << rs)));
    } else if (a.pre_tab) {
        const u64* pre = a.pre_tab + (u64)coset * 3072;
#pragma unroll
        for (int i = 0; i < 16; i++) x[i] = gl_mul(x[i], powtab_eval(pre, base + ((u64)(i * Q + q) << rs)));
    }
    radix16_dif<INV>(x);
    __syncthreads();   // twiddle table ready
#pragma unroll
    for (int i = 1; i < 16; i++) x[i] = gl_mul(x[i], wt[brev4(i) * q]);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned p = i * Q + q;
        unsigned idx = STRIDED ? ((p << logW) + d) : ((d << M) + (p ^ (((p >> (M - 4)) & 15u) << (M - 8))));
        sm[idx] = x[i];
    }
    __syncthreads();
    // ---- round 2: rows phigh * Q + r * 2^(M-8) + plow ------------------------------------------------
    constexpr unsigned LB = M - 8;                 // bits left for round 3
    const unsigned plow = q & ((1u << LB) - 1), phigh = q >> LB;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned p = phigh * Q + ((unsigned)i << LB) + plow;
        unsigned idx = STRIDED ? ((p << logW) + d) : ((d << M) + (p ^ (((p >> (M - 4)) & 15u) << (M - 8))));
        x[i] = sm[idx];
    }
    radix16_dif<INV>(x);
    if (LB == 0) {
        // ---- store: in-place position p holds output index brev_M(p) -----------------------------------
        const u64 jlo = STRIDED ? (base & (((u64)1 << s) - 1)) : 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            unsigned p = phigh * Q + (unsigned)i;
            u64 v = x[i];
            if (STRIDED && a.post_direct) {
                u64 k = __brev(p) >> (32 - M);
                v = gl_mul(v, __ldg(a.post_direct + k * jlo));
            } else if (STRIDED && a.post_tab) {
                u64 k = __brev(p) >> (32 - M);
                v = gl_mul(v, powtab_eval(a.post_tab, k * jlo));
            }
            if (a.canonical_out) v = gl_canon(v);
            out[(u64)p << rs] = v;
        }
        return;
    }
#pragma unroll
    for (int i = 1; i < 16; i++) x[i] = gl_mul(x[i], wt[(brev4(i) * plow) << 4]);
    // ---- round 3: the last LB <= 2 bits.  Second exchange (its own swizzle), then each thread holds the LB-bit
    // field f = p[LB-1..0] for 2^(4-LB) values g of the TOP bits of p, i.e. 2^(4-LB) independent radix-2^LB blocks.
    constexpr unsigned GB = 4 - LB;                      // bits of g
    __syncthreads();                                     // everyone is done reading the first exchange
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned p = phigh * Q + ((unsigned)i << LB) + plow;
        unsigned idx = STRIDED ? (((p << logW) + d) ^ (LB == 2 ? (((p >> 2) & 3u) << logW) : 0u))
                               : ((d << M) + (p ^ ((p >> 4) & 15u)));
        sm[idx] = x[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned g = (unsigned)i >> LB, f = (unsigned)i & ((1u << LB) - 1);
        unsigned p = (g << (M - GB)) | (q << LB) | f;
        unsigned idx = STRIDED ? (((p << logW) + d) ^ (LB == 2 ? (((p >> 2) & 3u) << logW) : 0u))
                               : ((d << M) + (p ^ ((p >> 4) & 15u)));
        x[i] = sm[idx];
    }
    if (LB == 2) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
            u64 a0 = x[4 * g], a1 = x[4 * g + 1], a2 = x[4 * g + 2], a3 = x[4 * g + 3];
            u64 s02 = gl_add(a0, a2), d02 = gl_sub(a0, a2);
            u64 s13 = gl_add(a1, a3);
            // (a1 - a3) * w_4, w_4 = 2^48 (inverse: 2^144 = -2^48)
            u64 d13 = gl_mul(INV ? gl_sub(a3, a1) : gl_sub(a1, a3), pow2_mod_p(48));
            x[4 * g] = gl_add(s02, s13);
            x[4 * g + 1] = gl_sub(s02, s13);
            x[4 * g + 2] = gl_add(d02, d13);
            x[4 * g + 3] = gl_sub(d02, d13);
        }
    } else {
#pragma unroll
        for (int g = 0; g < 8; g++) {
            u64 a0 = x[2 * g], a1 = x[2 * g + 1];
            x[2 * g] = gl_add(a0, a1);
            x[2 * g + 1] = gl_sub(a0, a1);
        }
    }
    // ---- store: in-place position p holds output index brev_M(p) ---------------------------------------
    const u64 jlo = STRIDED ? (base & (((u64)1 << s) - 1)) : 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        unsigned g = (unsigned)i >> LB, f = (unsigned)i & ((1u << LB) - 1);
        unsigned p = (g << (M - GB)) | (q << LB) | f;
        u64 v = x[i];
        if (STRIDED && a.post_direct) {
            u64 k = __brev(p) >> (32 - M);
            v = gl_mul(v, __ldg(a.post_direct + k * jlo));
        } else if (STRIDED && a.post_tab) {
            u64 k = __brev(p) >> (32 - M);
            v = gl_mul(v, powtab_eval(a.post_tab, k * jlo));
        }
        if (a.canonical_out) v = gl_canon(v);
        x[i] = v;
        if (STRIDED) out[(u64)p << rs] = v;
    }
    if (!STRIDED) {
        // f is the low bits of p: a thread owns 2^LB consecutive outputs per g -> 16-byte stores
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            unsigned g = (unsigned)i >> LB, f = (unsigned)i & ((1u << LB) - 1);
            unsigned p = (g << (M - GB)) | (q << LB) | f;
            *reinterpret_cast<ulonglong2*>(out + p) = make_ulonglong2(x[i], x[i + 1]);
        }
    }
}

__global__ void __launch_bounds__(256) k_fill_powers(u64* __restrict__ out, u64 len, const u64* __restrict__ tab) {
    for (u64 i = blockIdx.x * (u64)256 + threadIdx.x; i < len; i += (u64)gridDim.x * 256)
        out[i] = gl_canon(powtab_eval(tab, i));
}
void launch_fill_powers(u64* out, u64 len, const u64* powtab, cudaStream_t st) {
    u64 blocks = (len + 255) / 256;
    if (blocks > 4096) blocks = 4096;
    { k_fill_powers<<<(unsigned)blocks, 256, 0, st>>>(out, len, powtab); ++g_gl_launches; }
}

template <int M>
static void launch_ntt16_m(const ntt16_args& a, bool inverse, bool strided, dim3 grid, unsigned threads, size_t smem,
                           cudaStream_t st) {
    auto set = [&](void (*kern)(ntt16_args)) {
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        kern<<<grid, threads, smem, st>>>(a);
        ++g_gl_launches;
    };
    if (inverse) {
        if (strided) set(k_ntt16<M, true, true>);
        else set(k_ntt16<M, true, false>);
    } else {
        if (strided) set(k_ntt16<M, false, true>);
        else set(k_ntt16<M, false, false>);
    }
}

bool launch_ntt16(const ntt16_args& a0, unsigned M, bool inverse, u64 n, u32 columns, u32 cosets, cudaStream_t st) {
    if (M < 8 || M > 10) return false;
    ntt16_args a = a0;
    const bool strided = a.s != 0;
    unsigned logW = 13 - M;
    if (strided) {
        if (logW > a.s) logW = a.s;
        if (logW + (M - 8) > 5) logW = 5 - (M - 8);
    } else {
        u64 blocks = n >> M;
        while (((u64)1 << logW) > blocks) logW--;
    }
    unsigned threads = (1u << (M - 4)) << logW;
    if (threads < 32) return false;
    a.logW = logW;
    a.n = n;
    u64 tiles = n >> (M + logW);
    dim3 grid((unsigned)tiles, columns, cosets);
    size_t smem = (((size_t)1 << (M + logW)) + ((size_t)1 << M)) * sizeof(u64);
    switch (M) {
        case 8: launch_ntt16_m<8>(a, inverse, strided, grid, threads, smem, st); break;
        case 9: launch_ntt16_m<9>(a, inverse, strided, grid, threads, smem, st); break;
        default: launch_ntt16_m<10>(a, inverse, strided, grid, threads, smem, st); break;
    }
    return true;
}
