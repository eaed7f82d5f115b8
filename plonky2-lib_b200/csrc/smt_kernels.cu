// N2: bulk construction of the compact (circomlib-style) sparse Merkle tree of src/smt/tree.rs from m distinct
// (key, value) pairs -- the root and node set that m successive `set` calls (tree.rs:143-155, insert :255-387)
// leave behind, computed level-synchronously.
//
// Path bit i of a key = bit (i mod 64) of element i / 64, LSB first (src/smt/goldilocks_poseidon/mod.rs:27-48).
// After sorting the keys by path, the tree is the Cartesian tree of the adjacent-pair LCP array: the pair
// (i, i+1) with common prefix length d is the internal node at depth d where the two sides diverge; a set of
// >= 2 keys sharing a longer prefix than its parent's depth climbs through one-child nodes H(x, 0) / H(0, x);
// a single key is hoisted (its leaf hash stands for the whole subtree).  Levels are processed from the deepest
// LCP up to depth 0; every node of a level is independent.
//
// This file is #included at the end of hash_kernels.cu (the Poseidon constants live in that translation unit);
// it uses poseidon_permute_call / two_to_one_call defined there.
#include <cub/device/device_radix_sort.cuh>

#include "smt_kernels.h"

#define SMT_BLOCK 128
#define smt_permute_call poseidon_permute_call
#define smt_two_to_one two_to_one_call

__global__ void __launch_bounds__(256) k_smt_keys(const u64* __restrict__ keys, u64 m, u64* __restrict__ rk, u32* __restrict__ perm) {
    u64 i = blockIdx.x * (u64)256 + threadIdx.x;
    if (i >= m) return;
#pragma unroll
    for (int j = 0; j < 4; j++) rk[(u64)j * m + i] = __brevll(gl_canon(keys[4 * i + j]));
    perm[i] = (u32)i;
}
// gather plane `src` through perm into dst (sort key of the next radix pass)
__global__ void __launch_bounds__(256) k_smt_gather(const u64* __restrict__ src, const u32* __restrict__ perm, u64 m, u64* __restrict__ dst) {
    u64 i = blockIdx.x * (u64)256 + threadIdx.x;
    if (i < m) dst[i] = src[perm[i]];
}

// leaf hashes in sorted order (+ input order copy), and the LCP of every adjacent pair
__global__ void __launch_bounds__(SMT_BLOCK)
k_smt_leaves(smt_build_buffers b) {
    u64 i = blockIdx.x * (u64)SMT_BLOCK + threadIdx.x;
    if (i >= b.m) return;
    const u32 src = b.perm[i];
    u64 s[12];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        s[j] = b.keys[4 * (u64)src + j];
        s[4 + j] = b.values[4 * (u64)src + j];
        s[8 + j] = 0;
    }
    // SparseMerkleTree::insert: "value must be non-zero" (a bulk build has no removals)
    if (b.zero_values && (gl_canon(s[4]) | gl_canon(s[5]) | gl_canon(s[6]) | gl_canon(s[7])) == 0) atomicAdd(b.zero_values, 1u);
    smt_permute_call(s);
    s[0] = 1; s[1] = 1; s[2] = 0; s[3] = 1;   // hash_no_pad([k, v, 1, 1, 0, 1]): second chunk overwrites lanes 0..3
    smt_permute_call(s);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        u64 h = gl_canon(s[j]);
        b.leafh[4 * i + j] = h;
        if (b.leaf_hashes) b.leaf_hashes[4 * (u64)src + j] = h;
    }
    b.form_depth[i] = 0xFFFF;
    b.last_valid[i] = 0;
    if (i + 1 < b.m) {
        unsigned l = 256;
#pragma unroll
        for (int j = 3; j >= 0; j--) {
            u64 x = b.rk[(u64)j * b.m + b.perm[i]] ^ b.rk[(u64)j * b.m + b.perm[i + 1]];
            if (x) l = 64 * j + (unsigned)__clzll((long long)x);
        }
        b.lcp[i] = (uint16_t)l;
        atomicAdd(b.hist + l, 1u);
    }
}

GL_D int smt_path_bit(const smt_build_buffers& b, u64 sorted_index, unsigned d) {
    // rk holds bit-reversed limbs: path bit d of limb d/64 is bit 63 - (d mod 64)
    u64 w = b.rk[(u64)(d >> 6) * b.m + b.perm[sorted_index]];
    return (int)((w >> (63 - (d & 63))) & 1);
}
GL_D void smt_emit(const smt_build_buffers& b, const u64 h[4], const u64 l[4], const u64 r[4]) {
    if (!b.nodes) return;
    unsigned long long at = atomicAdd(b.node_count, 1ULL);
    if (at >= b.nodes_cap) return;
    u64* o = b.nodes + 12 * at;
#pragma unroll
    for (int j = 0; j < 4; j++) { o[j] = h[j]; o[4 + j] = l[j]; o[8 + j] = r[j]; }
}

// live groups (>= 2 keys, formed deeper than d) that do not split at depth d climb one level: H(v, 0) or H(0, v)
__global__ void __launch_bounds__(SMT_BLOCK)
k_smt_chain(smt_build_buffers b, unsigned d) {
    u64 s = blockIdx.x * (u64)SMT_BLOCK + threadIdx.x;
    if (s >= b.m) return;
    const unsigned fd = b.form_depth[s];
    if (fd == 0xFFFF || fd <= d) return;
    const u64 e = b.end_of[s];
    if ((s > 0 && b.lcp[s - 1] == d) || (e + 1 < b.m && b.lcp[e] == d)) return;   // merged at this depth instead
    u64 v[4], z[4] = {0, 0, 0, 0}, o[4];
#pragma unroll
    for (int j = 0; j < 4; j++) v[j] = b.val_first[4 * s + j];
    const int bit = smt_path_bit(b, s, d);
    if (bit) smt_two_to_one(z, v, o);
    else smt_two_to_one(v, z, o);
    smt_emit(b, o, bit ? z : v, bit ? v : z);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        b.val_first[4 * s + j] = o[j];
        b.val_last[4 * e + j] = o[j];
    }
}
// the pair (i, i+1) with LCP d is the internal node at depth d: left = the side whose bit d is 0 = key i's side
__global__ void __launch_bounds__(SMT_BLOCK)
k_smt_merge(smt_build_buffers b, unsigned d) {
    u64 i = blockIdx.x * (u64)SMT_BLOCK + threadIdx.x;
    if (i + 1 >= b.m || b.lcp[i] != d) return;
    u64 lv[4], rv[4], o[4];
    u64 ls = i, re = i + 1;
    if (b.last_valid[i]) {
        ls = b.start_of[i];
#pragma unroll
        for (int j = 0; j < 4; j++) lv[j] = b.val_last[4 * i + j];
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) lv[j] = b.leafh[4 * i + j];
    }
    if (b.form_depth[i + 1] != 0xFFFF) {
        re = b.end_of[i + 1];
#pragma unroll
        for (int j = 0; j < 4; j++) rv[j] = b.val_first[4 * (i + 1) + j];
    } else {
#pragma unroll
        for (int j = 0; j < 4; j++) rv[j] = b.leafh[4 * (i + 1) + j];
    }
    smt_two_to_one(lv, rv, o);
    smt_emit(b, o, lv, rv);
    b.last_valid[i] = 0;
    b.form_depth[i + 1] = 0xFFFF;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        b.val_first[4 * ls + j] = o[j];
        b.val_last[4 * re + j] = o[j];
    }
    b.end_of[ls] = (u32)re;
    b.start_of[re] = (u32)ls;
    b.form_depth[ls] = (uint16_t)d;
    b.last_valid[re] = 1;
}

size_t smt_sort_temp_bytes(uint64_t m) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const u64*)nullptr, (u64*)nullptr, (const u32*)nullptr, (u32*)nullptr, (int)m);
    return bytes;
}

int smt_build_prepare(const smt_build_buffers& b, cudaStream_t st) {
    const u64 m = b.m;
    const unsigned blocks256 = (unsigned)((m + 255) / 256);
    k_smt_keys<<<blocks256, 256, 0, st>>>(b.keys, m, b.rk, b.perm);
    ++g_gl_launches;
    // LSD radix sort over the four 64-bit planes, least significant path bits (element 3) first; stable
    u32 *pin = b.perm, *pout = b.perm_alt;
    for (int plane = 3; plane >= 0; plane--) {
        k_smt_gather<<<blocks256, 256, 0, st>>>(b.rk + (u64)plane * m, pin, m, b.rk_alt);
        size_t tmp = b.sort_tmp_bytes;
        // keys in rk_alt are consumed (sorted keys written over leafh scratch is not needed: use val_first as sink)
        cudaError_t e = cub::DeviceRadixSort::SortPairs(b.sort_tmp, tmp, (const u64*)b.rk_alt, (u64*)b.val_first, (const u32*)pin,
                                                        pout, (int)m, 0, 64, st);
        if (e != cudaSuccess) return (int)e;
        g_gl_launches += 4;
        u32* t = pin; pin = pout; pout = t;
    }
    // four passes: the result is back in b.perm
    k_smt_leaves<<<(unsigned)((m + SMT_BLOCK - 1) / SMT_BLOCK), SMT_BLOCK, 0, st>>>(b);
    ++g_gl_launches;
    return 0;
}

void smt_build_level(const smt_build_buffers& b, unsigned d, cudaStream_t st) {
    const unsigned blocks = (unsigned)((b.m + SMT_BLOCK - 1) / SMT_BLOCK);
    k_smt_chain<<<blocks, SMT_BLOCK, 0, st>>>(b, d);
    k_smt_merge<<<blocks, SMT_BLOCK, 0, st>>>(b, d);
    g_gl_launches += 2;
}
