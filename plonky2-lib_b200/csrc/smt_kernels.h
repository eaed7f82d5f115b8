// Internal interface of smt_kernels.cu: bulk build of the compact sparse Merkle tree.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
extern std::atomic<unsigned long long> g_gl_launches;

struct smt_build_buffers {
    // inputs (device)
    const uint64_t* keys;    // [m][4]
    const uint64_t* values;  // [m][4]
    uint64_t m;
    // work arrays (device), all sized for m entries
    uint64_t* rk;        // [4][m] bit-reversed limbs, sorted, limb-major (sort key planes)
    uint64_t* rk_alt;    // [m] double buffer for the radix sort
    uint32_t* perm;      // [m] sorted position -> input index
    uint32_t* perm_alt;  // [m]
    uint64_t* leafh;     // [m][4] leaf hashes in SORTED order
    uint16_t* lcp;       // [m]: lcp[i] = common path-bit prefix of sorted keys i and i+1 (i < m-1)
    uint64_t* val_first; // [m][4]
    uint64_t* val_last;  // [m][4]
    uint32_t* end_of;    // [m]
    uint32_t* start_of;  // [m]
    uint16_t* form_depth;// [m]: 0xFFFF = key is not the first key of a live group
    uint8_t* last_valid; // [m]
    uint32_t* hist;      // [257] pairs per lcp value (256 = duplicate keys)
    uint32_t* zero_values; // counter of entries whose value is all zero (mod p), or null when zero values are allowed
    void* sort_tmp;
    size_t sort_tmp_bytes;
    // outputs (device)
    uint64_t* nodes;        // [nodes_cap][12] hash, left, right of every internal node, or null
    uint64_t nodes_cap;
    unsigned long long* node_count;
    uint64_t* leaf_hashes;  // [m][4] in INPUT order, or null
};

size_t smt_sort_temp_bytes(uint64_t m);
// sorts by path order, hashes the leaves, fills lcp + histogram
int smt_build_prepare(const smt_build_buffers& b, cudaStream_t st);
// one level of the tree (depth d): chain the live groups that do not split here, merge the pairs that do
void smt_build_level(const smt_build_buffers& b, unsigned d, cudaStream_t st);
