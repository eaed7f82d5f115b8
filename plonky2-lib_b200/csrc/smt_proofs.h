// Internal interface of smt_proofs.cu: the process proofs of a batch of inserts and updates (second half).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gl_b200.h"

struct smt_proof_buffers {
    uint64_t m;                  // events: the sets, then (gl_smt_find_batch) the queries
    uint64_t m_sets;             // events with time >= m_sets are queries
    uint32_t stride;             // siblings kept per event while sweeping = max(1, bottom)
    uint32_t bottom;             // deepest LCP between different keys + 1: below it every position holds one key
    // from smt_build_prepare (sorted by path order)
    const uint64_t* keys;        // [m][4] input order
    const uint64_t* values;      // [m][4] input order
    const uint64_t* rk;          // [4][m] bit-reversed limbs, input order
    const uint32_t* perm;        // sorted position -> input index = insertion time
    const uint16_t* lcp;         // [m - 1]
    const uint64_t* leafh;       // [m][4] sorted order
    // segmentation and (segment, time) order of the depth being computed (cur) and of the one below it (nxt)
    uint32_t *a_cur, *end_cur, *ord_cur, *inv_cur, *tm_cur;
    uint32_t *a_nxt, *end_nxt, *ord_nxt, *inv_nxt, *tm_nxt;
    uint64_t *val_cur, *val_nxt; // [m][4]
    uint32_t *dc_cur, *dc_nxt;   // [m] distinct keys below the position after the event
    uint32_t *rep_cur, *rep_nxt; // [m] while dc == 1: sorted position of the latest event of that one key
    uint32_t* pos_of_time;       // [m] insertion time -> sorted position
    uint32_t *deep_dc, *deep_rep;// [m] per event: keys below (and, if one, which) the sibling of the deepest internal node on its path
    uint32_t* other;             // [m] per position of the current order: index in val_nxt of the other child's value
    uint8_t* bit;                // [m] path bit of the key at this depth (1: the key is in the right child)
    // per key, indexed by insertion time
    uint64_t* sib;               // [m][stride][4]
    uint32_t* stop_depth;        // where `find` stops
    uint32_t* stop_old;          // sorted position of the key found there, SP_NONE for an empty slot
    gl_smt_proof_hdr* hdr;       // [m] process proofs (set mode)
    gl_smt_inclusion_hdr* inc;   // [m - m_sets] inclusion proofs (find mode)
};

size_t smt_proof_temp_bytes(uint64_t m);
int smt_proofs_check_values(const uint64_t* values, uint64_t m, uint32_t* bad, cudaStream_t st);
int smt_proofs_sweep(smt_proof_buffers p, int dmax, const uint32_t* hist, uint32_t* counts /* [m + 1] */, void* tmp,
                     size_t tmp_bytes, cudaStream_t st);
int smt_proofs_offsets(const uint32_t* counts, uint64_t* off, uint64_t m, void* tmp, size_t tmp_bytes, cudaStream_t st);
void smt_proofs_gather(const smt_proof_buffers& p, const uint64_t* off, uint64_t cap, uint64_t* pool, cudaStream_t st);
