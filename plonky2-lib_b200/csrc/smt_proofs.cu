// N2, second half: the m SparseMerkleProcessProofs that m successive `tree.set(key_t, value_t)` calls
// (src/smt/tree.rs:143-155, find :588-676, update :174-253, insert :255-387) return when they start from an EMPTY
// tree and no value is zero -- all at once, although proof t is a statement about the tree the first t calls left.
// A key may occur several times: its first occurrence is an insert, the later ones are updates.
//
// Sort the events (key, value, time) by path order, ties (the same key) in time order.  The events below a trie
// position of depth d are a contiguous segment (maximal run of adjacent pairs with LCP >= d).  A position holds:
// nothing (hash 0), one key (its leaf hash, hoisted), or >= 2 keys (an internal node, possibly with one empty
// child).  Order the events of every segment by time:
//   Val_d[i] = hash at that position right after the i-th event,  dc_d[i] = distinct keys below it at that time
//            = the value of the child holding the event's key                 when dc == 1 (hoisted leaf),
//              H(child slot 0, child slot 1)                                   when dc >= 2,
// the child slots being Val_{d+1} of the two depth-(d+1) segments inside, taken at the same time: the child with
// the event's key at the event's own position in the depth-(d+1) order, the other one by a binary search over
// times.  The same search gives the time order of depth d without sorting (segment start + rank in the own child +
// events of the other child before it).  The sweep starts from the same-key groups ("depth 256") and runs from the
// deepest LCP between different keys up to depth 0.  With n = distinct keys below the position just before event t:
// n >= 2 means an internal node on t's `find` path and the other child's value just before t is siblings[d]; the
// shallowest depth with n <= 1 is where `find` stops -- an empty slot (is_old0), another key's leaf (old_key /
// old_value), or the key's own leaf (an update: old_value = its previous value).  Depth 0 yields old_root / new_root.
// Work: one permutation per (event, depth above its stopping point) -- the hashes the sequential calls compute,
// ~ m log2 m in total.
//
// #included at the end of hash_kernels.cu after smt_kernels.cu (shares the Poseidon constants and helpers).
#include <cub/device/device_scan.cuh>

#include "smt_proofs.h"

#define SP_NONE 0xFFFFFFFFu

// path bit d of the key with input index `src` (rk holds bit-reversed limbs: bit d of limb d / 64 is bit 63 - d % 64)
GL_D int smt_path_bit_perm(const u64* __restrict__ rk, u64 m, u32 src, unsigned d) {
    return (int)((rk[(u64)(d >> 6) * m + src] >> (63 - (d & 63))) & 1);
}

// "depth 256": segments = the events of one key (already in time order: the radix sort is stable)
// `set(key, value)`: a zero value removes.  What an event does depends on whether its key is in the tree just before it,
// i.e. on the value of the previous event of the same key (sorted position j - 1 when it is in the same group).
enum { SP_NOOP = 0, SP_UPDATE = 1, SP_INSERT = 2, SP_REMOVE = 3, SP_FIND = 4 };   // 0..3 = ProcessMerkleProofRole
// Events with time >= m_sets are `tree.find(key)` queries against the tree the sets left (gl_smt_find_batch): they
// come after every set of their key in the sorted order and change nothing.
GL_D bool sp_is_find(const smt_proof_buffers& p, u32 j) { return p.perm[j] >= p.m_sets; }
// the last `set` of event j's key at or before j (sorted position), or SP_NONE
GL_D u32 sp_last_set(const smt_proof_buffers& p, u32 j) {
    while (sp_is_find(p, j)) {
        if (j == 0 || p.lcp[j - 1] < 256) return SP_NONE;
        j--;
    }
    return j;
}
GL_D bool sp_value_nonzero(const smt_proof_buffers& p, u32 j) {
    const u64* v = p.values + 4 * (u64)p.perm[j];
    return (gl_canon(v[0]) | gl_canon(v[1]) | gl_canon(v[2]) | gl_canon(v[3])) != 0;
}
GL_D bool sp_present(const smt_proof_buffers& p, u32 j) {            // is the key in the tree right after event j?
    const u32 s = sp_last_set(p, j);
    return s != SP_NONE && sp_value_nonzero(p, s);
}
GL_D int sp_kind(const smt_proof_buffers& p, u32 j) {
    if (sp_is_find(p, j)) return SP_FIND;
    const bool same_key_before = j > 0 && p.lcp[j - 1] >= 256;
    const bool was = same_key_before && sp_present(p, j - 1), is = sp_value_nonzero(p, j);
    return is ? (was ? SP_UPDATE : SP_INSERT) : (was ? SP_REMOVE : SP_NOOP);
}

__global__ void __launch_bounds__(256) k_sp_init(smt_proof_buffers p) {
    u64 j = blockIdx.x * (u64)256 + threadIdx.x;
    if (j >= p.m) return;
    // a_nxt comes from the scan of the depth-256 flags; ends by scatter
    if (j + 1 == p.m || p.lcp[j] < 256) p.end_nxt[p.a_nxt[j]] = (u32)j + 1;
    p.ord_nxt[j] = (u32)j;
    p.inv_nxt[j] = (u32)j;
    p.tm_nxt[j] = p.perm[j];
    const u32 ls = sp_last_set(p, (u32)j);              // j itself unless j is a query
    const bool present = ls != SP_NONE && sp_value_nonzero(p, ls);
#pragma unroll
    for (int k = 0; k < 4; k++) p.val_nxt[4 * j + k] = present ? p.leafh[4 * (u64)ls + k] : 0;
    p.dc_nxt[j] = present ? 1 : 0;
    p.rep_nxt[j] = present ? ls : (u32)j;
    const u32 t = p.perm[j];
    p.pos_of_time[t] = (u32)j;
    // below every LCP between different keys the position holds this key alone, or nothing
    const int kind = sp_kind(p, (u32)j);
    p.stop_depth[t] = p.bottom;
    p.stop_old[t] = (kind == SP_UPDATE || kind == SP_REMOVE) ? (u32)j - 1 : SP_NONE;
    p.deep_dc[t] = 0;
    p.deep_rep[t] = SP_NONE;
}

// start flags of the depth-d segments, as scan input (index where a segment starts, else 0)
__global__ void __launch_bounds__(256) k_sp_flags(smt_proof_buffers p, unsigned d, u32* __restrict__ flag_idx) {
    u64 j = blockIdx.x * (u64)256 + threadIdx.x;
    if (j >= p.m) return;
    flag_idx[j] = (j == 0 || p.lcp[j - 1] < d) ? (u32)j : 0u;
}
// segment ends, indexed by segment start
__global__ void __launch_bounds__(256) k_sp_ends(smt_proof_buffers p, unsigned d) {
    u64 j = blockIdx.x * (u64)256 + threadIdx.x;
    if (j >= p.m) return;
    if (j + 1 == p.m || p.lcp[j] < d) p.end_cur[p.a_cur[j]] = (u32)j + 1;
}
// The (segment, time) order of depth d without sorting: a depth-d segment is the concatenation of (at most) two
// depth-(d+1) segments that are already in time order, so the new position of a key is
//   segment start + its rank in its own child + the number of keys of the other child inserted before it,
// and that count is the binary search the hash needs anyway (kept in `other` for k_sp_level).
__global__ void __launch_bounds__(256) k_sp_place(smt_proof_buffers p, unsigned d) {
    u64 j = blockIdx.x * (u64)256 + threadIdx.x;
    if (j >= p.m) return;
    const u32 a = p.a_cur[j], e = p.end_cur[a];
    const u32 a1 = p.a_nxt[j], e1 = p.end_nxt[a1];
    const u32 t = p.perm[j];
    u32 lo = 0, hi = 0, bit;
    if (a1 > a) { lo = a; hi = a1; bit = 1; }          // the key is in the right child
    else if (e1 < e) { lo = e1; hi = e; bit = 0; }
    else bit = (u32)smt_path_bit_perm(p.rk, p.m, t, d);
    u32 l = lo, h = hi;
    while (l < h) {                                     // keys of the other child inserted before t
        u32 mid = (l + h) >> 1;
        if (p.tm_nxt[mid] < t) l = mid + 1;
        else h = mid;
    }
    const u32 i = a + (p.inv_nxt[j] - a1) + (l - lo);
    p.ord_cur[i] = (u32)j;
    p.inv_cur[j] = i;
    p.tm_cur[i] = t;
    p.other[i] = (l > lo ? (l - 1) : SP_NONE) | 0u;     // where the other child's slot value at that time lives
    p.bit[i] = (uint8_t)bit;
}

__global__ void __launch_bounds__(SMT_BLOCK) k_sp_level(smt_proof_buffers p, unsigned d, int merged) {
    u64 i = blockIdx.x * (u64)SMT_BLOCK + threadIdx.x;
    if (i >= p.m) return;
    const u32 j = p.ord_cur[i];
    const u32 t = p.tm_cur[i];
    const u32 at = p.inv_nxt[j];                       // the event's position in the order one level down
    const u32 o = merged ? p.other[i] : SP_NONE;       // latest earlier event of the other child, if any
    const u32 dc_own = p.dc_nxt[at], dc_sib = o != SP_NONE ? p.dc_nxt[o] : 0;
    const u32 dc = dc_own + dc_sib;
    u64 own[4], sib[4] = {0, 0, 0, 0}, out[4];
#pragma unroll
    for (int k = 0; k < 4; k++) own[k] = p.val_nxt[4 * (u64)at + k];
    if (o != SP_NONE) {
#pragma unroll
        for (int k = 0; k < 4; k++) sib[k] = p.val_nxt[4 * (u64)o + k];
    }
    p.dc_cur[i] = dc;
    if (dc <= 1) {
        // nothing, or one key below this position: its leaf hash stands for the subtree, whichever child holds it
        const bool in_own = dc_own == 1;
#pragma unroll
        for (int k = 0; k < 4; k++) out[k] = dc == 0 ? 0 : (in_own ? own[k] : sib[k]);
        p.rep_cur[i] = dc == 0 ? SP_NONE : (in_own ? p.rep_nxt[at] : p.rep_nxt[o]);
    } else {
        const int bit = merged ? (int)p.bit[i] : smt_path_bit_perm(p.rk, p.m, t, d);
        if (bit) smt_two_to_one(sib, own, out);
        else smt_two_to_one(own, sib, out);
        p.rep_cur[i] = SP_NONE;
    }
#pragma unroll
    for (int k = 0; k < 4; k++) p.val_cur[4 * i + k] = out[k];
    // the proof of event t at this depth
    const int kind = sp_kind(p, j);
    const u32 before = dc - (kind == SP_INSERT ? 1u : 0u) + (kind == SP_REMOVE ? 1u : 0u);   // distinct keys just before t
    if (before >= 2) {
        u64* s = p.sib + ((u64)t * p.stride + d) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) s[k] = sib[k];
        if (d + 1 == p.stop_depth[t]) {                // the deepest internal node on the path: what hangs next to the leaf
            p.deep_dc[t] = dc_sib;
            p.deep_rep[t] = dc_sib == 1 ? p.rep_nxt[o] : SP_NONE;
        }
    } else {
        p.stop_depth[t] = d;
        if (kind == SP_UPDATE || kind == SP_REMOVE) p.stop_old[t] = j - 1;   // its own leaf: the previous event of the key
        else if (before == 0) p.stop_old[t] = SP_NONE;
        else if (kind == SP_FIND) p.stop_old[t] = dc_own == 1 ? p.rep_nxt[at] : p.rep_nxt[o];   // a query adds nothing to its child
        else p.stop_old[t] = dc_sib == 1 ? p.rep_nxt[o] : (at ? p.rep_nxt[at - 1] : SP_NONE);   // the one other key, in either child
    }
}

// after depth 0 (one segment, time order): roots and trimmed sibling counts
GL_D void sp_copy_kv(const smt_proof_buffers& p, u32 sorted_pos, u64* key_out, u64* value_out) {
    const u64 src = p.perm[sorted_pos];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        key_out[k] = gl_canon(p.keys[4 * src + k]);
        value_out[k] = gl_canon(p.values[4 * src + k]);
    }
}
// tree.find(key) against the final tree (src/smt/tree.rs:588-676): SparseMerkleInclusionProof
GL_D void sp_inclusion(const smt_proof_buffers& p, u64 t, u32 j, u32* __restrict__ counts) {
    gl_smt_inclusion_hdr* h = p.inc + (t - p.m_sets);
    u64 key[4], unused[4];
    sp_copy_kv(p, j, key, unused);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        h->root[k] = p.val_cur[4 * t + k];
        h->key[k] = key[k];
        h->value[k] = h->not_found_key[k] = h->not_found_value[k] = 0;
    }
    h->found = 0;
    h->is_old0 = 0;
    const u32 ls = sp_last_set(p, j);
    if (ls != SP_NONE && sp_value_nonzero(p, ls)) {
        h->found = 1;
        sp_copy_kv(p, ls, unused, h->value);
    } else if (p.stop_old[t] == SP_NONE) {
        h->is_old0 = 1;
    } else {
        sp_copy_kv(p, p.stop_old[t], h->not_found_key, h->not_found_value);
    }
    counts[t] = p.stop_depth[t];   // find returns every sibling down to where it stops
}
__global__ void __launch_bounds__(256) k_sp_roots(smt_proof_buffers p, u32* __restrict__ counts) {
    u64 t = blockIdx.x * (u64)256 + threadIdx.x;
    if (t >= p.m) return;
    const u32 j = p.pos_of_time[t];
    if (p.m_sets < p.m) {          // gl_smt_find_batch: only the queries produce output
        if (t >= p.m_sets) sp_inclusion(p, t, j, counts);
        else counts[t] = 0;
        return;
    }
    gl_smt_proof_hdr* h = p.hdr + t;
    const int kind = sp_kind(p, j);
    u64 key[4], value[4], zero4[4] = {0, 0, 0, 0};
    sp_copy_kv(p, j, key, value);
#pragma unroll
    for (int k = 0; k < 4; k++) {
        h->new_root[k] = p.val_cur[4 * t + k];
        h->old_root[k] = t ? p.val_cur[4 * (t - 1) + k] : 0;
        h->old_key[k] = h->old_value[k] = h->new_key[k] = h->new_value[k] = 0;
    }
    h->fnc = (uint32_t)kind;
    h->is_old0 = 0;
    const u64* s = p.sib + (u64)t * p.stride * 4;
    auto is_zero = [&](u32 lvl) { return (s[4 * lvl] | s[4 * lvl + 1] | s[4 * lvl + 2] | s[4 * lvl + 3]) == 0; };
    u32 ns = p.stop_depth[t];
    const u32 so = p.stop_old[t];
    if (kind == SP_NOOP) {                 // removing a key that is not there: nothing happens, nothing to show
#pragma unroll
        for (int k = 0; k < 4; k++) h->old_key[k] = h->new_key[k] = key[k];
        h->is_old0 = 1;
        ns = 0;
    } else if (kind == SP_INSERT) {
#pragma unroll
        for (int k = 0; k < 4; k++) { h->new_key[k] = key[k]; h->new_value[k] = value[k]; }
        if (so == SP_NONE) h->is_old0 = 1;
        else sp_copy_kv(p, so, h->old_key, h->old_value);
        while (ns > 0 && is_zero(ns - 1)) ns--;          // insert trims the trailing zero siblings
    } else if (kind == SP_UPDATE) {
        sp_copy_kv(p, so, h->old_key, h->old_value);      // the key itself with its previous value
#pragma unroll
        for (int k = 0; k < 4; k++) { h->new_key[k] = key[k]; h->new_value[k] = value[k]; }
    } else {                               // remove = the insert of this key into the tree it leaves, old and new swapped
        sp_copy_kv(p, so, h->old_key, h->old_value);
        if (ns > 0 && p.deep_dc[t] == 1) {                // a lone leaf hung next to it: that leaf moves up
            sp_copy_kv(p, p.deep_rep[t], h->new_key, h->new_value);
            ns--;                                         // its own level goes away with the node
            while (ns > 0 && is_zero(ns - 1)) ns--;
        } else {                                          // an internal node next to it (or the tree is empty now)
#pragma unroll
            for (int k = 0; k < 4; k++) h->new_key[k] = key[k];
            h->is_old0 = 1;
        }
    }
    (void)zero4;
    counts[t] = ns;
}
__global__ void __launch_bounds__(256) k_sp_gather(smt_proof_buffers p, const u64* __restrict__ off, u64 cap, u64* __restrict__ pool) {
    u64 t = blockIdx.x * (u64)256 + threadIdx.x;
    if (t >= p.m) return;
    const u64 o = off[t], n = off[t + 1] - o;
    if (o + n > cap) return;
    const u64* s = p.sib + (u64)t * p.stride * 4;
    for (u64 k = 0; k < 4 * n; k++) pool[4 * o + k] = s[k];
}
// any value that is all zero?  (`set` with the default value is a removal, not an insert)
__global__ void __launch_bounds__(256) k_sp_check_values(const u64* __restrict__ values, u64 m, u32* __restrict__ bad) {
    u64 t = blockIdx.x * (u64)256 + threadIdx.x;
    if (t >= m) return;
    if ((gl_canon(values[4 * t]) | gl_canon(values[4 * t + 1]) | gl_canon(values[4 * t + 2]) | gl_canon(values[4 * t + 3])) == 0) atomicAdd(bad, 1u);
}

size_t smt_proof_temp_bytes(uint64_t m) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceScan::InclusiveScan(nullptr, b, (const u32*)nullptr, (u32*)nullptr, cub::Max(), (int)m);
    cub::DeviceScan::ExclusiveSum(nullptr, c, (const u32*)nullptr, (u64*)nullptr, (int)m + 1);
    size_t r = a > b ? a : b;
    return r > c ? r : c;
}

int smt_proofs_check_values(const u64* values, u64 m, u32* bad, cudaStream_t st) {
    k_sp_check_values<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(values, m, bad);
    ++g_gl_launches;
    return 0;
}

// depths dmax .. 0; hist[d] = number of adjacent pairs with LCP == d (segments merge only at those depths).
// The two sets of order arrays alternate: `nxt` is the live set (one level down), `cur` the one being written.
int smt_proofs_sweep(smt_proof_buffers p, int dmax, const uint32_t* hist, u32* counts, void* tmp, size_t tmp_bytes, cudaStream_t st) {
    const u64 m = p.m;
    const unsigned b256 = (unsigned)((m + 255) / 256), bl = (unsigned)((m + SMT_BLOCK - 1) / SMT_BLOCK);
    {   // same-key groups: starts by the scan of the depth-256 flags
        k_sp_flags<<<b256, 256, 0, st>>>(p, 256u, counts);
        size_t tb = tmp_bytes;
        cudaError_t e = cub::DeviceScan::InclusiveScan(tmp, tb, (const u32*)counts, p.a_nxt, cub::Max(), (int)m, st);
        if (e != cudaSuccess) return (int)e;
        k_sp_init<<<b256, 256, 0, st>>>(p);
        g_gl_launches += 4;
    }
    u32* spare[5] = {p.a_cur, p.end_cur, p.ord_cur, p.inv_cur, p.tm_cur};
    for (int d = dmax; d >= 0; d--) {
        const bool merges = hist[d] != 0;
        if (merges) {
            p.a_cur = spare[0]; p.end_cur = spare[1]; p.ord_cur = spare[2]; p.inv_cur = spare[3]; p.tm_cur = spare[4];
            // new segmentation: starts by a max-scan of the start flags, ends by scatter, then the merged time order
            k_sp_flags<<<b256, 256, 0, st>>>(p, (unsigned)d, counts);
            size_t tb = tmp_bytes;
            cudaError_t e = cub::DeviceScan::InclusiveScan(tmp, tb, (const u32*)counts, p.a_cur, cub::Max(), (int)m, st);
            if (e != cudaSuccess) return (int)e;
            k_sp_ends<<<b256, 256, 0, st>>>(p, (unsigned)d);
            k_sp_place<<<b256, 256, 0, st>>>(p, (unsigned)d);
            g_gl_launches += 5;
        } else {
            // no pair diverges at this depth: same segments and order as one level down (pure chain steps)
            p.a_cur = p.a_nxt; p.end_cur = p.end_nxt; p.ord_cur = p.ord_nxt; p.inv_cur = p.inv_nxt; p.tm_cur = p.tm_nxt;
        }
        k_sp_level<<<bl, SMT_BLOCK, 0, st>>>(p, (unsigned)d, merges ? 1 : 0);
        ++g_gl_launches;
        if (merges) {   // this depth becomes "one level down"; the old live set is the spare one now
            spare[0] = p.a_nxt; spare[1] = p.end_nxt; spare[2] = p.ord_nxt; spare[3] = p.inv_nxt; spare[4] = p.tm_nxt;
            p.a_nxt = p.a_cur; p.end_nxt = p.end_cur; p.ord_nxt = p.ord_cur; p.inv_nxt = p.inv_cur; p.tm_nxt = p.tm_cur;
        }
        u64* v = p.val_nxt; p.val_nxt = p.val_cur; p.val_cur = v;
        u32* x = p.dc_nxt; p.dc_nxt = p.dc_cur; p.dc_cur = x;
        x = p.rep_nxt; p.rep_nxt = p.rep_cur; p.rep_cur = x;
    }
    p.val_cur = p.val_nxt;   // the depth-0 values (time order)
    k_sp_roots<<<b256, 256, 0, st>>>(p, counts);
    ++g_gl_launches;
    return 0;
}

int smt_proofs_offsets(const u32* counts, u64* off, u64 m, void* tmp, size_t tmp_bytes, cudaStream_t st) {
    size_t tb = tmp_bytes;
    cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tb, counts, off, (int)m + 1, st);
    g_gl_launches += 2;
    return (int)e;
}
void smt_proofs_gather(const smt_proof_buffers& p, const u64* off, u64 cap, u64* pool, cudaStream_t st) {
    k_sp_gather<<<(unsigned)((p.m + 255) / 256), 256, 0, st>>>(p, off, cap, pool);
    ++g_gl_launches;
}
