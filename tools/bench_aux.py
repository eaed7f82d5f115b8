#!/usr/bin/env python
"""Throughput of the secondary entry points of the C ABI (device-resident buffers, blocking calls, wall clock):
Poseidon batches, row-major MerkleTree::new, SMT process-proof batches (BASELINE config 4), FRI layer pieces.
Prints one JSON object; numbers quoted in DESIGN.md come from here (profiles/r1_bench_aux.json)."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

glb = importlib.import_module("plonky2-lib_b200")

ctx = glb.Context(0)
lib, N = ctx._lib, glb._native
import ctypes as C  # noqa: E402

dev = torch.device("cuda", 0)
P = glb.host.P
out = {}


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t)
    return best


def rand_dev(shape):
    return torch.randint(0, 2**62, shape, dtype=torch.int64, device=dev)


m = 1 << 22
st = rand_dev((m, 12))
t = timeit(lambda: ctx.check(lib.gl_poseidon_permute_batch(ctx._h, st.data_ptr(), m, N.GL_DEVICE)))
out["permute_batch"] = {"states": m, "ms": t * 1e3, "perms_per_s": m / t}
l, r, d = rand_dev((m, 4)), rand_dev((m, 4)), torch.empty((m, 4), dtype=torch.int64, device=dev)
t = timeit(lambda: ctx.check(lib.gl_poseidon_two_to_one_batch(ctx._h, l.data_ptr(), r.data_ptr(), d.data_ptr(), m, N.GL_DEVICE)))
out["two_to_one_batch"] = {"pairs": m, "ms": t * 1e3, "perms_per_s": m / t}

nl, ll = 1 << 20, 135
leaves = rand_dev((nl, ll))
dig = torch.empty((2 * (nl - 16), 4), dtype=torch.int64, device=dev)
cap = torch.empty((16, 4), dtype=torch.int64, device=dev)
t = timeit(lambda: ctx.check(lib.gl_merkle_build(ctx._h, leaves.data_ptr(), nl, ll, 4, dig.data_ptr(), cap.data_ptr(), N.GL_DEVICE)), 3)
out["merkle_build_rows"] = {"leaves": nl, "leaf_len": ll, "ms": t * 1e3, "perms_per_s": (nl * 17 + nl - 16) / t}
hd = torch.empty((nl, 4), dtype=torch.int64, device=dev)
t = timeit(lambda: ctx.check(lib.gl_poseidon_hash_no_pad_batch(ctx._h, leaves.data_ptr(), ll, nl, hd.data_ptr(), N.GL_DEVICE)), 3)
out["hash_no_pad_rows"] = {"rows": nl, "len": ll, "ms": t * 1e3, "perms_per_s": nl * 17 / t}
del leaves, dig

# BASELINE config 4 (i): SparseMerkleProcessProof::check over a batch of 2^20 proofs.  The batch is made of VALID insert
# proofs of the shape a tree of 2^20 entries produces: 20 non-zero siblings, the new leaf goes into an empty slot
# (is_old0).  Their roots are computed here with the library's own batch hashes, level by level over the whole batch
# (no tree and no CPU hashing needed).  The reference's verifier hashes 2 x 256 levels + leaf hashes = 516
# permutations per proof; 474 of them are discarded (levels in state Na, old side below the insertion level) and
# the kernel does not compute those: 42 permutations per proof here, same statuses.
rng = np.random.default_rng(4)
mm, NS = 1 << 20, 20
new_key, new_val = rand_dev((mm, 4)), rand_dev((mm, 4))
sib = rand_dev((mm, NS, 4))
prev_new = torch.empty((mm, 4), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
ctx.check(lib.gl_smt_leaf_hash_batch(ctx._h, new_key.data_ptr(), new_val.data_ptr(), prev_new.data_ptr(), mm, N.GL_DEVICE))
prev_old = torch.zeros((mm, 4), dtype=torch.int64, device=dev)
nxt = torch.empty((mm, 4), dtype=torch.int64, device=dev)
for i in range(NS - 1, -1, -1):
    pos = ((new_key[:, i >> 6] >> (i & 63)) & 1).bool()[:, None]
    s_i = sib[:, i, :].contiguous()
    for prev in (prev_old, prev_new):
        l, r = torch.where(pos, s_i, prev).contiguous(), torch.where(pos, prev, s_i).contiguous()
        torch.cuda.synchronize()   # torch's stream and the library's are different streams
        ctx.check(lib.gl_poseidon_two_to_one_batch(ctx._h, l.data_ptr(), r.data_ptr(), nxt.data_ptr(), mm, N.GL_DEVICE))
        prev.copy_(nxt)
hdr_all = np.zeros(mm, dtype=glb.host.SMT_HDR_DTYPE)
hdr_all["old_root"] = prev_old.cpu().numpy().view(np.uint64)
hdr_all["new_root"] = prev_new.cpu().numpy().view(np.uint64)
hdr_all["new_key"] = (new_key.cpu().numpy().view(np.uint64)) % np.uint64(P)
hdr_all["new_value"] = (new_val.cpu().numpy().view(np.uint64)) % np.uint64(P)
hdr_all["is_old0"] = 1
hdr_all["fnc"] = 2
off_all = (np.arange(mm + 1, dtype=np.uint64) * np.uint64(NS))
d_hdr = torch.from_numpy(hdr_all.view(np.uint8)).to(dev)
d_pool = sib.reshape(mm * NS, 4)
d_off = torch.from_numpy(off_all.view(np.int64)).to(dev)
d_status = torch.empty(mm, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
t = timeit(lambda: ctx.check(lib.gl_smt_verify_process_batch(ctx._h, d_hdr.data_ptr(), d_pool.data_ptr(), d_off.data_ptr(), mm,
                                                            d_status.data_ptr(), N.GL_DEVICE)), 3)
assert int(d_status.abs().sum().item()) == 0, "synthetic insert proofs must verify: %s" % torch.unique(d_status, return_counts=True).__repr__()
# one flipped bit in a sibling must be caught
d_pool[12345, 1] ^= 1
torch.cuda.synchronize()
ctx.check(lib.gl_smt_verify_process_batch(ctx._h, d_hdr.data_ptr(), d_pool.data_ptr(), d_off.data_ptr(), mm, d_status.data_ptr(), N.GL_DEVICE))
assert int((d_status != 0).sum().item()) == 1 and int(d_status[12345 // NS].item()) != 0
out["smt_verify_process_batch"] = {"proofs": mm, "kind": "valid ProcessInsert proofs, %d siblings, is_old0" % NS, "ms": t * 1e3,
                                   "proofs_per_s": mm / t, "permutations_computed_per_proof": 2 * NS + 2,
                                   "perms_per_s": mm * (2 * NS + 2) / t, "reference_permutations_per_proof": 516}
del d_pool, d_hdr, hdr_all, sib, new_key, new_val, prev_old, prev_new, nxt

# N2: bulk build of the sparse Merkle tree over 2^20 entries (BASELINE config 4: "batch of 2^20 native leaf updates")
mk = 1 << 20
kk = rng.integers(0, P, (mk, 4), dtype=np.uint64)
vv = rng.integers(1, P, (mk, 4), dtype=np.uint64)
t = timeit(lambda: glb.host.smt_build_tree(kk, vv), 2)
out["smt_build_tree"] = {"entries": mk, "ms": t * 1e3, "entries_per_s": mk / t, "note": "host buffers: 64 MB H2D inside the call"}
# N2, second half: the 2^20 process proofs of those inserts (in call order, from an empty tree), device-resident
# outputs, then the batch verifier over what was emitted: BASELINE config 4 as a pipeline
dk, dvv = torch.from_numpy(kk.view(np.int64)).to(dev), torch.from_numpy(vv.view(np.int64)).to(dev)
d_hdr2 = torch.empty(mk * glb.host.SMT_HDR_DTYPE.itemsize, dtype=torch.uint8, device=dev)
cap2 = 24 * mk
d_pool2 = torch.empty((cap2, 4), dtype=torch.int64, device=dev)
d_off2 = torch.empty(mk + 1, dtype=torch.int64, device=dev)
tot = C.c_uint64(0)
torch.cuda.synchronize()
t = timeit(lambda: ctx.check(lib.gl_smt_set_proofs(ctx._h, dk.data_ptr(), dvv.data_ptr(), mk, d_hdr2.data_ptr(), d_pool2.data_ptr(), cap2,
                                                     d_off2.data_ptr(), C.byref(tot), N.GL_DEVICE)), 2)
assert tot.value <= cap2
d_st2 = torch.empty(mk, dtype=torch.int32, device=dev)
tv = timeit(lambda: ctx.check(lib.gl_smt_verify_process_batch(ctx._h, d_hdr2.data_ptr(), d_pool2.data_ptr(), d_off2.data_ptr(), mk,
                                                             d_st2.data_ptr(), N.GL_DEVICE)), 2)
assert int(d_st2.abs().sum().item()) == 0, "emitted proofs must verify"
out["smt_set_proofs"] = {"entries": mk, "ms": t * 1e3, "proofs_per_s": mk / t, "siblings_total": int(tot.value),
                            "avg_siblings": tot.value / mk, "verify_emitted_ms": tv * 1e3,
                            "note": "device-resident inputs and outputs; one permutation per (key, depth above its stopping point); the time order of a depth is merged from its children by the same binary search the hash needs"}
# config 4's "256-proof membership circuit": tree.find for 256 stored keys against the 2^20-entry tree (the witnesses of
# the inclusion circuits), tree build included
nq = 256
dq = dk[:: mk // nq].contiguous()
d_inc = torch.empty(nq * glb.host.SMT_INCLUSION_DTYPE.itemsize, dtype=torch.uint8, device=dev)
d_poolq = torch.empty((64 * nq, 4), dtype=torch.int64, device=dev)
d_offq = torch.empty(nq + 1, dtype=torch.int64, device=dev)
totq = C.c_uint64(0)
torch.cuda.synchronize()
tq = timeit(lambda: ctx.check(lib.gl_smt_find_batch(ctx._h, dk.data_ptr(), dvv.data_ptr(), mk, dq.data_ptr(), nq, d_inc.data_ptr(), d_poolq.data_ptr(),
                                                    64 * nq, d_offq.data_ptr(), C.byref(totq), N.GL_DEVICE)), 2)
inc = np.frombuffer(d_inc.cpu().numpy().tobytes(), dtype=glb.host.SMT_INCLUSION_DTYPE)
assert inc["found"].all() and totq.value <= 64 * nq
out["smt_find_batch"] = {"entries": mk, "queries": nq, "ms": tq * 1e3, "avg_siblings": totq.value / nq,
                         "note": "sort + versioned sweep over 2^20 sets and 256 queries; every query found"}
del dq, d_inc, d_poolq, d_offq

# the same number of `set` calls as account updates: 2^19 inserts, then 2^19 updates / removals of random existing keys
half = mk // 2
pick = torch.randint(0, half, (half,), device=dev)
dk2 = torch.cat([dk[:half], dk[:half][pick]]).contiguous()
dv2 = torch.cat([dvv[:half], dvv[half:]]).contiguous()
dv2[half + (half * 3) // 4:] = 0                           # the last eighth of the calls remove their key (or do nothing)
torch.cuda.synchronize()
t2 = timeit(lambda: ctx.check(lib.gl_smt_set_proofs(ctx._h, dk2.data_ptr(), dv2.data_ptr(), mk, d_hdr2.data_ptr(), d_pool2.data_ptr(), cap2,
                                                    d_off2.data_ptr(), C.byref(tot), N.GL_DEVICE)), 2)
assert tot.value <= cap2
ctx.check(lib.gl_smt_verify_process_batch(ctx._h, d_hdr2.data_ptr(), d_pool2.data_ptr(), d_off2.data_ptr(), mk, d_st2.data_ptr(), N.GL_DEVICE))
assert int(d_st2.abs().sum().item()) == 0, "emitted proofs must verify"
out["smt_set_proofs_mixed"] = {"calls": mk, "what": "2^19 inserts, then 2^19 calls on random existing keys: 3/4 updates, 1/4 removals or no-ops",
                               "ms": t2 * 1e3, "proofs_per_s": mk / t2, "siblings_total": int(tot.value)}
del kk, vv, dk, dvv, dk2, dv2, d_hdr2, d_pool2, d_off2

# verifier's side: 2^18 openings of a 2^20-leaf tree over 135-element rows checked against the cap in one call
lgv, kv = 20, 1 << 18
vrows = rand_dev((1 << lgv, 135))
vdig = torch.empty((2 * ((1 << lgv) - 16), 4), dtype=torch.int64, device=dev)
vcap = torch.empty((16, 4), dtype=torch.int64, device=dev)
torch.cuda.synchronize()
ctx.check(lib.gl_merkle_build(ctx._h, vrows.data_ptr(), 1 << lgv, 135, 4, vdig.data_ptr(), vcap.data_ptr(), N.GL_DEVICE))
vidx = torch.randint(0, 1 << lgv, (kv,), dtype=torch.int64, device=dev)
# sibling paths gathered from the digest buffer with torch (plonky2's in-order layout: MerkleTree::prove)
L = lgv - 4
sub = vidx >> L
pair = vidx & ((1 << L) - 1)
per_sub = 2 * ((1 << L) - 1)
vpaths = torch.empty((kv, L, 4), dtype=torch.int64, device=dev)
pp = pair.clone()
for i in range(L):
    parity = pp & 1
    pp = pp >> 1
    sib = 2 * ((pp << (i + 1)) + (1 << i) - 1) + (1 - parity)
    vpaths[:, i, :] = vdig[sub * per_sub + sib]
vleaf = vrows[vidx].contiguous()
vok = torch.empty(kv, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
t = timeit(lambda: ctx.check(lib.gl_merkle_verify_batch(ctx._h, vleaf.data_ptr(), 135, vidx.data_ptr(), vpaths.data_ptr(), L, vcap.data_ptr(), 4,
                                                      kv, vok.data_ptr(), N.GL_DEVICE)), 3)
assert int(vok.sum().item()) == kv, "every opening must verify"
out["merkle_verify_batch"] = {"openings": kv, "leaf_len": 135, "path_len": L, "ms": t * 1e3, "openings_per_s": kv / t,
                              "perms_per_s": kv * (17 + L) / t}
del vrows, vdig, vpaths, vleaf

# FRI: first reduction layer of a 2^20-row proof (N = 2^23 extension values, arity 16)
ln = 1 << 23
vals = rand_dev((ln, 2))
fd = torch.empty((2 * ((ln >> 4) - 16), 4), dtype=torch.int64, device=dev)
t = timeit(lambda: ctx.check(lib.gl_fri_layer_tree(ctx._h, vals.data_ptr(), ln, 4, 4, fd.data_ptr(), cap.data_ptr(), N.GL_DEVICE)), 3)
out["fri_layer_tree"] = {"ext_values": ln, "arity_bits": 4, "ms": t * 1e3}
fo = torch.empty((ln >> 4, 2), dtype=torch.int64, device=dev)
nx = torch.empty((ln >> 4, 2), dtype=torch.int64, device=dev)
beta = (C.c_uint64 * 2)(12345678901234567, 7654321987654321)
t = timeit(lambda: ctx.check(lib.gl_fri_fold(ctx._h, vals.data_ptr(), ln, 4, beta, pow(7, 16, P), fo.data_ptr(), nx.data_ptr(), N.GL_DEVICE)), 3)
out["fri_fold"] = {"ext_coeffs": ln, "ms": t * 1e3}
state = (C.c_uint64 * 12)(*[int(x) for x in rng.integers(0, P, 12, dtype=np.uint64)])
w = C.c_uint64()
t = timeit(lambda: ctx.check(lib.gl_pow_grind(ctx._h, state, 0, 16, C.byref(w))), 3)
out["pow_grind_16_bits"] = {"ms": t * 1e3, "witness": w.value}
# N1: PolynomialBatch::prove_openings at the shapes of a 2^20-row proof (constants+sigmas 84, wires 135, Z 20, quotient 16)
del vals, fd, fo, nx
ctx.trim()
fri = importlib.import_module("plonky2-lib_b200.fri")
lg = 20
batches = []
for k, c in enumerate((84, 135, 20, 16)):
    v = rand_dev((c, 1 << lg)).cpu().numpy().view(np.uint64) % np.uint64(P)
    batches.append(glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=False))
    del v
zeta = (1234567890123, 987654321987)
g = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - lg), P)
cols = (84, 135, 20, 16)
instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
            ((zeta[0] * g % P, zeta[1] * g % P), [(2, pi) for pi in range(20)])]
params = fri.FriParams.for_degree(glb.FriConfig(), lg)
t = timeit(lambda: [b.eval_at(zeta) for b in batches], 3)
out["opening_set_2^20"] = {"polynomials": sum(cols), "ms": t * 1e3,
                           "note": "OpeningSet::new at zeta for the 4 oracles (gl_commit_eval): 255 polynomials of 2^20 coefficients, results to the host"}


def run_open():
    ch = fri.Challenger()
    for b in batches:
        ch.observe_cap(b.merkle_tree.cap)
    return fri.prove_openings(batches, instance, ch, params)


def run_open_one_call():
    ch = fri.Challenger()
    for b in batches:
        ch.observe_cap(b.merkle_tree.cap)
    return fri.prove_openings_device(batches, instance, ch, params)


t = timeit(run_open_one_call, 3)
out["prove_openings_2^20_one_call"] = {"oracles": list(cols), "ms": t * 1e3, "note": "gl_fri_prove: the same proof through one C-ABI call, transcript on the device"}
t0 = time.perf_counter()
alpha = (111, 222)
fc, fv = fri.fri_final_poly(batches, instance, alpha, 3)
t_final = time.perf_counter() - t0
t = timeit(run_open, 2)
out["prove_openings_2^20"] = {"oracles": list(cols), "ms": t * 1e3, "fri_final_poly_ms_incl_d2h": t_final * 1e3,
                              "note": "alpha reduction of 275 polynomial openings, division, LDE, 4 FRI layers, 16-bit PoW, 28 queries; "
                                      "the FRI polynomial and every layer stay in HBM; only caps, the final coefficients and the 28 opened rows + paths cross PCIe (host Challenger)"}
# "Plonky2 prove ms per circuit" (BASELINE.json metric, first half) as far as this path goes: the commit + opening
# trace of one data.prove(pw) -- wires, Z/partial products, quotient commits and prove_openings over those three plus
# the build-time constants+sigmas oracle -- at the row counts SURVEY 8d estimates for BASELINE configs 1, 3, 4(ii), 5.
# Witness generation and compute_quotient_polys stay on the host in the reference and are not part of this number.
for b in batches:
    b.free()
ctx.trim()


def quotient_bench(lg):
    """gl_quotient_polys at standard_recursion_config geometry with the reference's three custom gates (+ Noop / Constant /
    PublicInput): data independent cost, so random oracles are as good as a satisfying witness."""
    n_ = 1 << lg
    circuit = (lg, 135, 80, 4, 2, 2, 8, 6)
    gates = [(0, 0, 0, 0, 3), (1, 2, 0, 0, 3), (2, 0, 0, 0, 3), (3, 3, 1, 3, 6), (4, 2, 1, 3, 6), (5, 2, 1, 3, 6)]
    k_is = np.array([pow(7, j, P) for j in range(80)], dtype=np.uint64)
    bs = [glb.PolynomialBatch.from_values(rand_dev((c, n_)), 3, False, 4, want_coeffs=False) for c in (84, 135, 20)]
    ch = np.array([3, 5], dtype=np.uint64)
    pih = np.arange(4, dtype=np.uint64)
    t_ = timeit(lambda: glb.host.compute_quotient_polys(circuit, gates, k_is, *bs, pih, ch, ch + np.uint64(9), ch + np.uint64(77)), 3)
    for b in bs:
        b.free()
    ctx.trim()
    rows = n_ * 8
    return {"rows": n_, "points": rows, "ms": t_ * 1e3, "lde_bytes_read": rows * (84 + 135 + 20 + 2) * 8,
            "gate_constraints_per_point": 3 * 34 + 2 * 67 + 2 * 67 + 2 + 4, "what": "kernel + coset_ifft of 2 x 8n values + 16 n coefficients D2H"}


out["quotient_polys"] = [quotient_bench(lg) for lg in (16, 18, 20)]


def proof_trace(lg, one_call=True):
    n_ = 1 << lg
    vals_ = [rand_dev((c, n_)).cpu().numpy().view(np.uint64) % np.uint64(P) for c in (84, 135, 20, 16)]
    const_sigmas = glb.PolynomialBatch.from_values(vals_[0], 3, False, 4, want_coeffs=False)   # CircuitBuilder::build
    g_ = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - lg), P)
    inst = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
            ((zeta[0] * g_ % P, zeta[1] * g_ % P), [(2, pi) for pi in range(20)])]
    prm = fri.FriParams.for_degree(glb.FriConfig(), lg)

    def once():
        bs = [const_sigmas] + [glb.PolynomialBatch.from_values(v, 3, False, 4, want_coeffs=True) for v in vals_[1:]]
        ch = fri.Challenger()
        for b in bs:
            ch.observe_cap(b.merkle_tree.cap)
        (fri.prove_openings_device if one_call else fri.prove_openings)(bs, inst, ch, prm)
        for b in bs[1:]:
            b.free()

    t_ = timeit(once, 3)
    const_sigmas.free()
    return t_ * 1e3


out["proof_trace_multi_call_ms"] = {"what": "the same trace with the Python-driven prover (host Challenger, ~70 dependent round trips)",
                                    "config1_ecdsa_2^16_rows": proof_trace(16, False), "config5_outer_recursion_2^13_rows": proof_trace(13, False)}
out["proof_trace_ms"] = {
    "what": "3 commits (135 + 20 + 16 columns, pageable host buffers in and coefficients out) + prove_openings over 4 oracles (gl_fri_prove: one call, transcript on the device)",
    "config1_ecdsa_2^16_rows": proof_trace(16),
    "config4_smt_256_inclusions_2^15_rows": proof_trace(15),
    "config3_keccak_64_blocks_2^18_rows": proof_trace(18),
    "config5_outer_recursion_2^13_rows": proof_trace(13),
}
print(json.dumps(out))
