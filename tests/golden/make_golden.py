"""Generates the fixtures in tests/golden/.

poseidon_kat.json : the permutation vectors of SURVEY.md 8c.  The all-zero vector's first four lanes are
                    the reference's own KAT (src/zkdsa/circuits/mod.rs:85-105); the others were PROBED with
                    the constants recipe and match upstream plonky2's poseidon_goldilocks test_vectors.
commit_caps.json  : Merkle caps of PolynomialBatch::from_values on the synthetic input of SURVEY 8d,
                    produced by the oracle itself ("parity unpinned" rows: regression fixtures only).
smt_sets.json     : a fixed sequence of 40 `tree.set` calls on 10 keys (inserts, updates, removals, no-ops, keys that come
                    back; (1,2), (12,1), (5,51) of src/smt/gadgets/verify/mod.rs:24-34 first) with the process proof of
                    every call and `find` results against the final tree, produced by the oracle's restatement of
                    src/smt/tree.rs ("parity unpinned": regression fixture).
fri_proof.json    : prove_openings of two small oracles (2^5 rows; 3 + 2 polynomials; 8-bit PoW; 28 queries) by the oracle's
                    restatement of plonky2::fri::prover, flattened field by field ("parity unpinned": regression fixture).
Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
P = o.P

perm = []
for name, vec in [("zeros", [0] * 12), ("iota", list(range(12))), ("p_minus_1", [P - 1] * 12)]:
    out = o.permute(np.array(vec, dtype=np.uint64))
    perm.append({"name": name, "in": [f"{x:016x}" for x in vec], "out": [f"{int(x):016x}" for x in out]})
expected = {
    "zeros": "3c18a9786cb0b359 c4055e3364a246c3 7953db0ab48808f4 c71603f33a1144ca d7709673896996dc 46a84e87642f44ed d032648251ee0b3c 1c687363b207df62 df8565563e8045fe 40f5b37ff4254dae d070f637b431067c 1792b1c4342109d7",
    "iota": "d64e1e3efc5b8e9e 53666633020aaa47 d40285597c6a8825 613a4f81e81231d2 414754bfebd051f0 cb1f8980294a023f 6eb2a9e4d54a9d0f 1902bc3af467e056 f045d5eafdc6021f e4150f77caaa3be5 c9bfd01d39b50cce 5c0a27fcb0e1459b",
    "p_minus_1": "be0085cfc57a8357 d95af71847d05c09 cf55a13d33c1c953 95803a74f4530e82 fcd99eb30a135df1 e095905e913a3029 de0392461b42919b 7d3260e24e81d031 10d3d0465d9deaa0 a87571083dfc2a47 e18263681e9958f8 e28e96f1ae5e60d3",
}
for p_ in perm:
    assert p_["out"] == expected[p_["name"]].split(), p_["name"]
json.dump({"source": "SURVEY.md 8c; zeros[0:4] = src/zkdsa/circuits/mod.rs:85-105", "permutation": perm},
          open(os.path.join(HERE, "poseidon_kat.json"), "w"), indent=1)

cases = []
for lg_n, c, r, h in [(3, 5, 3, 4), (6, 20, 3, 4), (8, 135, 3, 4), (10, 16, 3, 4), (12, 136, 3, 4), (14, 135, 3, 4)]:
    v = o.synthetic_values(c, 1 << lg_n)
    res = o.commit_from_values(v, r, h, want_leaves=False)
    cases.append({"lg_n": lg_n, "c": c, "rate_bits": r, "cap_height": h,
                  "cap": [f"{int(x):016x}" for x in res["cap"].reshape(-1)]})
json.dump({"source": "oracle/gl_oracle.c glo_commit_from_values on pyoracle.synthetic_values (seed 0x706C6F6E6B7932)",
           "cases": cases}, open(os.path.join(HERE, "commit_caps.json"), "w"), indent=1)


def hx(a):
    return [f"{int(x):016x}" for x in np.asarray(a).reshape(-1)]


rng = np.random.default_rng(0x736D74)
pool = [o.from_u128(k) for k in (1, 12, 5)] + [rng.integers(0, P, 4, dtype=np.uint64) for _ in range(7)]
twin = pool[3].copy()
twin[2] ^= np.uint64(1) << np.uint64(9)            # shares 137 path bits with pool[3]
pool[9] = twin
seq_k = [pool[0], pool[1], pool[2]] + [pool[int(i)] for i in rng.integers(0, 10, 37)]
seq_v = [o.from_u128(2), o.from_u128(1), o.from_u128(51)]
for i in range(37):
    seq_v.append(np.zeros(4, dtype=np.uint64) if rng.random() < 0.3 else rng.integers(1, P, 4, dtype=np.uint64))
tree = o.Smt()
calls = []
for k, v in zip(seq_k, seq_v):
    r = tree.set(k, v)
    ns = int(r["num_siblings"])
    calls.append({"key": hx(k), "value": hx(v), "fnc": int(r["fnc"]), "is_old0": int(r["is_old0"]), "old_root": hx(r["old_root"]),
                  "new_root": hx(r["new_root"]), "old_key": hx(r["old_key"]), "old_value": hx(r["old_value"]), "new_key": hx(r["new_key"]),
                  "new_value": hx(r["new_value"]), "siblings": hx(r["siblings"][:ns])})
finds = []
for q in pool + [rng.integers(0, P, 4, dtype=np.uint64) for _ in range(4)]:
    f = tree.find(q)
    finds.append({"key": hx(q), "found": bool(f["found"]), "is_old0": bool(f["is_old0"]), "value": hx(f["value"]),
                  "not_found_key": hx(f["not_found_key"]), "siblings": hx(f["siblings"])})
assert {c["fnc"] for c in calls} == {0, 1, 2, 3}
json.dump({"source": "oracle/gl_oracle.c glo_smt_set / glo_smt_find (restatement of src/smt/tree.rs); calls[0:3] = src/smt/gadgets/verify/mod.rs:24-34",
           "root": hx(tree.root()), "calls": calls, "finds": finds}, open(os.path.join(HERE, "smt_sets.json"), "w"), indent=1)
from oracle import fri_oracle as fo  # noqa: E402


def flatten_proof(proof):
    out = []
    for cap in proof["commit_phase_merkle_caps"]:
        out += [int(x) for x in np.asarray(cap).reshape(-1)]
    out += [int(x) for x in np.asarray(proof["final_poly"]).reshape(-1)]
    out.append(int(proof["pow_witness"]))
    for r in proof["query_round_proofs"]:
        out.append(int(r["x_index"]))
        for row, path in r["initial_trees_proof"]:
            out += [int(x) for x in np.asarray(row).reshape(-1)] + [int(x) for x in np.asarray(path).reshape(-1)]
        for st in r["steps"]:
            out += [int(x) for x in np.asarray(st["evals"]).reshape(-1)] + [int(x) for x in np.asarray(st["merkle_proof"]).reshape(-1)]
    return out


DEG, RATE, CAPH, POW, ROUNDS = 5, 3, 4, 8, 28
cols = (3, 2)
vals, polys, trees = [], [], []
for k, c in enumerate(cols):
    v = o.synthetic_values(c, 1 << DEG, seed=4242 + k)
    res = o.commit_from_values(v, RATE, CAPH)
    vals.append(v)
    polys.append(res["coeffs"])
    trees.append(fo.MerkleTree(res["leaves"], CAPH))
zeta = (0x1234567, 0x89ABCDE)
gen = o.lib().glo_primitive_root_of_unity(DEG)
instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]), (fo.ext_scalar(zeta, gen), [(1, 0), (1, 1)])]
ch = fo.Challenger()
for t_ in trees:
    ch.observe_cap(t_.cap)
proof = fo.prove_openings(polys, trees, instance, ch, DEG, RATE, CAPH, POW, ROUNDS)
json.dump({"source": "oracle/fri_oracle.py prove_openings (restatement of plonky2::fri::prover / oracle)",
           "degree_bits": DEG, "rate_bits": RATE, "cap_height": CAPH, "proof_of_work_bits": POW, "num_query_rounds": ROUNDS,
           "values": [[hx(col) for col in v] for v in vals],
           "instance": [[list(int(x) for x in pt), [list(p_) for p_ in ps]] for pt, ps in instance],
           "caps": [hx(t_.cap) for t_ in trees],
           "openings": [[[int(x) for x in ov] for ov in b] for b in fo.opening_set(polys, instance)],
           "proof_flat": [f"{x:x}" for x in flatten_proof(proof)]}, open(os.path.join(HERE, "fri_proof.json"), "w"), indent=0)
print("wrote", HERE)
