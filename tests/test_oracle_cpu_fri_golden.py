"""tests/golden/fri_proof.json (make_golden.py) replayed by the oracle on CPU: the fixture is what the oracle produces."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_fri_proof_reproduced_by_the_oracle(oracle):
    from oracle import fri_oracle as fo

    g = json.load(open(os.path.join(GOLDEN, "fri_proof.json")))
    un = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)  # noqa: E731
    polys, trees = [], []
    for v in g["values"]:
        vals = np.stack([un(col) for col in v])
        res = oracle.commit_from_values(vals, g["rate_bits"], g["cap_height"])
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], g["cap_height"]))
    for t, cap in zip(trees, g["caps"]):
        assert np.array_equal(t.cap.reshape(-1), un(cap))
    instance = [(tuple(pt), [tuple(p) for p in ps]) for pt, ps in g["instance"]]
    assert [[[int(x) for x in ov] for ov in b] for b in fo.opening_set(polys, instance)] == g["openings"]
    ch = fo.Challenger()
    for t in trees:
        ch.observe_cap(t.cap)
    proof = fo.prove_openings(polys, trees, instance, ch, g["degree_bits"], g["rate_bits"], g["cap_height"], g["proof_of_work_bits"],
                              g["num_query_rounds"])
    flat = []
    for cap in proof["commit_phase_merkle_caps"]:
        flat += [int(x) for x in np.asarray(cap).reshape(-1)]
    flat += [int(x) for x in np.asarray(proof["final_poly"]).reshape(-1)]
    flat.append(int(proof["pow_witness"]))
    for r in proof["query_round_proofs"]:
        flat.append(int(r["x_index"]))
        for row, path in r["initial_trees_proof"]:
            flat += [int(x) for x in np.asarray(row).reshape(-1)] + [int(x) for x in np.asarray(path).reshape(-1)]
        for st in r["steps"]:
            flat += [int(x) for x in np.asarray(st["evals"]).reshape(-1)] + [int(x) for x in np.asarray(st["merkle_proof"]).reshape(-1)]
    assert [f"{x:x}" for x in flat] == g["proof_flat"]
