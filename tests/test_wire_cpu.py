"""plonky2's binary proof layout (plonky2-lib_b200/wire.py): writer and reader round-trip the golden FRI proof, the size
formula is exact, and malformed inputs are rejected instead of mis-parsed (host-side formatting, no device needed)."""
import importlib
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fri_proof.json")


def _golden_proof():
    fri = importlib.import_module("plonky2-lib_b200.fri")
    host = importlib.import_module("plonky2-lib_b200.host")
    g = json.load(open(GOLD))
    flat = np.array([int(x, 16) for x in g["proof_flat"]], dtype=np.uint64)
    cfg = host.FriConfig(rate_bits=g["rate_bits"], cap_height=g["cap_height"], proof_of_work_bits=g["proof_of_work_bits"],
                         num_query_rounds=g["num_query_rounds"])
    params = fri.FriParams.for_degree(cfg, g["degree_bits"])
    cols = [len(v) for v in g["values"]]
    return g, fri.parse_flat_proof(flat, cols, params), cols, params


def test_fri_proof_bytes_round_trip_and_layout():
    wire = importlib.import_module("plonky2-lib_b200.wire")
    g, proof, cols, params = _golden_proof()
    data = wire.fri_proof_to_bytes(proof)
    shape = (cols, g["degree_bits"], g["rate_bits"], g["cap_height"], list(params.reduction_arity_bits), g["num_query_rounds"])
    assert len(data) == wire.fri_proof_num_bytes(*shape)
    back = wire.fri_proof_from_bytes(data, *shape)
    assert back["pow_witness"] == proof["pow_witness"] and np.array_equal(back["final_poly"], proof["final_poly"])
    for a, b in zip(back["query_round_proofs"], proof["query_round_proofs"]):
        for (ra, pa), (rb, pb) in zip(a["initial_trees_proof"], b["initial_trees_proof"]):
            assert np.array_equal(ra, rb) and np.array_equal(pa, pb)
        for sa, sb in zip(a["steps"], b["steps"]):
            assert np.array_equal(sa["evals"], sb["evals"]) and np.array_equal(sa["merkle_proof"], sb["merkle_proof"])
    assert wire.fri_proof_to_bytes(back) == data
    # first bytes: the first cap's first digest, 4 x u64 little endian (HashOut::to_bytes, pinned by
    # src/zkdsa/circuits/mod.rs:149 + src/smt/goldilocks_poseidon/hash/mod.rs:84-95); last 8 bytes: pow_witness
    if proof["commit_phase_merkle_caps"]:
        first = proof["commit_phase_merkle_caps"][0][0]
        assert data[:32] == b"".join(int(x).to_bytes(8, "little") for x in first)
    assert data[-8:] == int(proof["pow_witness"]).to_bytes(8, "little")
    # a Merkle proof is length-prefixed with ONE byte
    lgN, h = g["degree_bits"] + g["rate_bits"], g["cap_height"]
    at = len(proof["commit_phase_merkle_caps"]) * (32 << h) + 8 * cols[0]
    assert data[at] == lgN - h
    with pytest.raises(wire.WireError):
        wire.fri_proof_from_bytes(data[:-1], *shape)
    with pytest.raises(wire.WireError):
        wire.fri_proof_from_bytes(data + b"\x00" * 8, *shape)
    bad = bytearray(data)
    bad[at] += 1                                               # a sibling count the circuit does not allow
    with pytest.raises(wire.WireError):
        wire.fri_proof_from_bytes(bytes(bad), *shape)
    bad = bytearray(data)
    bad[0:8] = (0xFFFFFFFF00000001).to_bytes(8, "little")      # p itself: not canonical
    with pytest.raises(wire.WireError):
        wire.fri_proof_from_bytes(bytes(bad), *shape)


def test_proof_with_public_inputs_round_trip():
    wire = importlib.import_module("plonky2-lib_b200.wire")
    g, proof, cols, params = _golden_proof()
    rng = np.random.default_rng(3)
    P = 0xFFFFFFFF00000001
    caps = [rng.integers(0, P, (1 << g["cap_height"], 4), dtype=np.uint64) for _ in range(3)]
    lens = {"constants": 4, "plonk_sigmas": 80, "wires": 135, "plonk_zs": 2, "plonk_zs_next": 2, "partial_products": 18, "quotient_polys": 16}
    openings = {k: rng.integers(0, P, (v, 2), dtype=np.uint64) for k, v in lens.items()}
    pis = rng.integers(0, P, 7, dtype=np.uint64)
    data = wire.proof_with_public_inputs_to_bytes(caps, openings, proof, pis)
    shape = (cols, g["degree_bits"], g["rate_bits"], g["cap_height"], list(params.reduction_arity_bits), g["num_query_rounds"])
    assert len(data) == 3 * (32 << g["cap_height"]) + 16 * sum(lens.values()) + wire.fri_proof_num_bytes(*shape) + 8 * 7
    c2, o2, f2, p2 = wire.proof_with_public_inputs_from_bytes(data, lens, *shape)
    assert all(np.array_equal(a, b) for a, b in zip(c2, caps)) and all(np.array_equal(o2[k], openings[k]) for k in lens)
    assert np.array_equal(p2, pis) and wire.fri_proof_to_bytes(f2) == wire.fri_proof_to_bytes(proof)
    with pytest.raises(wire.WireError):
        wire.proof_with_public_inputs_from_bytes(data + b"\x01", lens, *shape)
