"""N1: plonky2::fri::prover::fri_proof on the device (resident layer trees, device PoW, gathered
query openings, host Challenger) against the oracle's fri_proof, field by field, and through the oracle's
restatement of the upstream verifier."""
import importlib

import numpy as np
import pytest

from conftest import P, rand_field

pytestmark = pytest.mark.gpu


def _same_proof(a, b):
    assert a["pow_witness"] == b["pow_witness"]
    assert np.array_equal(a["final_poly"], b["final_poly"])
    assert len(a["commit_phase_merkle_caps"]) == len(b["commit_phase_merkle_caps"])
    for x, y in zip(a["commit_phase_merkle_caps"], b["commit_phase_merkle_caps"]):
        assert np.array_equal(x, y)
    assert len(a["query_round_proofs"]) == len(b["query_round_proofs"])
    for ra, rb in zip(a["query_round_proofs"], b["query_round_proofs"]):
        assert ra["x_index"] == rb["x_index"]
        for (rowa, patha), (rowb, pathb) in zip(ra["initial_trees_proof"], rb["initial_trees_proof"]):
            assert np.array_equal(rowa, rowb) and np.array_equal(patha, pathb)
        for sa, sb in zip(ra["steps"], rb["steps"]):
            assert np.array_equal(sa["evals"], sb["evals"]) and np.array_equal(sa["merkle_proof"], sb["merkle_proof"])


@pytest.mark.parametrize("degree_bits,pow_bits,cols", [(6, 4, (3,)), (10, 8, (5, 20)), (12, 16, (135, 20, 16))])
def test_fri_proof_matches_oracle_and_verifies(glb, ctx, oracle, rng, degree_bits, pow_bits, cols):
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    rate_bits, cap_height, rounds = 3, 4, 28
    n, lde = 1 << degree_bits, 1 << (degree_bits + rate_bits)
    coeffs = np.zeros((lde, 2), dtype=np.uint64)
    coeffs[:n] = rand_field(rng, (n, 2))
    values = glb.coset_fft(np.ascontiguousarray(coeffs.T), 7)          # extension coset_fft = two base transforms
    values = np.ascontiguousarray(values.T)
    assert np.array_equal(values, oracle.ext_coset_fft(coeffs, 7))
    # the initial oracles: resident commits on the device, mirrored trees on the CPU
    batches, cpu_trees = [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=900 + k)
        b = glb.PolynomialBatch.from_values(v, rate_bits, False, cap_height)
        batches.append(b)
        cpu_trees.append(fo.MerkleTree(oracle.commit_from_values(v, rate_bits, cap_height)["leaves"], cap_height))
        assert np.array_equal(b.merkle_tree.cap, cpu_trees[-1].cap)
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)
    assert list(params.reduction_arity_bits) == fo.reduction_arity_bits(degree_bits, rate_bits, cap_height)
    ch = fri.Challenger()
    och = fo.Challenger()
    for t in cpu_trees:
        ch.observe_cap(t.cap)
        och.observe_cap(t.cap)
    got = fri.fri_proof(batches, coeffs, values, ch, params)
    want = fo.fri_proof(cpu_trees, coeffs, values, och, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    _same_proof(got, want)
    assert ch.get_challenge() == och.get_challenge()                  # transcripts end in the same state
    vch = fo.Challenger()
    for t in cpu_trees:
        vch.observe_cap(t.cap)

    def first_layer_eval(x_index, rows, subgroup_x):
        return fo.eval_ext_poly(coeffs[:n], (subgroup_x, 0))

    assert fo.verify_fri_proof(got, [t.cap for t in cpu_trees], cap_height, vch, degree_bits, first_layer_eval,
                               rate_bits, cap_height, pow_bits, rounds)
    for b in batches:
        b.free()


def test_challenger_matches_oracle(glb, ctx, oracle, rng):
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    a, b = fri.Challenger(), fo.Challenger()
    for step in range(40):
        if step % 3 == 0:
            assert a.get_challenge() == b.get_challenge()
        else:
            xs = rand_field(rng, (int(rng.integers(1, 12)),))
            a.observe_elements(xs)
            b.observe_elements(xs)
    assert a.get_extension_challenge() == b.get_extension_challenge()


def test_challenger_chained_duplexing_matches_element_by_element(glb, ctx, oracle, rng):
    """observe_elements sends runs of full input buffers to the device as ONE gl_poseidon_duplex_chain call; the
    transcript (sponge state, pending inputs, buffered outputs) must be what upstream's element-by-element
    observation leaves, whatever the alignment of the run with the buffer."""
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    for pre in range(0, 9):                       # elements already waiting in input_buffer
        for length in (15, 16, 17, 64, 71, 256):  # a Merkle cap is 64
            a, b = fri.Challenger(), fo.Challenger()
            head = rand_field(rng, (pre,))
            for e in head.tolist():
                a.observe_element(e)
                b.observe_element(e)
            xs = rand_field(rng, (length,))
            xs[0] = np.uint64(P)                  # non-canonical input: observed as 0
            a.observe_elements(xs)
            for e in xs.tolist():
                b.observe_element(e)
            assert [int(v) for v in a.sponge_state] == [int(v) for v in b.sponge_state]
            assert list(a.input_buffer) == [int(v) for v in b.input_buffer]
            assert list(a.output_buffer) == [int(v) for v in b.output_buffer]
            assert [a.get_challenge() for _ in range(9)] == [b.get_challenge() for _ in range(9)]
    # the raw entry point: a zero-length chain is a no-op, NULL arguments are rejected
    st = np.arange(12, dtype=np.uint64)
    assert ctx._lib.gl_poseidon_duplex_chain(ctx._h, st.ctypes.data, None, 0) == 0 and st[5] == 5
    assert ctx._lib.gl_poseidon_duplex_chain(ctx._h, None, None, 1) == glb._native.GL_E_ARG


@pytest.mark.parametrize("degree_bits,cols", [(5, (3, 2)), (8, (4, 9, 3)), (12, (84, 135, 20, 16))])
def test_prove_openings_matches_oracle_and_verifies(glb, ctx, oracle, rng, degree_bits, cols):
    """PolynomialBatch::prove_openings end to end on the device (alpha reduction, division by X - z, LDE, FRI),
    against the oracle and through its restatement of the upstream verifier (fri_combine_initial included).
    The last case has the oracle shapes of a real proof: constants+sigmas, wires, Z/partial products, quotient."""
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    rate_bits, cap_height, pow_bits, rounds = 3, 4, 10, 28
    n = 1 << degree_bits
    batches_dev, polys, trees = [], [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=700 + k)
        b = glb.PolynomialBatch.from_values(v, rate_bits, False, cap_height)
        res = oracle.commit_from_values(v, rate_bits, cap_height)
        batches_dev.append(b)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], cap_height))
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    g = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    zs = min(len(cols) - 1, 2)                                    # the oracle holding the Z polynomials
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
                (fo.ext_scalar(zeta, g), [(zs, pi) for pi in range(min(2, cols[zs]))])]
    # the FRI polynomial itself
    alpha = tuple(int(x) for x in rand_field(rng, (2,)))
    want_final = fo.final_poly_of_openings(polys, instance, alpha)
    got_coeffs, got_values = fri.fri_final_poly(batches_dev, instance, alpha, rate_bits)
    assert np.array_equal(got_coeffs[:n], want_final) and not got_coeffs[n:].any()
    assert np.array_equal(got_values, oracle.ext_coset_fft(got_coeffs, 7))
    # the whole opening proof
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)
    ch, och, vch = fri.Challenger(), fo.Challenger(), fo.Challenger()
    for t in trees:
        for c in (ch, och, vch):
            c.observe_cap(t.cap)
    got = fri.prove_openings(batches_dev, instance, ch, params)
    want = fo.prove_openings(polys, trees, instance, och, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    _same_proof(got, want)
    openings = fo.opening_set(polys, instance)
    assert fo.verify_openings(got, openings, [t.cap for t in trees], instance, vch, degree_bits, rate_bits, cap_height,
                              pow_bits, rounds)
    for b in batches_dev:
        b.free()


@pytest.mark.parametrize("degree_bits,cols", [(5, (3, 2)), (9, (4, 9, 3)), (12, (84, 135, 20, 16))])
def test_device_verifier_accepts_proofs_and_rejects_tampering(glb, ctx, oracle, rng, degree_bits, cols):
    """plonky2::fri::verifier on the product side (fri_verifier.py: Merkle openings batched on the device, transcript on
    the product Challenger, interpolation in exact integers on the host): accepts the device prover's proof AND the
    oracle prover's proof, and rejects each kind of tampering where the oracle's restatement of the verifier rejects it."""
    import copy

    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    fv = importlib.import_module("plonky2-lib_b200.fri_verifier")
    rate_bits, cap_height, pow_bits, rounds = 3, 4, 8, 28
    n = 1 << degree_bits
    batches_dev, polys, trees = [], [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=300 + k)
        batches_dev.append(glb.PolynomialBatch.from_values(v, rate_bits, False, cap_height))
        res = oracle.commit_from_values(v, rate_bits, cap_height)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], cap_height))
    caps = [t.cap for t in trees]
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    g = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    zs = min(len(cols) - 1, 2)
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
                (fo.ext_scalar(zeta, g), [(zs, pi) for pi in range(min(2, cols[zs]))])]
    openings = fo.opening_set(polys, instance)
    assert fri.opening_set(batches_dev, instance) == [[tuple(int(x) for x in v) for v in b] for b in openings]   # OpeningSet::new on the device
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)

    def transcript():
        ch = fri.Challenger()
        for cap in caps:
            ch.observe_cap(cap)
        return ch

    proof = fri.prove_openings(batches_dev, instance, transcript(), params)
    assert fv.verify_openings(instance, openings, caps, proof, transcript(), params) is True
    # a proof made by the oracle's prover
    och = fo.Challenger()
    for cap in caps:
        och.observe_cap(cap)
    oproof = fo.prove_openings(polys, trees, instance, och, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    assert fv.verify_openings(instance, openings, caps, oproof, transcript(), params) is True

    def oracle_accepts(p, ops):
        vch = fo.Challenger()
        for cap in caps:
            vch.observe_cap(cap)
        try:
            return bool(fo.verify_openings(p, ops, caps, instance, vch, degree_bits, rate_bits, cap_height, pow_bits, rounds))
        except AssertionError:
            return False

    def tampered(kind):
        p, ops = copy.deepcopy(proof), copy.deepcopy(openings)
        r = p["query_round_proofs"][3]
        if kind == "pow":
            p["pow_witness"] = int(p["pow_witness"]) + 1
        elif kind == "final":
            p["final_poly"] = np.array(p["final_poly"], copy=True)
            p["final_poly"][0, 0] ^= np.uint64(1)
        elif kind == "eval" and r["steps"]:
            r["steps"][0]["evals"] = np.array(r["steps"][0]["evals"], copy=True)
            r["steps"][0]["evals"][1, 1] ^= np.uint64(2)
        elif kind == "sibling" and r["steps"] and len(r["steps"][0]["merkle_proof"]):
            r["steps"][0]["merkle_proof"] = np.array(r["steps"][0]["merkle_proof"], copy=True)
            r["steps"][0]["merkle_proof"][0, 3] ^= np.uint64(1)
        elif kind == "row":
            row, path = r["initial_trees_proof"][0]
            row = np.array(row, copy=True)
            row[0] ^= np.uint64(1)
            r["initial_trees_proof"][0] = (row, path)
        elif kind == "opening":
            ops[0][0] = ((int(ops[0][0][0]) + 1) % P, int(ops[0][0][1]))
        elif kind == "cap" and p["commit_phase_merkle_caps"]:
            p["commit_phase_merkle_caps"][0] = np.array(p["commit_phase_merkle_caps"][0], copy=True)
            p["commit_phase_merkle_caps"][0][0, 0] ^= np.uint64(1)
        else:
            return None
        return p, ops

    for kind in ("pow", "final", "eval", "sibling", "row", "opening", "cap"):
        t = tampered(kind)
        if t is None:
            continue
        p, ops = t
        assert not oracle_accepts(p, ops), kind
        with pytest.raises(fv.FriVerifyError):
            fv.verify_openings(instance, ops, caps, p, transcript(), params)
    for b in batches_dev:
        b.free()


@pytest.mark.parametrize("lg_n,c", [(0, 3), (3, 5), (8, 20), (12, 135), (13, 7), (14, 2)])
def test_commit_eval_matches_oracle(glb, ctx, oracle, rng, lg_n, c):
    """gl_commit_eval = OpeningSet::new: every polynomial of a resident commit at an extension-field point (one chunk,
    exactly one chunk, several chunks of 4096 coefficients)."""
    n = 1 << lg_n
    values = oracle.synthetic_values(c, n, seed=40 + lg_n)
    b = glb.PolynomialBatch.from_values(values, 3 if lg_n else 0, False, 0)
    coeffs = np.asarray(b.polynomials)
    for point in (tuple(int(x) for x in rand_field(rng, (2,))), (0, 0), (1, 0), (P - 1, P - 1), (5, 0)):
        got = b.eval_at(point)
        for j in (0, c // 2, c - 1):
            assert tuple(int(x) for x in got[j]) == oracle.eval_base_poly_at_ext(coeffs[j], np.array(point, dtype=np.uint64)), (point, j)
    b.free()


def test_golden_fri_proof(glb, ctx):
    """The committed fixture tests/golden/fri_proof.json against the device prover and the product verifier, without the
    oracle in the loop: commits, OpeningSet, every field of the opening proof."""
    import json
    import os

    fri = importlib.import_module("plonky2-lib_b200.fri")
    fv = importlib.import_module("plonky2-lib_b200.fri_verifier")
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fri_proof.json")))
    un = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)  # noqa: E731
    batches = [glb.PolynomialBatch.from_values(np.stack([un(col) for col in v]), g["rate_bits"], False, g["cap_height"]) for v in g["values"]]
    caps = [b.merkle_tree.cap for b in batches]
    for cap, want in zip(caps, g["caps"]):
        assert np.array_equal(cap.reshape(-1), un(want))
    instance = [(tuple(pt), [tuple(p) for p in ps]) for pt, ps in g["instance"]]
    openings = fri.opening_set(batches, instance)
    assert [[list(ov) for ov in b] for b in openings] == g["openings"]
    cfg = glb.FriConfig(rate_bits=g["rate_bits"], cap_height=g["cap_height"], proof_of_work_bits=g["proof_of_work_bits"],
                        num_query_rounds=g["num_query_rounds"])
    params = fri.FriParams.for_degree(cfg, g["degree_bits"])

    def transcript():
        ch = fri.Challenger()
        for cap in caps:
            ch.observe_cap(cap)
        return ch

    proof = fri.prove_openings(batches, instance, transcript(), params)
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
    assert fv.verify_openings(instance, openings, caps, proof, transcript(), params) is True
    for b in batches:
        b.free()


@pytest.mark.parametrize("degree_bits,cols,pow_bits,times_x", [(5, (3, 2), 6, False), (8, (4, 9, 3), 10, False), (8, (4, 9, 3), 10, True),
                                                                (12, (84, 135, 20, 16), 16, False), (13, (84, 135, 20, 16), 16, False)])
def test_one_call_prover_matches_oracle(glb, ctx, oracle, rng, degree_bits, cols, pow_bits, times_x):
    """gl_fri_prove: prove_openings + fri_proof in ONE C-ABI call with the transcript on the device, against the oracle's
    prover field by field (both fork forms), through the oracle's verifier and the product's, and with the Challenger
    coming back in the state the oracle's ends in."""
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    fv = importlib.import_module("plonky2-lib_b200.fri_verifier")
    rate_bits, cap_height, rounds = 3, 4, 28
    n = 1 << degree_bits
    batches_dev, polys, trees = [], [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=300 + k)
        batches_dev.append(glb.PolynomialBatch.from_values(v, rate_bits, False, cap_height))
        res = oracle.commit_from_values(v, rate_bits, cap_height)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], cap_height))
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    g = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    zs = min(len(cols) - 1, 2)
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]),
                (fo.ext_scalar(zeta, g), [(zs, pi) for pi in range(min(2, cols[zs]))])]
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)
    params.final_poly_times_x = times_x
    ch, och, vch, pch = fri.Challenger(), fo.Challenger(), fo.Challenger(), fri.Challenger()
    for t in trees:
        for c in (ch, och, vch, pch):
            c.observe_cap(t.cap)
    ch.observe_elements([1, 2, 3])             # a ragged input buffer when the call starts
    och.observe_elements([1, 2, 3])
    vch.observe_elements([1, 2, 3])
    pch.observe_elements([1, 2, 3])
    got = fri.prove_openings_device(batches_dev, instance, ch, params)
    want = fo.prove_openings(polys, trees, instance, och, degree_bits, rate_bits, cap_height, pow_bits, rounds, times_x=times_x)
    _same_proof(got, want)
    assert ch.get_challenge() == och.get_challenge()                  # the transcript continues from the same state
    openings = fo.opening_set(polys, instance)
    assert fo.verify_openings(got, openings, [t.cap for t in trees], instance, vch, degree_bits, rate_bits, cap_height, pow_bits,
                              rounds, times_x=times_x)
    assert fv.verify_openings(instance, openings, [t.cap for t in trees], got, pch, params)
    # the Python-driven multi-call prover gives the same proof
    ch2 = fri.Challenger()
    for t in trees:
        ch2.observe_cap(t.cap)
    ch2.observe_elements([1, 2, 3])
    _same_proof(fri.prove_openings(batches_dev, instance, ch2, params), want)
    for b in batches_dev:
        b.free()


def test_one_call_prover_with_blinded_oracles(glb, ctx, oracle, rng):
    """Salted (blinding = true) oracles through gl_fri_prove and the multi-call prover: the opened rows carry the SALT_SIZE
    salt elements after the polynomial values, the proof equals the oracle prover's over the same salted trees, and both
    verifiers (which index the evaluations by polynomial, i.e. ignore the salt as upstream's unsalted_evals does) accept."""
    from oracle import fri_oracle as fo

    fri = importlib.import_module("plonky2-lib_b200.fri")
    fv = importlib.import_module("plonky2-lib_b200.fri_verifier")
    degree_bits, cols, pow_bits, rate_bits, cap_height, rounds = 8, (6, 9, 3), 8, 3, 4, 28
    n = 1 << degree_bits
    ctx.set_salt_seed(99)
    batches_dev, polys, trees = [], [], []
    for k, c in enumerate(cols):
        v = oracle.synthetic_values(c, n, seed=500 + k)
        b = glb.PolynomialBatch.from_values(v, rate_bits, k != 1, cap_height)      # oracles 0 and 2 salted, 1 plain
        batches_dev.append(b)
        polys.append(oracle.commit_from_values(v, rate_bits, cap_height)["coeffs"])
        trees.append(fo.MerkleTree(b.merkle_tree.leaves, cap_height))
        assert np.array_equal(trees[-1].cap, b.merkle_tree.cap)
        assert b.leaf_len == c + (4 if k != 1 else 0)
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)])]
    cfg = glb.FriConfig(rate_bits=rate_bits, cap_height=cap_height, proof_of_work_bits=pow_bits, num_query_rounds=rounds)
    params = fri.FriParams.for_degree(cfg, degree_bits)
    ch, och, vch, pch, ch2 = fri.Challenger(), fo.Challenger(), fo.Challenger(), fri.Challenger(), fri.Challenger()
    for t in trees:
        for c in (ch, och, vch, pch, ch2):
            c.observe_cap(t.cap)
    got = fri.prove_openings_device(batches_dev, instance, ch, params)
    want = fo.prove_openings(polys, trees, instance, och, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    _same_proof(got, want)
    for rnd in got["query_round_proofs"]:
        assert [len(row) for row, _ in rnd["initial_trees_proof"]] == [10, 9, 7]
    openings = fo.opening_set(polys, instance)
    assert fo.verify_openings(got, openings, [t.cap for t in trees], instance, vch, degree_bits, rate_bits, cap_height, pow_bits, rounds)
    assert fv.verify_openings(instance, openings, [t.cap for t in trees], got, pch, params)
    _same_proof(fri.prove_openings(batches_dev, instance, ch2, params), want)
    for b in batches_dev:
        b.free()


def test_one_call_prover_matches_golden(glb, ctx):
    """tests/golden/fri_proof.json replayed through gl_fri_prove: the flat word stream IS the fixture's proof_flat."""
    import json
    import os

    fri = importlib.import_module("plonky2-lib_b200.fri")
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "fri_proof.json")))
    vals = [np.array([[int(x, 16) for x in col] for col in v], dtype=np.uint64) for v in g["values"]]
    bs = [glb.PolynomialBatch.from_values(v, g["rate_bits"], False, g["cap_height"]) for v in vals]
    cfg = glb.FriConfig(rate_bits=g["rate_bits"], cap_height=g["cap_height"], proof_of_work_bits=g["proof_of_work_bits"],
                        num_query_rounds=g["num_query_rounds"])
    params = fri.FriParams.for_degree(cfg, g["degree_bits"])
    ch = fri.Challenger()
    for b in bs:
        ch.observe_cap(b.merkle_tree.cap)
    inst = [(tuple(pt), [tuple(p_) for p_ in ps]) for pt, ps in g["instance"]]
    flat = fri.prove_openings_device(bs, inst, ch, params, flat=True)
    assert [f"{int(x):x}" for x in flat] == g["proof_flat"]
    for b in bs:
        b.free()


def test_verifier_validates_the_proof_shape_first(glb, ctx, oracle, rng):
    """validate_fri_proof_shape: path lengths, step counts, evals per step, row widths and canonical words are
    fixed by the parameters; a malformed proof is a FriVerifyError, not an IndexError, and never reaches the hashing."""
    fri = importlib.import_module("plonky2-lib_b200.fri")
    fv = importlib.import_module("plonky2-lib_b200.fri_verifier")
    degree_bits, cols = 8, (4, 3)
    n = 1 << degree_bits
    bs = [glb.PolynomialBatch.from_values(oracle.synthetic_values(c, n, seed=60 + k), 3, False, 4) for k, c in enumerate(cols)]
    zeta = tuple(int(x) for x in rand_field(rng, (2,)))
    inst = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)])]
    cfg = glb.FriConfig(rate_bits=3, cap_height=4, proof_of_work_bits=6, num_query_rounds=8)
    params = fri.FriParams.for_degree(cfg, degree_bits)

    def fresh():
        c_ = fri.Challenger()
        for b in bs:
            c_.observe_cap(b.merkle_tree.cap)
        return c_

    proof = fri.prove_openings_device(bs, inst, fresh(), params)
    openings = fri.opening_set(bs, inst)
    caps = [b.merkle_tree.cap for b in bs]
    assert fv.verify_openings(inst, openings, caps, proof, fresh(), params)
    import copy

    def broken(edit):
        p_ = copy.deepcopy(proof)
        edit(p_)
        with pytest.raises(fv.FriVerifyError):
            fv.verify_openings(inst, openings, caps, p_, fresh(), params)

    broken(lambda p_: p_["query_round_proofs"][0]["steps"].pop())                                        # a step missing
    broken(lambda p_: p_["query_round_proofs"][1]["steps"][0].__setitem__("merkle_proof", p_["query_round_proofs"][1]["steps"][0]["merkle_proof"][:-1]))   # shorter path
    broken(lambda p_: p_["query_round_proofs"][2]["steps"][0].__setitem__("evals", p_["query_round_proofs"][2]["steps"][0]["evals"][:-1]))
    broken(lambda p_: p_["query_round_proofs"][3].__setitem__("initial_trees_proof", p_["query_round_proofs"][3]["initial_trees_proof"][:1]))
    broken(lambda p_: p_.__setitem__("pow_witness", p_["pow_witness"] + P))                              # non-canonical witness
    broken(lambda p_: p_["final_poly"].__setitem__((0, 0), np.uint64(P)))                               # non-canonical coefficient
    broken(lambda p_: p_.__setitem__("commit_phase_merkle_caps", p_["commit_phase_merkle_caps"][:-1]))
    broken(lambda p_: p_["query_round_proofs"].pop())
    for b in bs:
        b.free()
