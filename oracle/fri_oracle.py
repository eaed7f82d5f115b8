"""CPU ORACLE for plonky2::fri::{prover, verifier} and plonky2::iop::challenger (TEST INFRASTRUCTURE ONLY).

Restates upstream v0.1.4 (source absent; "parity unpinned": the reference holds no
FRI vectors) on top of the C oracle's primitives.  The verifier half (`verify_fri_proof`) is what gives the
prover half its evidence: a proof produced here or by the device must pass the same checks the upstream
verifier runs (proof of work, Merkle openings against every cap, fold consistency layer by layer, final
polynomial).  Reached in the reference through every `data.prove(pw)` / `data.verify(proof)` call site
(e.g. src/ecdsa/gadgets/ecdsa.rs:349,352).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import pyoracle as o

P = o.P
W = 7  # F[X]/(X^2 - 7)
RATE = 8


# ---- QuadraticExtension<GoldilocksField> on python ints ------------------------------------------------
def ext_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def ext_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def ext_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def ext_inv(a):
    # (a0 + a1 X)^-1 = (a0 - a1 X) / (a0^2 - 7 a1^2)
    d = pow((a[0] * a[0] - W * a[1] * a[1]) % P, P - 2, P)
    return (a[0] * d % P, (-a[1]) * d % P)


def ext_scalar(a, s):
    return (a[0] * s % P, a[1] * s % P)


class Challenger:
    """iop::challenger::Challenger: overwrite-mode duplex sponge over the Poseidon permutation."""

    def __init__(self):
        self.sponge_state = np.zeros(12, dtype=np.uint64)
        self.input_buffer: List[int] = []
        self.output_buffer: List[int] = []

    def observe_element(self, e):
        self.output_buffer = []
        self.input_buffer.append(int(e) % P)
        if len(self.input_buffer) == RATE:
            self.duplexing()

    def observe_elements(self, es):
        for e in np.asarray(es, dtype=np.uint64).reshape(-1).tolist():
            self.observe_element(e)

    observe_cap = observe_elements
    observe_hash = observe_elements
    observe_extension_elements = observe_elements

    def get_challenge(self) -> int:
        if self.input_buffer or not self.output_buffer:
            self.duplexing()
        return self.output_buffer.pop()

    def get_extension_challenge(self):
        return [self.get_challenge(), self.get_challenge()]

    def duplexing(self):
        for i, v in enumerate(self.input_buffer):
            self.sponge_state[i] = v
        self.input_buffer = []
        self.sponge_state = o.permute(self.sponge_state)
        self.output_buffer = [int(x) for x in self.sponge_state[:RATE]]


class MerkleTree:
    """hash::merkle_tree::MerkleTree on host arrays."""

    def __init__(self, leaves: np.ndarray, cap_height: int):
        self.leaves = np.ascontiguousarray(leaves, dtype=np.uint64)
        self.cap_height = cap_height
        self.digests, self.cap = o.merkle_tree(self.leaves, cap_height)

    def get(self, i):
        return self.leaves[i]

    def prove(self, i):
        return o.merkle_prove(self.digests, self.leaves.shape[0], self.cap_height, i)


def reduction_arity_bits(degree_bits, rate_bits, cap_height, arity_bits=4, final_poly_bits=5):
    out = []
    while degree_bits > final_poly_bits and degree_bits + rate_bits - arity_bits >= cap_height:
        out.append(arity_bits)
        degree_bits -= arity_bits
    return out


def fri_committed_trees(coeffs, values, challenger, arities, rate_bits, cap_height):
    coeffs, values = np.array(coeffs, dtype=np.uint64), np.array(values, dtype=np.uint64)
    trees = []
    shift = 7
    for ab in arities:
        leaves, _, _ = o.fri_layer_tree(values, ab, cap_height)      # reverse_index_bits + chunks(arity) + flatten
        tree = MerkleTree(leaves, cap_height)
        challenger.observe_cap(tree.cap)
        trees.append(tree)
        beta = challenger.get_extension_challenge()
        coeffs = o.fri_fold(coeffs, ab, np.array(beta, dtype=np.uint64))
        shift = pow(shift, 1 << ab, P)
        values = o.ext_coset_fft(coeffs, shift)
    coeffs = coeffs[: coeffs.shape[0] >> rate_bits].copy()
    challenger.observe_extension_elements(coeffs)
    return trees, coeffs


def fri_proof_of_work(challenger, pow_bits):
    state = challenger.sponge_state.copy()
    pos = len(challenger.input_buffer)
    for i, v in enumerate(challenger.input_buffer):
        state[i] = v
    w = 0
    while True:  # smallest witness (upstream: any witness, rayon find_any)
        w_found = o.pow_grind(state, pos, pow_bits, w, 1 << 20)
        if w_found != 0xFFFFFFFFFFFFFFFF:
            break
        w += 1 << 20
    challenger.observe_element(w_found)
    resp = challenger.get_challenge()
    assert pow_bits == 0 or resp >> (64 - pow_bits) == 0
    return w_found


def fri_proof(initial_trees: Sequence[MerkleTree], coeffs, values, challenger, degree_bits, rate_bits=3, cap_height=4,
              pow_bits=16, num_query_rounds=28):
    n = np.asarray(values).shape[0]
    arities = reduction_arity_bits(degree_bits, rate_bits, cap_height)
    trees, final_coeffs = fri_committed_trees(coeffs, values, challenger, arities, rate_bits, cap_height)
    pow_witness = fri_proof_of_work(challenger, pow_bits)
    rounds = []
    for _ in range(num_query_rounds):
        x_index = challenger.get_challenge() % n
        x0 = x_index
        initial = [(t.get(x_index).copy(), t.prove(x_index)) for t in initial_trees]
        steps = []
        for i, tree in enumerate(trees):
            ab = arities[i]
            steps.append({"evals": tree.get(x_index >> ab).reshape(-1, 2).copy(), "merkle_proof": tree.prove(x_index >> ab)})
            x_index >>= ab
        rounds.append({"x_index": x0, "initial_trees_proof": initial, "steps": steps})
    return {"commit_phase_merkle_caps": [t.cap for t in trees], "query_round_proofs": rounds, "final_poly": final_coeffs,
            "pow_witness": pow_witness}


# ---- verifier ---------------------------------------------------------------------------------------
def _rev(x, bits):
    return int(o.lib().glo_reverse_bits(x, bits)) if bits else 0


def compute_evaluation(x, x_index_within_coset, arity_bits, evals, beta):
    """fri::verifier::compute_evaluation: interpolate {(x g^i, P(x g^i))} and evaluate at beta (Lagrange)."""
    arity = 1 << arity_bits
    g = o.lib().glo_primitive_root_of_unity(arity_bits)
    ev = [tuple(int(v) for v in evals[_rev(i, arity_bits)]) for i in range(arity)]   # reverse_index_bits_in_place
    rev_within = _rev(x_index_within_coset, arity_bits)
    coset_start = x * pow(g, arity - rev_within, P) % P
    pts = [coset_start * pow(g, i, P) % P for i in range(arity)]
    acc = (0, 0)
    for i in range(arity):
        num, den = (1, 0), 1
        for j in range(arity):
            if i != j:
                num = ext_mul(num, ext_sub(beta, (pts[j], 0)))
                den = den * ((pts[i] - pts[j]) % P) % P
        acc = ext_add(acc, ext_mul(ev[i], ext_scalar(num, pow(den, P - 2, P))))
    return acc


def eval_ext_poly(coeffs, x):
    acc = (0, 0)
    for c in coeffs[::-1]:
        acc = ext_add(ext_mul(acc, x), (int(c[0]), int(c[1])))
    return acc


def verify_fri_proof(proof, initial_caps, initial_cap_height, challenger, degree_bits, first_layer_eval, rate_bits=3, cap_height=4,
                     pow_bits=16, num_query_rounds=28):
    """The checks of fri::verifier::verify_fri_proof, replaying the transcript.  `first_layer_eval(x_index,
    initial_rows)` stands in for fri_combine_initial (the alpha-combination of the opened rows, which belongs
    to prove_openings): it must return the value of the FRI polynomial at that LDE point."""
    lg_n = degree_bits + rate_bits
    n = 1 << lg_n
    arities = reduction_arity_bits(degree_bits, rate_bits, cap_height)
    betas = []
    for cap in proof["commit_phase_merkle_caps"]:
        challenger.observe_cap(cap)
        betas.append(tuple(challenger.get_extension_challenge()))
    challenger.observe_extension_elements(proof["final_poly"])
    challenger.observe_element(proof["pow_witness"])
    resp = challenger.get_challenge()
    assert pow_bits == 0 or resp >> (64 - pow_bits) == 0, "invalid proof of work"
    assert len(proof["final_poly"]) == (1 << degree_bits) >> sum(arities)
    assert len(proof["query_round_proofs"]) == num_query_rounds
    for rnd in proof["query_round_proofs"]:
        x_index = challenger.get_challenge() % n
        assert x_index == rnd["x_index"]
        for (row, path), cap in zip(rnd["initial_trees_proof"], initial_caps):
            assert o.merkle_verify(row, x_index, path, cap, initial_cap_height), "initial Merkle proof"
        subgroup_x = 7 * pow(o.lib().glo_primitive_root_of_unity(lg_n), _rev(x_index, lg_n), P) % P
        old_eval = first_layer_eval(x_index, [r for r, _ in rnd["initial_trees_proof"]], subgroup_x)
        for i, ab in enumerate(arities):
            evals = rnd["steps"][i]["evals"]
            coset_index, within = x_index >> ab, x_index & ((1 << ab) - 1)
            assert tuple(int(v) for v in evals[within]) == old_eval, f"layer {i}: inconsistent with the previous evaluation"
            old_eval = compute_evaluation(subgroup_x, within, ab, evals, betas[i])
            assert o.merkle_verify(np.ascontiguousarray(evals).reshape(-1), coset_index, rnd["steps"][i]["merkle_proof"],
                                   proof["commit_phase_merkle_caps"][i], cap_height), f"layer {i}: Merkle proof"
            subgroup_x = pow(subgroup_x, 1 << ab, P)
            x_index = coset_index
        assert eval_ext_poly(proof["final_poly"], (subgroup_x, 0)) == old_eval, "final polynomial mismatch"
    return True


# ---- fri::oracle::PolynomialBatch::prove_openings and the verifier's fri_combine_initial ---------------
# An instance is a list of batches: (point (a0, a1), [(oracle_index, polynomial_index), ...]); upstream's
# FriInstanceInfo for a proof has two: every polynomial at zeta, the Z polynomials at g * zeta.
def ext_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = ext_mul(r, a)
        a = ext_mul(a, a)
        e >>= 1
    return r


def final_poly_of_openings(oracle_polys, batches, alpha, times_x=False):
    """final_poly = sum_i alpha^(k_i) (F_i(X) - F_i(z_i)) / (X - z_i),  F_i = sum_j alpha^j f_ij.
    times_x: the older upstream form (PR #436) multiplies the result by X (`final_poly.coeffs.insert(0, ZERO)`; the quotient
    is then NOT padded with a zero, so the length is the same); the later form, the default, does not."""
    n = oracle_polys[0].shape[1]
    final = np.zeros((n, 2), dtype=np.uint64)
    for point, polys in batches:
        comp = o.reduce_polys_base([oracle_polys[oi][pi] for oi, pi in polys], alpha)
        quot = o.divide_by_linear(comp, point)           # remainder dropped, padded back with a zero
        shift = ext_pow(tuple(int(x) for x in alpha), len(polys))   # ReducingFactor::shift_poly
        final = o.ext_poly_scale_add(final, shift, quot)
    if times_x:
        assert not final[-1].any()
        final = np.concatenate([np.zeros((1, 2), dtype=np.uint64), final[:-1]])
    return final


def prove_openings(oracle_polys, oracle_trees, batches, challenger, degree_bits, rate_bits=3, cap_height=4, pow_bits=16,
                   num_query_rounds=28, times_x=False):
    alpha = challenger.get_extension_challenge()
    final = final_poly_of_openings(oracle_polys, batches, alpha, times_x)
    n = final.shape[0]
    lde_coeffs = np.zeros((n << rate_bits, 2), dtype=np.uint64)
    lde_coeffs[:n] = final
    lde_values = o.ext_coset_fft(lde_coeffs, 7)
    return fri_proof(oracle_trees, lde_coeffs, lde_values, challenger, degree_bits, rate_bits, cap_height, pow_bits,
                     num_query_rounds)


def opening_set(oracle_polys, batches):
    """OpeningSet: f_ij(z_i) for every polynomial of every batch (what the proof carries next to the FRI proof)."""
    return [[o.eval_base_poly_at_ext(oracle_polys[oi][pi], point) for oi, pi in polys] for point, polys in batches]


def fri_combine_initial(batches, openings, initial_rows, alpha, subgroup_x, times_x=False):
    """fri::verifier::fri_combine_initial: sum_i alpha^(k_i) (reduce(evals_i) - reduce(openings_i)) / (x - z_i);
    times_x: the older upstream form returns sum * subgroup_x (the prover multiplied final_poly by X)."""
    alpha = tuple(int(x) for x in alpha)
    total = (0, 0)
    for (point, polys), opened in zip(batches, openings):
        red_e, red_o = (0, 0), (0, 0)
        for (oi, pi), ov in zip(reversed(polys), reversed(opened)):     # ReducingFactor::reduce: Horner from the back
            red_e = ext_add(ext_mul(red_e, alpha), (int(initial_rows[oi][pi]), 0))
            red_o = ext_add(ext_mul(red_o, alpha), ov)
        num = ext_sub(red_e, red_o)
        den = ext_sub((subgroup_x, 0), (int(point[0]), int(point[1])))
        total = ext_mul(total, ext_pow(alpha, len(polys)))
        total = ext_add(total, ext_mul(num, ext_inv(den)))
    return ext_scalar(total, subgroup_x) if times_x else total


def verify_openings(proof, openings, oracle_caps, batches, challenger, degree_bits, rate_bits=3, cap_height=4, pow_bits=16,
                    num_query_rounds=28, times_x=False):
    """verify_fri_proof with the real fri_combine_initial: ties the opened rows of the initial trees and the
    claimed openings to the first FRI layer."""
    alpha = challenger.get_extension_challenge()

    def first_layer_eval(x_index, rows, subgroup_x):
        return fri_combine_initial(batches, openings, rows, alpha, subgroup_x, times_x)

    return verify_fri_proof(proof, oracle_caps, cap_height, challenger, degree_bits, first_layer_eval, rate_bits, cap_height,
                            pow_bits, num_query_rounds)
