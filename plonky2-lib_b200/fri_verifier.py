"""plonky2::fri::verifier on top of the C ABI: the FRI half of `data.verify(proof)` (42 call sites in the reference, e.g.
src/ecdsa/gadgets/ecdsa.rs:352, src/hash/keccak256.rs:249, src/smt/gadgets/process/mod.rs:84).

What is data-parallel goes to the device: every Merkle opening of the proof (28 query rounds x (initial trees + layer
trees)) is checked with one `gl_merkle_verify_batch` per tree, and the transcript's permutations are the product
Challenger's.  What is a few hundred field operations per query round (alpha-reduction of the opened rows,
`compute_evaluation`'s interpolation over a coset of 16 points, the final polynomial) stays on the host in exact
Python integers, as it stays in Rust upstream.  Structure and names follow upstream: `fri_challenges`,
`PrecomputedReducedOpenings`, `fri_combine_initial`, `compute_evaluation`, `fri_verifier_query_round`,
`verify_fri_proof`; a failed `ensure!` raises FriVerifyError with upstream's message.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

from .fri import Challenger, FriParams
from .host import Context, _ctx, merkle_verify_batch

P = 0xFFFFFFFF00000001
W = 7                       # QuadraticExtension: X^2 - 7
Ext = Tuple[int, int]


class FriVerifyError(ValueError):
    """an `ensure!` of plonky2::fri::verifier failed"""


# ---- F and F[X]/(X^2 - 7) in exact integers -----------------------------------------------------------------------
def _e(x) -> Ext:
    return int(x[0]) % P, int(x[1]) % P


def _eadd(x: Ext, y: Ext) -> Ext:
    return (x[0] + y[0]) % P, (x[1] + y[1]) % P


def _esub(x: Ext, y: Ext) -> Ext:
    return (x[0] - y[0]) % P, (x[1] - y[1]) % P


def _emul(x: Ext, y: Ext) -> Ext:
    return (x[0] * y[0] + W * x[1] * y[1]) % P, (x[0] * y[1] + x[1] * y[0]) % P


def _escale(x: Ext, s: int) -> Ext:
    return x[0] * s % P, x[1] * s % P


def _einv(x: Ext) -> Ext:
    norm = (x[0] * x[0] - W * x[1] * x[1]) % P          # x * conj(x)
    ni = pow(norm, P - 2, P)
    return x[0] * ni % P, (P - x[1]) * ni % P


def _epow(x: Ext, e: int) -> Ext:
    acc: Ext = (1, 0)
    while e:
        if e & 1:
            acc = _emul(acc, x)
        x = _emul(x, x)
        e >>= 1
    return acc


def _root_of_unity(bits: int) -> int:
    return pow(pow(7, (P - 1) >> 32, P), 1 << (32 - bits), P)


def _bitrev(x: int, bits: int) -> int:
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


# ---- challenges (plonky2::fri::challenges) ---------------------------------------------------------------------------
def fri_challenges(challenger: Challenger, proof: dict, degree_bits: int, params: FriParams) -> dict:
    """Challenger::fri_challenges: alpha, one beta per committed layer, the proof-of-work response, the query indices."""
    cfg = params.config
    alpha = tuple(challenger.get_extension_challenge())
    betas = []
    for cap in proof["commit_phase_merkle_caps"]:
        challenger.observe_cap(cap)
        betas.append(tuple(challenger.get_extension_challenge()))
    challenger.observe_extension_elements(proof["final_poly"])
    challenger.observe_element(int(proof["pow_witness"]))
    pow_response = challenger.get_challenge()
    lde_size = 1 << (degree_bits + cfg.rate_bits)
    indices = [challenger.get_challenge() % lde_size for _ in range(cfg.num_query_rounds)]
    return {"fri_alpha": alpha, "fri_betas": betas, "fri_pow_response": pow_response, "fri_query_indices": indices}


def precomputed_reduced_openings(openings: Sequence[Sequence[Ext]], alpha: Ext) -> List[Ext]:
    """PrecomputedReducedOpenings::from_os_and_alpha: ReducingFactor(alpha).reduce of every batch's opened values."""
    out = []
    for batch in openings:
        acc: Ext = (0, 0)
        for v in reversed(list(batch)):
            acc = _eadd(_emul(acc, alpha), _e(v))
        out.append(acc)
    return out


def fri_combine_initial(instance, initial_rows, alpha: Ext, subgroup_x: int, reduced_openings: Sequence[Ext],
                        times_x: bool = False) -> Ext:
    """sum over the batches of alpha^(polys so far) * (reduce(evals) - reduce(openings)) / (x - z), accumulated the way
    ReducingFactor::shift does (multiply what is there by alpha^len before adding the next quotient).  times_x
    (FriParams.final_poly_times_x, GL_COMPAT_FRI_FINAL_POLY_TIMES_X): the older upstream form returns sum * subgroup_x."""
    total: Ext = (0, 0)
    for (point, polys), red_open in zip(instance, reduced_openings):
        red_eval: Ext = (0, 0)
        for oi, pi in reversed(list(polys)):
            red_eval = _eadd(_emul(red_eval, alpha), (int(initial_rows[oi][pi]) % P, 0))
        quotient = _emul(_esub(red_eval, red_open), _einv(_esub((subgroup_x, 0), _e(point))))
        total = _eadd(_emul(total, _epow(alpha, len(polys))), quotient)
    return _escale(total, subgroup_x) if times_x else total


def compute_evaluation(x: int, x_index_within_coset: int, arity_bits: int, evals, beta: Ext) -> Ext:
    """The value at beta of the polynomial of degree < arity through the coset's points (barycentric form, as upstream's
    `interpolate`): the evaluations arrive in bit-reversed order, the coset starts at x * g^(-rev(index))."""
    arity = 1 << arity_bits
    g = _root_of_unity(arity_bits)
    ordered = [_e(evals[_bitrev(i, arity_bits)]) for i in range(arity)]
    start = x * pow(g, (arity - _bitrev(x_index_within_coset, arity_bits)) % arity, P) % P
    pts = [start * pow(g, i, P) % P for i in range(arity)]
    for p, v in zip(pts, ordered):                 # beta on the coset itself (probability ~ 2^-124, but exact)
        if beta == (p, 0):
            return v
    # weights w_i = 1 / prod_{j != i} (x_i - x_j); for a coset of a subgroup: prod = arity * x_i^(arity-1)
    l_beta: Ext = (1, 0)
    for p in pts:
        l_beta = _emul(l_beta, _esub(beta, (p, 0)))
    acc: Ext = (0, 0)
    for p, v in zip(pts, ordered):
        w = pow(arity * pow(p, arity - 1, P) % P, P - 2, P)
        acc = _eadd(acc, _emul(_escale(v, w), _einv(_esub(beta, (p, 0)))))
    return _emul(l_beta, acc)


def _eval_final_poly(coeffs, x: int) -> Ext:
    acc: Ext = (0, 0)
    for c in reversed(list(coeffs)):
        acc = _eadd(_escale(acc, x), _e(c))
    return acc


# ---- validate_fri_proof_shape ------------------------------------------------------------------------------------------
def validate_fri_proof_shape(proof: dict, instance, initial_merkle_caps, params: FriParams, oracle_columns=None) -> None:
    """plonky2::fri::validate_shape::validate_fri_proof_shape: everything about the proof's SHAPE is fixed by the parameters,
    never taken from the proof -- cap sizes, number of layers / rounds / steps, evals per step, Merkle path lengths
    (lg N - cap_height, minus the arity bits folded so far for layer trees), row widths per oracle -- and every word must be
    a canonical field element.  Raises FriVerifyError before any hashing happens."""
    cfg = params.config
    arities = list(params.reduction_arity_bits)
    lg_n, h = params.degree_bits + cfg.rate_bits, cfg.cap_height

    def canonical(a, what):
        a = np.asarray(a)
        if a.dtype.kind not in "ui" or (a.astype(np.uint64, copy=False) >= np.uint64(P)).any():
            raise FriVerifyError(f"{what}: not a canonical field element")
        return a

    try:
        caps = proof["commit_phase_merkle_caps"]
        rounds = proof["query_round_proofs"]
        final = np.asarray(proof["final_poly"])
        pow_witness = int(proof["pow_witness"])
    except (KeyError, TypeError, ValueError) as e:
        raise FriVerifyError(f"malformed proof: {e!r}")
    if not 0 <= pow_witness < P:
        raise FriVerifyError("pow_witness: not a canonical field element")
    if len(caps) != len(arities):
        raise FriVerifyError("The number of committed layers does not match the reduction strategy.")
    for cap in list(caps) + list(initial_merkle_caps):
        if np.asarray(cap).shape != (1 << h, 4):
            raise FriVerifyError("Merkle cap of the wrong height.")
        canonical(cap, "cap")
    if sum(arities) > params.degree_bits or final.shape != ((1 << params.degree_bits) >> sum(arities), 2):
        raise FriVerifyError("Final polynomial has wrong degree.")
    canonical(final, "final_poly")
    if len(rounds) != cfg.num_query_rounds:
        raise FriVerifyError("Number of query rounds does not match config.")
    if oracle_columns is None:
        oracle_columns = [None] * len(initial_merkle_caps)
        for _, polys in instance:
            for oi, pi in polys:
                if oi >= len(oracle_columns):
                    raise FriVerifyError("instance refers to an oracle without a cap")
                oracle_columns[oi] = max(oracle_columns[oi] or 0, pi + 1)
        exact = False
    else:
        exact = True
    for rnd in rounds:
        try:
            init, steps = rnd["initial_trees_proof"], rnd["steps"]
        except (KeyError, TypeError) as e:
            raise FriVerifyError(f"malformed query round: {e!r}")
        if len(init) != len(initial_merkle_caps):
            raise FriVerifyError("Wrong number of initial trees in a query round.")
        for (row, path), want in zip(init, oracle_columns):
            row, path = canonical(row, "leaf"), canonical(path, "sibling")
            if row.ndim != 1 or (want is not None and (row.shape[0] != want if exact else row.shape[0] < want)):
                raise FriVerifyError("Initial-tree leaf of the wrong width.")
            if path.reshape(-1, 4).shape != (lg_n - h, 4):
                raise FriVerifyError("Initial-tree Merkle proof of the wrong length.")
        if len(steps) != len(arities):
            raise FriVerifyError("Wrong number of reduction steps in a query round.")
        cur = lg_n
        for st, ab in zip(steps, arities):
            ev, path = canonical(st["evals"], "evals"), canonical(st["merkle_proof"], "sibling")
            if ev.reshape(-1, 2).shape != (1 << ab, 2) or ev.size != 2 << ab:
                raise FriVerifyError("Wrong number of evaluations in a reduction step.")
            cur -= ab
            if cur < h or path.reshape(-1, 4).shape != (cur - h, 4):
                raise FriVerifyError("Reduction-layer Merkle proof of the wrong length.")


# ---- verify_fri_proof --------------------------------------------------------------------------------------------------
def verify_fri_proof(instance, openings, challenges: dict, initial_merkle_caps, proof: dict, params: FriParams,
                     ctx: Context = None) -> bool:
    """plonky2::fri::verifier::verify_fri_proof.  instance: [(point, [(oracle_index, polynomial_index), ...]), ...]
    (FriInstanceInfo.batches); openings: the claimed values per batch (FriOpenings); initial_merkle_caps: the caps of the
    oracles.  Raises FriVerifyError where upstream returns Err; returns True otherwise."""
    ctx = _ctx(ctx)
    cfg = params.config
    arities = list(params.reduction_arity_bits)
    lg_n = params.degree_bits + cfg.rate_bits
    validate_fri_proof_shape(proof, instance, initial_merkle_caps, params)
    times_x = bool(getattr(params, "final_poly_times_x", False))
    if len(proof["final_poly"]) != (1 << params.degree_bits) >> sum(arities):
        raise FriVerifyError("Final polynomial has wrong degree.")
    # fri_verify_proof_of_work
    if cfg.proof_of_work_bits and challenges["fri_pow_response"] >> (64 - cfg.proof_of_work_bits):
        raise FriVerifyError("Invalid proof-of-work.")
    rounds = proof["query_round_proofs"]
    if len(rounds) != cfg.num_query_rounds:
        raise FriVerifyError("Number of query rounds does not match config.")
    if len(proof["commit_phase_merkle_caps"]) != len(arities) or len(challenges["fri_betas"]) != len(arities):
        raise FriVerifyError("The number of committed layers does not match the reduction strategy.")
    alpha = challenges["fri_alpha"]
    reduced = precomputed_reduced_openings(openings, alpha)
    indices = [int(x) for x in challenges["fri_query_indices"]]

    # every Merkle opening of the proof, one device call per tree
    for ti, cap in enumerate(initial_merkle_caps):
        rows = np.stack([np.asarray(r["initial_trees_proof"][ti][0], dtype=np.uint64) for r in rounds])
        paths = np.stack([np.asarray(r["initial_trees_proof"][ti][1], dtype=np.uint64).reshape(-1, 4) for r in rounds])
        if not merkle_verify_batch(rows, np.array(indices, dtype=np.uint64), paths, cap, ctx=ctx).all():
            raise FriVerifyError("Invalid Merkle proof (initial tree %d)." % ti)
    layer_index = list(indices)
    for li, ab in enumerate(arities):
        layer_index = [x >> ab for x in layer_index]
        rows = np.stack([np.asarray(r["steps"][li]["evals"], dtype=np.uint64).reshape(-1) for r in rounds])
        paths = np.stack([np.asarray(r["steps"][li]["merkle_proof"], dtype=np.uint64).reshape(-1, 4) for r in rounds])
        if not merkle_verify_batch(rows, np.array(layer_index, dtype=np.uint64), paths, proof["commit_phase_merkle_caps"][li],
                                   ctx=ctx).all():
            raise FriVerifyError("Invalid Merkle proof (reduction layer %d)." % li)

    # fri_verifier_query_round: the consistency of the evaluations from layer to layer
    w_n = _root_of_unity(lg_n)
    for q, rnd in enumerate(rounds):
        x_index = indices[q]
        if "x_index" in rnd and int(rnd["x_index"]) != x_index:
            raise FriVerifyError("query index does not come from the transcript")
        subgroup_x = 7 * pow(w_n, _bitrev(x_index, lg_n), P) % P
        old_eval = fri_combine_initial(instance, [row for row, _ in rnd["initial_trees_proof"]], alpha, subgroup_x, reduced, times_x)
        for li, ab in enumerate(arities):
            evals = np.asarray(rnd["steps"][li]["evals"], dtype=np.uint64).reshape(-1, 2)
            within = x_index & ((1 << ab) - 1)
            if _e(evals[within]) != old_eval:
                raise FriVerifyError("Inconsistent evaluation with the previous layer (query %d, layer %d)." % (q, li))
            old_eval = compute_evaluation(subgroup_x, within, ab, evals, challenges["fri_betas"][li])
            subgroup_x = pow(subgroup_x, 1 << ab, P)
            x_index >>= ab
        if _eval_final_poly(proof["final_poly"], subgroup_x) != old_eval:
            raise FriVerifyError("Final polynomial evaluation is invalid.")
    return True


def verify_openings(instance, openings, initial_merkle_caps, proof: dict, challenger: Challenger, params: FriParams,
                    ctx: Context = None) -> bool:
    """The verifier's counterpart of fri.prove_openings: draw the challenges from the transcript (the caller has observed
    the caps and the openings as the prover's caller did), then verify_fri_proof."""
    ch = fri_challenges(challenger, proof, params.degree_bits, params)
    return verify_fri_proof(instance, openings, ch, initial_merkle_caps, proof, params, ctx)
