"""plonky2's binary proof format (plonky2::util::serialization, `Write::write_proof_with_public_inputs` and the readers
that take their lengths from CommonCircuitData): what `ProofWithPublicInputs::to_bytes` / `from_bytes` produce, so that a
proof made on the device leaves the box in the form the reference's callers hold it in
(/root/reference/src/ecdsa/gadgets/ecdsa.rs:299-316 round-trips CircuitData the same way; proofs travel as these bytes).

Layout (v0.1.4, from memory of upstream -- "parity unpinned", the one fork-dependent point is listed in FORMAT_NOTES):
  field element          8 bytes, canonical u64 little endian            (write_field)
  extension element      2 field elements                                 (write_field_ext)
  hash / cap             4 field elements per HashOut, 2^cap_height per cap, NO length prefix  (write_hash, write_merkle_cap)
  vectors                elements back to back, NO length prefix: the reader knows every length from the circuit
  Merkle proof           1 byte = number of siblings, then the siblings   (write_merkle_proof)
  Proof                  wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap, OpeningSet, FriProof
  OpeningSet             constants, plonk_sigmas, wires, plonk_zs, plonk_zs_next, partial_products, quotient_polys
  FriProof               commit_phase_merkle_caps; per query round: per oracle (leaf row, Merkle proof), per reduction step
                         (evals, Merkle proof); final_poly coefficients; pow_witness
  ProofWithPublicInputs  Proof, then the public inputs
The query index is not part of a FriQueryRound: the verifier re-derives it from the transcript.
"""
from __future__ import annotations

import io
from typing import Dict, List, Sequence

import numpy as np

P = 0xFFFFFFFF00000001
OPENING_SET_FIELDS = ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products", "quotient_polys")
FORMAT_NOTES = ("upstream revisions with lookup arguments append lookup_zs / lookup_zs_next to the OpeningSet (empty, i.e. zero "
                "bytes, for circuits without lookup tables -- none of the reference's circuits has one)")


class WireError(ValueError):
    pass


def _canon(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1)
    if (a >= np.uint64(P)).any():
        raise WireError("non-canonical field element")
    return a.astype("<u8", copy=False)


def _merkle_proof(out: io.BytesIO, path):
    sib = np.asarray(path, dtype=np.uint64).reshape(-1, 4)
    if sib.shape[0] > 255:
        raise WireError("Merkle proof length must fit in u8.")
    out.write(bytes([sib.shape[0]]))
    out.write(_canon(sib).tobytes())


def fri_proof_to_bytes(proof: dict) -> bytes:
    """write_fri_proof: proof = the dict fri.prove_openings / prove_openings_device returns."""
    out = io.BytesIO()
    for cap in proof["commit_phase_merkle_caps"]:
        out.write(_canon(cap).tobytes())
    for rnd in proof["query_round_proofs"]:
        for row, path in rnd["initial_trees_proof"]:
            out.write(_canon(row).tobytes())
            _merkle_proof(out, path)
        for st in rnd["steps"]:
            out.write(_canon(st["evals"]).tobytes())
            _merkle_proof(out, st["merkle_proof"])
    out.write(_canon(proof["final_poly"]).tobytes())
    out.write(_canon([int(proof["pow_witness"])]).tobytes())
    return out.getvalue()


def proof_with_public_inputs_to_bytes(caps: Sequence, openings: Dict[str, Sequence], fri_proof: dict, public_inputs) -> bytes:
    """write_proof_with_public_inputs.  caps = (wires_cap, plonk_zs_partial_products_cap, quotient_polys_cap);
    openings: OpeningSet field name -> [k][2] extension values."""
    out = io.BytesIO()
    if len(caps) != 3:
        raise WireError("a Proof carries three caps")
    for cap in caps:
        out.write(_canon(cap).tobytes())
    for f in OPENING_SET_FIELDS:
        out.write(_canon(np.asarray(openings.get(f, np.zeros((0, 2))), dtype=np.uint64)).tobytes())
    out.write(fri_proof_to_bytes(fri_proof))
    out.write(_canon(public_inputs).tobytes())
    return out.getvalue()


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.at = memoryview(data), 0

    def fields(self, k: int, shape=None) -> np.ndarray:
        end = self.at + 8 * k
        if end > len(self.b):
            raise WireError("unexpected end of proof bytes")
        a = np.frombuffer(self.b[self.at:end], dtype="<u8").astype(np.uint64)
        self.at = end
        if (a >= np.uint64(P)).any():
            raise WireError("non-canonical field element")      # read_field: from_canonical rejects >= p in debug builds
        return a.reshape(shape) if shape is not None else a

    def merkle_proof(self, expect: int) -> np.ndarray:
        if self.at >= len(self.b):
            raise WireError("unexpected end of proof bytes")
        k = self.b[self.at]
        self.at += 1
        if k != expect:
            raise WireError(f"Merkle proof of {k} siblings where the circuit fixes {expect}")
        return self.fields(4 * k, (k, 4))


def fri_proof_from_reader(r: _Reader, oracle_columns: Sequence[int], degree_bits: int, rate_bits: int, cap_height: int,
                          reduction_arity_bits: Sequence[int], num_query_rounds: int) -> dict:
    lgN = degree_bits + rate_bits
    caps = [r.fields(4 << cap_height, (1 << cap_height, 4)) for _ in reduction_arity_bits]
    rounds = []
    for _ in range(num_query_rounds):
        init = [(r.fields(c), r.merkle_proof(lgN - cap_height)) for c in oracle_columns]
        steps, cur = [], lgN
        for ab in reduction_arity_bits:
            ev = r.fields(2 << ab, (1 << ab, 2))
            cur -= ab
            steps.append({"evals": ev, "merkle_proof": r.merkle_proof(cur - cap_height)})
        rounds.append({"initial_trees_proof": init, "steps": steps})
    final_len = (1 << degree_bits) >> sum(reduction_arity_bits)
    final = r.fields(2 * final_len, (final_len, 2))
    pow_witness = int(r.fields(1)[0])
    return {"commit_phase_merkle_caps": caps, "query_round_proofs": rounds, "final_poly": final, "pow_witness": pow_witness}


def fri_proof_from_bytes(data: bytes, oracle_columns, degree_bits, rate_bits, cap_height, reduction_arity_bits, num_query_rounds) -> dict:
    r = _Reader(data)
    p = fri_proof_from_reader(r, oracle_columns, degree_bits, rate_bits, cap_height, reduction_arity_bits, num_query_rounds)
    if r.at != len(data):
        raise WireError("trailing bytes after the FRI proof")
    return p


def proof_with_public_inputs_from_bytes(data: bytes, opening_lengths: Dict[str, int], oracle_columns, degree_bits, rate_bits,
                                        cap_height, reduction_arity_bits, num_query_rounds):
    """read_proof_with_public_inputs: every length comes from the circuit (opening_lengths: OpeningSet field -> count);
    the public inputs are whatever field elements remain."""
    r = _Reader(data)
    caps = [r.fields(4 << cap_height, (1 << cap_height, 4)) for _ in range(3)]
    openings = {f: r.fields(2 * opening_lengths.get(f, 0), (opening_lengths.get(f, 0), 2)) for f in OPENING_SET_FIELDS}
    fp = fri_proof_from_reader(r, oracle_columns, degree_bits, rate_bits, cap_height, reduction_arity_bits, num_query_rounds)
    rest = len(data) - r.at
    if rest % 8:
        raise WireError("public inputs are not a whole number of field elements")
    return caps, openings, fp, r.fields(rest // 8)


def fri_proof_num_bytes(oracle_columns, degree_bits, rate_bits, cap_height, reduction_arity_bits, num_query_rounds) -> int:
    lgN = degree_bits + rate_bits
    per_round = sum(8 * c + 1 + 32 * (lgN - cap_height) for c in oracle_columns)
    cur = lgN
    for ab in reduction_arity_bits:
        cur -= ab
        per_round += 16 * (1 << ab) + 1 + 32 * (cur - cap_height)
    final_len = (1 << degree_bits) >> sum(reduction_arity_bits)
    return len(reduction_arity_bits) * (32 << cap_height) + num_query_rounds * per_round + 16 * final_len + 8
