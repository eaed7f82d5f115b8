"""ctypes binding of the CPU ORACLE (oracle/libgl_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never by the product package.  Parity status is stated in
oracle/gl_oracle.h ("pinned" for Poseidon, "parity unpinned" for NTT/LDE/Merkle/FRI layouts).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgl_oracle.so")
P = 0xFFFFFFFF00000001

u64p = C.POINTER(C.c_uint64)


def build(force: bool = False) -> str:
    """Compile oracle/gl_oracle.c with the committed Makefile (gcc, OpenMP)."""
    src = os.path.join(_HERE, "gl_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB_PATH)
    ):
        subprocess.check_call(["make", "-C", _HERE, "-s"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return _LIB_PATH


class SmtProcessProof(C.Structure):
    """Mirror of glo_smt_process_proof == SparseMerkleProcessProof (src/smt/proof/process.rs:12-23)."""

    _fields_ = [
        ("old_root", C.c_uint64 * 4),
        ("old_key", C.c_uint64 * 4),
        ("old_value", C.c_uint64 * 4),
        ("new_root", C.c_uint64 * 4),
        ("new_key", C.c_uint64 * 4),
        ("new_value", C.c_uint64 * 4),
        ("siblings", (C.c_uint64 * 4) * 256),
        ("num_siblings", C.c_uint32),
        ("is_old0", C.c_uint32),
        ("fnc", C.c_uint32),
        ("pad_", C.c_uint32),
    ]


SMT_PROOF_DTYPE = np.dtype(
    [
        ("old_root", "<u8", 4),
        ("old_key", "<u8", 4),
        ("old_value", "<u8", 4),
        ("new_root", "<u8", 4),
        ("new_key", "<u8", 4),
        ("new_value", "<u8", 4),
        ("siblings", "<u8", (256, 4)),
        ("num_siblings", "<u4"),
        ("is_old0", "<u4"),
        ("fnc", "<u4"),
        ("pad_", "<u4"),
    ]
)
assert SMT_PROOF_DTYPE.itemsize == C.sizeof(SmtProcessProof)

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    sz = C.c_size_t
    u = C.c_uint
    sigs = {
        "glo_add": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "glo_sub": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "glo_mul": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "glo_pow": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "glo_inv": (C.c_uint64, [C.c_uint64]),
        "glo_primitive_root_of_unity": (C.c_uint64, [u]),
        "glo_ext_mul": (None, [u64p, u64p, u64p]),
        "glo_poseidon_round_constants": (None, [u64p]),
        "glo_poseidon_permute": (None, [u64p]),
        "glo_poseidon_permute_naive": (None, [u64p]),
        "glo_poseidon_permute_slow": (None, [u64p]),
        "glo_hash_no_pad": (None, [u64p, sz, u64p]),
        "glo_hash_pad": (None, [u64p, sz, u64p]),
        "glo_hash_or_noop": (None, [u64p, sz, u64p]),
        "glo_two_to_one": (None, [u64p, u64p, u64p]),
        "glo_permute_batch": (None, [u64p, sz]),
        "glo_two_to_one_batch": (None, [u64p, u64p, u64p, sz]),
        "glo_hash_no_pad_batch": (None, [u64p, sz, sz, u64p]),
        "glo_fft": (None, [u64p, u, u]),
        "glo_ifft": (None, [u64p, u]),
        "glo_coset_fft": (None, [u64p, u, C.c_uint64, u]),
        "glo_coset_ifft": (None, [u64p, u, C.c_uint64]),
        "glo_reverse_bits": (sz, [sz, u]),
        "glo_merkle_tree": (C.c_int, [u64p, sz, sz, u, u64p, u64p]),
        "glo_merkle_prove": (None, [u64p, sz, u, sz, u64p]),
        "glo_merkle_verify": (C.c_int, [u64p, sz, sz, u64p, u, u64p, u]),
        "glo_commit_from_values": (C.c_int, [u64p, u, u, u, u, u64p, u64p, u64p, u64p]),
        "glo_commit_from_coeffs": (C.c_int, [u64p, u, u, u, u, u64p, u64p, u64p]),
        "glo_smt_leaf_hash": (None, [u64p, u64p, u64p]),
        "glo_smt_internal_hash": (None, [u64p, u64p, u64p]),
        "glo_smt_verify_process_proof": (C.c_int, [C.c_void_p]),
        "glo_smt_verify_process_batch": (None, [C.c_void_p, sz, C.POINTER(C.c_int32)]),
        "glo_smt_new": (C.c_void_p, []),
        "glo_smt_free": (None, [C.c_void_p]),
        "glo_smt_root": (None, [C.c_void_p, u64p]),
        "glo_smt_set": (C.c_int, [C.c_void_p, u64p, u64p, C.c_void_p]),
        "glo_smt_find": (C.c_int, [C.c_void_p, u64p, u64p, C.POINTER(C.c_uint32), u64p, u64p, C.POINTER(C.c_uint32)]),
        "glo_fri_layer_tree": (C.c_int, [u64p, sz, u, u, u64p, u64p, u64p]),
        "glo_fri_fold": (None, [u64p, sz, u, u64p, u64p]),
        "glo_ext_coset_fft": (None, [u64p, u, C.c_uint64]),
        "glo_pow_grind": (C.c_uint64, [u64p, u, u, u, C.c_uint64, C.c_uint64]),
        "glo_reduce_polys_base": (None, [C.POINTER(u64p), sz, sz, u64p, u64p]),
        "glo_divide_by_linear": (None, [u64p, sz, u64p, u64p]),
        "glo_ext_poly_scale_add": (None, [u64p, sz, u64p, u64p]),
        "glo_eval_base_poly_at_ext": (None, [u64p, sz, u64p, u64p]),
        "glo_num_gate_constraints": (C.c_uint, [C.c_void_p, u]),
        "glo_permutation_zs": (C.c_int, [C.c_void_p, u64p, u64p, u64p, u64p, u64p, u64p]),
        "glo_quotient_polys": (C.c_int, [C.c_void_p, C.c_void_p, u64p, u64p, u64p, u64p, u, u64p, u64p, u64p, u64p, u64p]),
        "glo_vanishing_at_point": (None, [C.c_void_p, C.c_void_p, u64p, C.c_uint64, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u64p, u64p]),
        "glo_num_threads": (C.c_int, []),
        "glo_set_num_threads": (None, [C.c_int]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _a(x) -> np.ndarray:
    return np.ascontiguousarray(x, dtype=np.uint64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(u64p) if a is not None else None


# ---- convenience wrappers (numpy in, numpy out) -------------------------------------------------
def round_constants() -> np.ndarray:
    out = np.zeros(360, dtype=np.uint64)
    lib().glo_poseidon_round_constants(_p(out))
    return out


def permute(state) -> np.ndarray:
    s = _a(state).copy()
    assert s.shape == (12,)
    lib().glo_poseidon_permute(_p(s))
    return s


def permute_batch(states) -> np.ndarray:
    s = _a(states).copy()
    assert s.ndim == 2 and s.shape[1] == 12
    lib().glo_permute_batch(_p(s), s.shape[0])
    return s


def hash_no_pad(x) -> np.ndarray:
    x = _a(x)
    out = np.zeros(4, dtype=np.uint64)
    lib().glo_hash_no_pad(_p(x), x.size, _p(out))
    return out


def hash_pad(x) -> np.ndarray:
    x = _a(x)
    out = np.zeros(4, dtype=np.uint64)
    lib().glo_hash_pad(_p(x), x.size, _p(out))
    return out


def hash_or_noop(x) -> np.ndarray:
    x = _a(x)
    out = np.zeros(4, dtype=np.uint64)
    lib().glo_hash_or_noop(_p(x), x.size, _p(out))
    return out


def two_to_one(l, r) -> np.ndarray:
    l, r = _a(l), _a(r)
    out = np.zeros(4, dtype=np.uint64)
    lib().glo_two_to_one(_p(l), _p(r), _p(out))
    return out


def two_to_one_batch(l, r) -> np.ndarray:
    l, r = _a(l), _a(r)
    out = np.zeros_like(l)
    lib().glo_two_to_one_batch(_p(l), _p(r), _p(out), l.shape[0])
    return out


def hash_no_pad_batch(x) -> np.ndarray:
    x = _a(x)
    out = np.zeros((x.shape[0], 4), dtype=np.uint64)
    lib().glo_hash_no_pad_batch(_p(x), x.shape[1], x.shape[0], _p(out))
    return out


def fft(a, zero_factor: int = 0) -> np.ndarray:
    a = _a(a).copy()
    lib().glo_fft(_p(a), int(a.size).bit_length() - 1, zero_factor)
    return a


def ifft(a) -> np.ndarray:
    a = _a(a).copy()
    lib().glo_ifft(_p(a), int(a.size).bit_length() - 1)
    return a


def coset_fft(a, shift: int = 7, zero_factor: int = 0) -> np.ndarray:
    a = _a(a).copy()
    lib().glo_coset_fft(_p(a), int(a.size).bit_length() - 1, shift, zero_factor)
    return a


def coset_ifft(a, shift: int = 7) -> np.ndarray:
    a = _a(a).copy()
    lib().glo_coset_ifft(_p(a), int(a.size).bit_length() - 1, shift)
    return a


def merkle_tree(leaves, cap_height: int):
    """MerkleTree::new(leaves, cap_height) -> (digests [2(N-2^h),4], cap [2^h,4])."""
    leaves = _a(leaves)
    n, ll = leaves.shape
    nd = 2 * (n - (1 << cap_height))
    digests = np.zeros((max(nd, 0), 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().glo_merkle_tree(_p(leaves), n, ll, cap_height, _p(digests), _p(cap))
    if rc:
        raise ValueError(f"MerkleTree::new would panic (code {rc})")
    return digests, cap


def merkle_prove(digests, num_leaves: int, cap_height: int, idx: int) -> np.ndarray:
    digests = _a(digests)
    L = int(num_leaves).bit_length() - 1 - cap_height
    sib = np.zeros((L, 4), dtype=np.uint64)
    lib().glo_merkle_prove(_p(digests), num_leaves, cap_height, idx, _p(sib))
    return sib


def merkle_verify(leaf, idx: int, siblings, cap, cap_height: int) -> bool:
    leaf, siblings, cap = _a(leaf), _a(siblings), _a(cap)
    return bool(
        lib().glo_merkle_verify(_p(leaf), leaf.size, idx, _p(siblings), siblings.shape[0], _p(cap), cap_height)
    )


def commit_from_values(values, rate_bits: int, cap_height: int, want_leaves: bool = True):
    """PolynomialBatch::from_values -> dict(coeffs [c,n], leaves [N,c], digests, cap)."""
    values = _a(values)
    c, n = values.shape
    lg_n = int(n).bit_length() - 1
    N = n << rate_bits
    coeffs = np.zeros_like(values)
    leaves = np.zeros((N, c), dtype=np.uint64) if want_leaves else None
    nd = 2 * (N - (1 << cap_height))
    digests = np.zeros((nd, 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().glo_commit_from_values(
        _p(values), lg_n, c, rate_bits, cap_height, _p(coeffs), _p(leaves) if want_leaves else None, _p(digests), _p(cap)
    )
    if rc:
        raise ValueError(f"PolynomialBatch::from_values would panic (code {rc})")
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


def commit_from_coeffs(coeffs, rate_bits: int, cap_height: int, want_leaves: bool = True):
    coeffs = _a(coeffs)
    c, n = coeffs.shape
    lg_n = int(n).bit_length() - 1
    N = n << rate_bits
    leaves = np.zeros((N, c), dtype=np.uint64) if want_leaves else None
    nd = 2 * (N - (1 << cap_height))
    digests = np.zeros((nd, 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().glo_commit_from_coeffs(
        _p(coeffs), lg_n, c, rate_bits, cap_height, _p(leaves) if want_leaves else None, _p(digests), _p(cap)
    )
    if rc:
        raise ValueError(f"PolynomialBatch::from_coeffs would panic (code {rc})")
    return {"coeffs": coeffs, "leaves": leaves, "digests": digests, "cap": cap}


def smt_leaf_hash(k, v) -> np.ndarray:
    k, v = _a(k), _a(v)
    out = np.zeros(4, dtype=np.uint64)
    lib().glo_smt_leaf_hash(_p(k), _p(v), _p(out))
    return out


class Smt:
    """PoseidonSparseMerkleTreeMemory (src/smt/goldilocks_poseidon/mod.rs:193) restated in C."""

    def __init__(self):
        self._h = lib().glo_smt_new()

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().glo_smt_free(self._h)
                self._h = None
        except Exception:  # interpreter shutdown
            pass

    def root(self) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        lib().glo_smt_root(self._h, _p(out))
        return out

    def set(self, key, value) -> np.ndarray:
        """tree.set(key, value) -> one SMT_PROOF_DTYPE record (the process proof)."""
        k, v = _a(key), _a(value)
        rec = np.zeros(1, dtype=SMT_PROOF_DTYPE)
        rc = lib().glo_smt_set(self._h, _p(k), _p(v), rec.ctypes.data)
        if rc:
            raise RuntimeError(f"smt set failed ({rc})")
        return rec[0]

    def find(self, key):
        k = _a(key)
        sib = np.zeros((256, 4), dtype=np.uint64)
        ns = C.c_uint32(0)
        nfk = np.zeros(4, dtype=np.uint64)
        val = np.zeros(4, dtype=np.uint64)
        old0 = C.c_uint32(0)
        found = lib().glo_smt_find(self._h, _p(k), _p(sib), C.byref(ns), _p(nfk), _p(val), C.byref(old0))
        return {"found": found == 1, "siblings": sib[: ns.value].copy(), "not_found_key": nfk, "value": val, "is_old0": bool(old0.value)}


def smt_verify_process_batch(proofs: np.ndarray) -> np.ndarray:
    proofs = np.ascontiguousarray(proofs, dtype=SMT_PROOF_DTYPE)
    status = np.zeros(proofs.shape[0], dtype=np.int32)
    lib().glo_smt_verify_process_batch(proofs.ctypes.data, proofs.shape[0], status.ctypes.data_as(C.POINTER(C.c_int32)))
    return status


def from_u128(x: int) -> np.ndarray:
    """GoldilocksHashOut::from_u128 (src/smt/goldilocks_poseidon/hash/mod.rs:254-267): four u32 limbs."""
    return np.array([(x >> (32 * i)) & 0xFFFFFFFF for i in range(4)], dtype=np.uint64)


def fri_layer_tree(values_ext, arity_bits: int, cap_height: int):
    v = _a(values_ext)
    ln = v.shape[0]
    nl = ln >> arity_bits
    leaves = np.zeros((nl, 2 << arity_bits), dtype=np.uint64)
    nd = 2 * (nl - (1 << cap_height))
    digests = np.zeros((nd, 4), dtype=np.uint64)
    cap = np.zeros((1 << cap_height, 4), dtype=np.uint64)
    rc = lib().glo_fri_layer_tree(_p(v), ln, arity_bits, cap_height, _p(leaves), _p(digests), _p(cap))
    if rc:
        raise ValueError(f"fri layer MerkleTree::new would panic ({rc})")
    return leaves, digests, cap


def fri_fold(coeffs_ext, arity_bits: int, beta) -> np.ndarray:
    cf, b = _a(coeffs_ext), _a(beta)
    out = np.zeros((cf.shape[0] >> arity_bits, 2), dtype=np.uint64)
    lib().glo_fri_fold(_p(cf), cf.shape[0], arity_bits, _p(b), _p(out))
    return out


def ext_coset_fft(a_ext, shift: int) -> np.ndarray:
    a = _a(a_ext).copy()
    lib().glo_ext_coset_fft(_p(a), int(a.shape[0]).bit_length() - 1, shift)
    return a


def pow_grind(state, pos: int, min_lz: int, start: int = 0, count: int = 1 << 22, out_pos: int = 7) -> int:
    s = _a(state)
    return int(lib().glo_pow_grind(_p(s), pos, out_pos, min_lz, start, count))


def reduce_polys_base(polys, alpha) -> np.ndarray:
    """ReducingFactor::reduce_polys_base: sum_j alpha^j * polys[j] -> [n][2]."""
    polys = [_a(p) for p in polys]
    n = polys[0].shape[0]
    arr = (u64p * len(polys))(*[_p(p) for p in polys])
    out = np.zeros((n, 2), dtype=np.uint64)
    a = _a(alpha)
    lib().glo_reduce_polys_base(arr, len(polys), n, _p(a), _p(out))
    return out


def divide_by_linear(poly_ext, z) -> np.ndarray:
    p = _a(poly_ext)
    out = np.zeros_like(p)
    zz = _a(z)
    lib().glo_divide_by_linear(_p(p), p.shape[0], _p(zz), _p(out))
    return out


def ext_poly_scale_add(acc_ext, scalar, add_ext) -> np.ndarray:
    acc = _a(acc_ext).copy()
    sc, ad = _a(scalar), _a(add_ext)
    lib().glo_ext_poly_scale_add(_p(acc), acc.shape[0], _p(sc), _p(ad))
    return acc


def eval_base_poly_at_ext(coeffs, point):
    c, pt = _a(coeffs), _a(point)
    out = np.zeros(2, dtype=np.uint64)
    lib().glo_eval_base_poly_at_ext(_p(c), c.shape[0], _p(pt), _p(out))
    return (int(out[0]), int(out[1]))


# ---- N3: permutation argument + compute_quotient_polys ----------------------------------------------------------
GATE_DTYPE = np.dtype([("kind", "<u4"), ("num_ops", "<u4"), ("selector_index", "<u4"), ("group_start", "<u4"), ("group_end", "<u4"),
                       ("reserved", "<u4")])
CIRCUIT_DTYPE = np.dtype([("degree_bits", "<u4"), ("num_wires", "<u4"), ("num_routed_wires", "<u4"), ("num_constants", "<u4"),
                          ("num_selectors", "<u4"), ("num_challenges", "<u4"), ("quotient_degree_factor", "<u4"), ("num_gates", "<u4")])
GATE_NOOP, GATE_CONSTANT, GATE_PUBLIC_INPUT, GATE_U32_INTERLEAVE, GATE_UNINTERLEAVE_TO_U32, GATE_UNINTERLEAVE_TO_B32 = range(6)


def permutation_zs(circuit, k_is, wires, sigmas, betas, gammas) -> np.ndarray:
    """Z and partial-product polynomials (values on the subgroup) [nch * (1 + num_prods)][n], prover column order."""
    cd = np.asarray(circuit, dtype=CIRCUIT_DTYPE).reshape(1)
    n = 1 << int(cd["degree_bits"][0])
    nch, R, deg = int(cd["num_challenges"][0]), int(cd["num_routed_wires"][0]), int(cd["quotient_degree_factor"][0])
    chunks = -(-R // deg)
    out = np.zeros((nch * chunks, n), dtype=np.uint64)
    w, sg, k, b, g = _a(wires), _a(sigmas), _a(k_is), _a(betas), _a(gammas)
    rc = lib().glo_permutation_zs(cd.ctypes.data, _p(k), _p(w), _p(sg), _p(b), _p(g), _p(out))
    if rc:
        raise ValueError("the grand product does not close: copy constraints violated or sigma is not a permutation")
    return out


def quotient_polys(circuit, gates, k_is, cs_leaves, wires_leaves, zs_leaves, rate_bits, pih, betas, gammas, alphas) -> np.ndarray:
    """compute_quotient_polys: [nch * quotient_degree_factor][n] coefficient chunks, from the row-major LDE leaves."""
    cd = np.asarray(circuit, dtype=CIRCUIT_DTYPE).reshape(1)
    gt = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    n = 1 << int(cd["degree_bits"][0])
    out = np.zeros((int(cd["num_challenges"][0]) * int(cd["quotient_degree_factor"][0]), n), dtype=np.uint64)
    a = [_a(x) for x in (k_is, cs_leaves, wires_leaves, zs_leaves, pih, betas, gammas, alphas)]
    rc = lib().glo_quotient_polys(cd.ctypes.data, gt.ctypes.data, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), rate_bits, _p(a[4]), _p(a[5]),
                                  _p(a[6]), _p(a[7]), _p(out))
    if rc:
        raise ValueError(f"compute_quotient_polys would panic (code {rc})")
    return out


def vanishing_at_point(circuit, gates, k_is, x, lcs, lw, lz, nz, pih, betas, gammas, alphas) -> np.ndarray:
    cd = np.asarray(circuit, dtype=CIRCUIT_DTYPE).reshape(1)
    gt = np.ascontiguousarray(gates, dtype=GATE_DTYPE)
    out = np.zeros(int(cd["num_challenges"][0]), dtype=np.uint64)
    a = [_a(v) for v in (k_is, lcs, lw, lz, nz, pih, betas, gammas, alphas)]
    lib().glo_vanishing_at_point(cd.ctypes.data, gt.ctypes.data, _p(a[0]), int(x), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]), _p(a[5]),
                                 _p(a[6]), _p(a[7]), _p(a[8]), _p(out))
    return out


def synthetic_values(c: int, n: int, seed: int = 0x706C6F6E6B7932, col0: int = 0) -> np.ndarray:
    """SURVEY 8d synthetic input: values[col][row] = splitmix64(seed ^ ((col<<32)+row)) mod p."""
    col = (np.arange(col0, col0 + c, dtype=np.uint64) << np.uint64(32))[:, None]
    row = np.arange(n, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        z = (np.uint64(seed) ^ (col + row)) + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return np.where(z >= np.uint64(P), z - np.uint64(P), z).astype(np.uint64)
