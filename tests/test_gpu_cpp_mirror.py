"""Builds and runs the C++ host mirror test (include/plonky2_b200.hpp over the C ABI) on the GPU box."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_cpp_host_mirror(glb, ctx, tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    libdir = os.path.join(ROOT, "plonky2-lib_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-L", libdir, "-lgl_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host_mirror_test ok" in out.stdout


def test_cpp_host_mirror_compiles(glb, tmp_path):
    """CPU-side: the header compiles and links against the C ABI (no compute call is made)."""
    for name in ("host_mirror_test", "fri_mirror_test"):
        obj = str(tmp_path / (name + ".o"))
        subprocess.check_call(["g++", "-std=c++17", "-c", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", obj])


def _flatten_proof(proof, last_challenge):
    out = []
    for cap in proof["commit_phase_merkle_caps"]:
        out += [int(x) for x in cap.reshape(-1)]
    out += [int(x) for x in proof["final_poly"].reshape(-1)]
    out.append(int(proof["pow_witness"]))
    for r in proof["query_round_proofs"]:
        out.append(int(r["x_index"]))
        for row, path in r["initial_trees_proof"]:
            out += [int(x) for x in row] + [int(x) for x in path.reshape(-1)]
        for s in r["steps"]:
            out += [int(x) for x in s["evals"].reshape(-1)] + [int(x) for x in s["merkle_proof"].reshape(-1)]
    out.append(int(last_challenge))
    return out


@pytest.mark.gpu
def test_cpp_prove_openings_matches_oracle(glb, ctx, oracle, tmp_path):
    """The C++ host mirror's prove_openings (Challenger, fri_proof_resident) against the oracle's, field by field."""
    from oracle import fri_oracle as fo

    exe, dump = str(tmp_path / "fri_mirror_test"), str(tmp_path / "proof.txt")
    libdir = os.path.join(ROOT, "plonky2-lib_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "fri_mirror_test.cpp"), "-L", libdir, "-lgl_b200",
                           f"-Wl,-rpath,{libdir}", "-o", exe])
    out = subprocess.run([exe, dump], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    got = [int(x) for x in open(dump).read().split()]
    degree_bits, cols = 8, (4, 9, 3)
    n = 1 << degree_bits
    polys, trees = [], []
    for k, c in enumerate(cols):
        res = oracle.commit_from_values(oracle.synthetic_values(c, n, seed=700 + k), 3, 4)
        polys.append(res["coeffs"])
        trees.append(fo.MerkleTree(res["leaves"], 4))
    zeta = (0x0123456789ABCDEF, 0x0FEDCBA987654321)
    g = oracle.lib().glo_primitive_root_of_unity(degree_bits)
    instance = [(zeta, [(oi, pi) for oi, c in enumerate(cols) for pi in range(c)]), (fo.ext_scalar(zeta, g), [(2, 0), (2, 1)])]
    och = fo.Challenger()
    for t in trees:
        och.observe_cap(t.cap)
    want = fo.prove_openings(polys, trees, instance, och, degree_bits, 3, 4, 10, 28)
    assert got == _flatten_proof(want, och.get_challenge())
