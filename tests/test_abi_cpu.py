"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, exports every
symbol include/gl_b200.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = header_symbols()
    for must in ["gl_commit_from_values", "gl_commit_from_coeffs", "gl_merkle_build", "gl_poseidon_two_to_one_batch",
                 "gl_poseidon_hash_no_pad_batch", "gl_smt_verify_process_batch", "gl_fri_layer_tree", "gl_pow_grind"]:
        assert must in syms


def test_library_exports_every_declared_symbol(glb):
    lib = ctypes.CDLL(glb._native.LIB_PATH)
    for s in header_symbols():
        assert hasattr(lib, s), f"{s} declared in include/gl_b200.h but not exported"


def test_binding_covers_every_declared_symbol(glb):
    assert sorted(glb._native.SIGNATURES) == header_symbols()


def test_no_cpu_fallback(glb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    with pytest.raises(glb.GlPanic) as e:
        glb.Context(0)
    assert "no CUDA device" in str(e.value)


def test_product_does_not_reference_oracle():
    """The product tree must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "plonky2-lib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "gl_oracle" not in text and "libgl_oracle" not in text, f


def test_fri_config_matches_standard_recursion_config(glb):
    cfg = glb.CircuitConfig.standard_recursion_config()
    assert (cfg.num_wires, cfg.num_routed_wires, cfg.num_challenges) == (135, 80, 2)
    f = cfg.fri_config
    assert (f.rate_bits, f.cap_height, f.proof_of_work_bits, f.num_query_rounds) == (3, 4, 16, 28)
    assert f.reduction_strategy.reduction_arity_bits(20, 3, 4) == [4, 4, 4, 4]
    assert glb.CircuitConfig.standard_ecc_config().num_wires == 136
    assert glb.CircuitConfig.wide_ecc_config().num_wires == 234


def test_hash_out_text_form_matches_the_reference():
    """WrappedHashOut Display / Serialize / Deserialize (src/smt/goldilocks_poseidon/hash/mod.rs:62-78, 84-137) and the
    zkdsa public-input JSON (src/zkdsa/circuits/mod.rs:136-153): host-side formatting, no device needed."""
    import importlib

    import numpy as np

    host = importlib.import_module("plonky2-lib_b200.host")
    one = np.array([1, 0, 0, 0], dtype=np.uint64)
    assert host.hash_out_to_hex(one) == "0x0000000000000000000000000000000000000000000000000000000000000001"
    assert host.hash_out_from_hex("0x01").tolist() == one.tolist()
    kat = np.array([4330397376401421145, 14124799381142128323, 8742572140681234676, 14345658006221440202], dtype=np.uint64)
    text = "0xc71603f33a1144ca7953db0ab48808f4c4055e3364a246c33c18a9786cb0b359"
    assert host.hash_out_to_hex(kat) == text and host.hash_out_from_hex(text).tolist() == kat.tolist()
    rng = np.random.default_rng(5)
    for _ in range(20):
        h = rng.integers(0, 2**64, 4, dtype=np.uint64)
        s = host.hash_out_to_hex(h)
        assert len(s) == 66 and host.hash_out_from_hex(s).tolist() == h.tolist()
    import pytest

    for bad in ("01", "0x1", "0xzz", "0x" + "00" * 33):
        with pytest.raises(ValueError):
            host.hash_out_from_hex(bad)


def test_smt_proof_json_forms_match_the_reference_layout():
    """serde_json of SparseMerkleProcessProof / SparseMerkleInclusionProof over GoldilocksHashOut
    (src/smt/proof/process.rs:12-23,53-59, src/smt/proof/inclusion.rs:5-33,62-80): fields in declaration order, hashes as
    0x-hex strings, the role as its variant name.  Host-side formatting only; uses the golden fixture for real proofs."""
    import importlib
    import json
    import os

    import numpy as np
    import pytest

    host = importlib.import_module("plonky2-lib_b200.host")
    g = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "smt_sets.json")))
    un = lambda xs: np.array([int(x, 16) for x in xs], dtype=np.uint64)  # noqa: E731
    m = len(g["calls"])
    hd = np.zeros(m, dtype=host.SMT_HDR_DTYPE)
    off = np.zeros(m + 1, dtype=np.uint64)
    sibs = []
    for t, c in enumerate(g["calls"]):
        for f in ("old_root", "old_key", "old_value", "new_root", "new_key", "new_value"):
            hd[f][t] = un(c[f])
        hd["is_old0"][t], hd["fnc"][t] = c["is_old0"], c["fnc"]
        sibs.append(un(c["siblings"]).reshape(-1, 4))
        off[t + 1] = off[t] + np.uint64(sibs[-1].shape[0])
    pool = np.concatenate(sibs)
    texts = host.smt_process_proofs_to_json(hd, pool, off)
    first = json.loads(texts[0])
    assert list(first) == ["old_root", "old_key", "old_value", "new_root", "new_key", "new_value", "siblings", "is_old0", "fnc"]
    assert first["fnc"] == "ProcessInsert" and first["is_old0"] is True and first["siblings"] == []
    assert first["new_key"] == "0x0000000000000000000000000000000000000000000000000000000000000001"   # key 1
    assert first["old_root"] == "0x" + "00" * 32 and " " not in texts[0]
    assert {json.loads(x)["fnc"] for x in texts} == set(host.PROCESS_ROLES)
    hd2, pool2, off2 = host.smt_process_proofs_from_json(texts)
    assert np.array_equal(hd2, hd) and np.array_equal(pool2, pool) and np.array_equal(off2, off)
    with pytest.raises(ValueError):
        host.smt_process_proofs_from_json([texts[0].replace("ProcessInsert", "ProcessUpsert")])
    # inclusion proofs: the shape of the reference's test_serialize_merkle_proof
    inc = np.zeros(1, dtype=host.SMT_INCLUSION_DTYPE)
    inc["root"][0][0], inc["key"][0][0], inc["value"][0][0] = 1, 2, 3
    inc["not_found_key"][0][0], inc["not_found_value"][0][0], inc["found"][0] = 5, 6, 1
    sib = np.array([[4, 0, 0, 0]], dtype=np.uint64)
    text = host.smt_inclusion_proofs_to_json(inc, sib, np.array([0, 1], dtype=np.uint64))[0]
    obj = json.loads(text)
    assert list(obj) == ["root", "found", "key", "value", "not_found_key", "not_found_value", "siblings", "is_old0"]
    assert obj["found"] is True and obj["is_old0"] is False and obj["siblings"] == ["0x" + "00" * 31 + "04"]
    back = host.smt_inclusion_proofs_from_json([text])
    assert np.array_equal(back[0], inc) and np.array_equal(back[1], sib) and back[2].tolist() == [0, 1]


def test_every_entry_point_has_its_reference_side_binding_documented():
    """INTEGRATION.md shows the Rust `extern "C"` declaration a maintainer of the fork would add for every function of
    include/gl_b200.h (instrumentation-only entry points may be named in a comment instead)."""
    import os
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "gl_b200.h")).read()
    integ = open(os.path.join(root, "INTEGRATION.md")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(gl_[a-z0-9_]+)\s*\(", header))
    bound = set(re.findall(r"pub fn (gl_[a-z0-9_]+)", integ))
    missing = {n for n in declared - bound if n not in integ}
    assert not missing, sorted(missing)
    # the crate a machine with cargo compiles unchanged: rust/plonky2_gl_b200_sys declares every entry point, and nothing else
    crate = open(os.path.join(root, "rust", "plonky2_gl_b200_sys", "src", "lib.rs")).read()
    in_crate = set(re.findall(r"pub fn (gl_[a-z0-9_]+)", crate))
    assert in_crate == declared, sorted(in_crate ^ declared)
    for f in ("Cargo.toml", "build.rs"):
        assert os.path.exists(os.path.join(root, "rust", "plonky2_gl_b200_sys", f))


def test_cpp_mirror_compiles_and_links_without_a_gpu(tmp_path):
    """include/plonky2_b200.hpp (the host side above the C ABI where the reference is compiled code) and the two C++ test
    programs build against the in-tree library with plain g++: every entry point the mirror uses exists with the signature
    it expects.  Nothing is run here (no device)."""
    import os
    import shutil
    import subprocess

    import pytest

    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "plonky2-lib_b200")
    if not os.path.exists(os.path.join(libdir, "libgl_b200.so")):
        pytest.skip("library not built")
    for name in ("host_mirror_test", "fri_mirror_test"):
        out = subprocess.run(["g++", "-std=c++17", "-O0", "-Wall", "-I", os.path.join(root, "include"),
                              os.path.join(root, "tests", "cpp", name + ".cpp"), "-L", libdir, "-lgl_b200", f"-Wl,-rpath,{libdir}",
                              "-o", str(tmp_path / name)], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr


def test_poseidon_linear_layer_tables_on_the_cpu(tmp_path):
    """tests/cpp/poseidon_tables_test.cpp: the constant tables the FP64 linear layers add (signed S-box offsets included)
    replayed with exact integers in the device's own schedule: every accumulator the fold sees stays in [0, 2^50) and the
    permutation equals its definition."""
    import os
    import shutil
    import subprocess

    import pytest

    if not shutil.which("g++"):
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "poseidon_tables_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", os.path.join(root, "tests", "cpp", "poseidon_tables_test.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "poseidon_tables_test ok" in out.stdout, out.stdout + out.stderr
