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
