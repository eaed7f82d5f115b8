import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure; oracle/gl_oracle.c restates plonky2 v0.1.4)."""
    from oracle import pyoracle

    pyoracle.lib()
    return pyoracle


@pytest.fixture(scope="session")
def glb():
    """The product package (directory name has a hyphen)."""
    mod = importlib.import_module("plonky2-lib_b200")
    build = importlib.import_module("plonky2-lib_b200.build")
    build.build()
    return mod


@pytest.fixture(scope="session")
def ctx(glb):
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    c = glb.Context.default()
    yield c


@pytest.fixture
def rng():
    return np.random.default_rng(0x706C6F6E6B7932)


P = 0xFFFFFFFF00000001


def rand_field(rng, shape, canonical=True):
    x = rng.integers(0, P, size=shape, dtype=np.uint64)
    return x


def adversarial_columns(n):
    """all-0, all-(p-1), bits, u32 limbs, p-1-row."""
    r = np.arange(n, dtype=np.uint64)
    return np.stack(
        [
            np.zeros(n, dtype=np.uint64),
            np.full(n, P - 1, dtype=np.uint64),
            r & np.uint64(1),
            (r * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF),
            np.uint64(P - 1) - r,
        ]
    )
