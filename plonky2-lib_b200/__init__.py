"""plonky2-lib_b200: the B200-native Goldilocks NTT/LDE + Poseidon Merkle commit core underneath
Orbiter-Finance/Plonky2-lib's `data.prove(pw)` / `builder.build::<C>()` / `PoseidonHash::*` calls.

The product is the CUDA library `libgl_b200.so` (C ABI: include/gl_b200.h, sources: csrc/).  This
package is the thin host-side mirror of the plonky2 plugin surface used by tests and bench.py.
The directory name contains a hyphen, so import it with
`importlib.import_module("plonky2-lib_b200")`.
"""
from . import _native  # noqa: F401
from .host import *  # noqa: F401,F403
from .host import (  # noqa: F401
    CircuitConfig, Context, FriConfig, Group, ShardedPolynomialBatch, FriReductionStrategy, GlPanic, MerkleTree, PolynomialBatch, PoseidonHash,
    PoseidonNodeHash, coset_fft, coset_ifft, fft, fri_fold, fri_layer_tree, fri_proof_of_work, ifft, pinned_empty,
    smt_check_process_proofs,
)
