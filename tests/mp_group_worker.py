"""Worker of tests/test_gpu_group.py::test_ranks_in_separate_processes: one rank of a gl_group per process."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as o  # noqa: E402  (the checker)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    glb = importlib.import_module("plonky2-lib_b200")
    tok = [glb.Group.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(tok, 0)
    ctx = glb.Context(local)
    g = glb.Group.from_token(ctx, rank, world, tok[0])
    for lg_n, c in [(12, 135), (10, 20)]:
        values = o.synthetic_values(c, 1 << lg_n)
        want = o.commit_from_values(values, 3, 4)
        b = g.commit(values, 3, 4)
        assert np.array_equal(b.cap, want["cap"]), "cap"
        mine = b.coeffs.any(axis=1)
        assert np.array_equal(b.coeffs[mine], want["coeffs"][mine]), "own coefficients"
        cnt = torch.tensor(mine.astype(np.int64))
        dist.all_reduce(cnt)
        assert bool((cnt == 1).all()), "every polynomial on exactly one rank"
        N = (1 << lg_n) << 3
        idx = [0, N // 2 - 1, N // 2, N - 1, 12345 % N, N // world, N // world - 1]
        rows, paths = b.open(idx)
        assert np.array_equal(rows, want["leaves"][idx]), "rows"
        for q, i in enumerate(idx):
            assert np.array_equal(paths[q], o.merkle_prove(want["digests"], N, 4, i)), "path"
        b.free()
    g.close()
    dist.barrier()
    dist.destroy_process_group()
    print("group worker ok", rank, flush=True)


if __name__ == "__main__":
    main()
