"""Multi-GPU plan of one sharded commit: who owns what, and the two exchanges.

One process per GPU (torch.distributed; NCCL on GPUs, gloo in the CPU tests).  A commit of c polynomials
of n coefficients with N = n * 2^rate_bits leaves is split over `world` ranks:

  * the IFFT by column slice (`column_slice`), followed by ONE all-gather of the coefficients;
  * the LDE + Merkle tree by leaf block: rank r owns leaves [r * N / world, (r + 1) * N / world)
    = LDE cosets k with bitrev_r(k) in that range = whole top-level subtrees and their cap entries
    (`gl_ctx_set_shard`); no exchange;
  * the cap by ONE all-gather of 2^cap_height / world digests per rank;
  * query openings are answered by `owner_of_leaf`.

Batches of independent items (SMT process proofs, P7; Poseidon batches) are cut evenly by `batch_range` with no
exchange at all; `slice_proof_batch` rebases the sibling offsets of one rank's share.
"""
from __future__ import annotations

from typing import Tuple


def column_slice(rank: int, world: int, cols: int) -> Tuple[int, int, int]:
    """(first column, one-past-last column, padded slice width) of the IFFT slice of `rank`."""
    per = (cols + world - 1) // world
    return min(rank * per, cols), min((rank + 1) * per, cols), per


def round_blocks(rank: int, world: int, cols: int, target_rounds: int = 5):
    """Pipelined plan: the columns are cut into rounds of world * G consecutive columns; inside round j rank r
    inverse-transforms the G columns [j * world * G + r * G, + G).  An all-gather of the G columns of every rank
    then yields the round's world * G coefficient columns contiguously, so the LDE of round j can run while
    round j + 1 is on the wire.  Returns (G, rounds, [(first, last) of this rank's real columns per round])."""
    G = max(1, -(-cols // (world * target_rounds)))
    rounds = -(-cols // (world * G))
    mine = []
    for j in range(rounds):
        first = j * world * G + rank * G
        mine.append((min(first, cols), min(first + G, cols)))
    return G, rounds, mine


def leaf_range(rank: int, world: int, num_leaves: int) -> Tuple[int, int]:
    per = num_leaves // world
    return rank * per, (rank + 1) * per


def owner_of_leaf(leaf_index: int, world: int, num_leaves: int) -> int:
    return leaf_index // (num_leaves // world)


def cap_range(rank: int, world: int, cap_height: int) -> Tuple[int, int]:
    per = (1 << cap_height) // world
    return rank * per, (rank + 1) * per


def check_shardable(world: int, rate_bits: int, cap_height: int) -> None:
    if world & (world - 1) or world > (1 << rate_bits) or world > (1 << cap_height):
        raise ValueError("the rank count must be a power of two dividing 2^rate_bits and 2^cap_height")


def all_gather_coefficients(dist, slice_buf, gathered, cols: int):
    """slice_buf [per][n] (this rank's IFFT output, zero padded) -> gathered [world * per][n]; the first
    `cols` rows of `gathered` are the coefficients of every column, in order, on every rank."""
    dist.all_gather_into_tensor(gathered, slice_buf)
    return gathered[:cols]


def all_gather_cap(dist, local_cap, cap_all):
    """local_cap [2^cap_height / world][4] -> cap_all [2^cap_height][4] (MerkleCap) on every rank."""
    dist.all_gather_into_tensor(cap_all, local_cap)
    return cap_all


def batch_range(rank: int, world: int, m: int) -> Tuple[int, int]:
    """[lo, hi) of the items of `rank` when m independent items are split as evenly as possible."""
    base, extra = divmod(m, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def slice_proof_batch(headers, sib_pool, sib_off, lo: int, hi: int):
    """The share [lo, hi) of a batch of SparseMerkleProcessProofs laid out as gl_smt_verify_process_batch takes it
    (headers [m], sib_pool [total][4], sib_off [m + 1]): the same three arrays for that share, offsets rebased."""
    first, last = int(sib_off[lo]), int(sib_off[hi])
    return headers[lo:hi], sib_pool[first:last], sib_off[lo:hi + 1] - sib_off[lo]
