"""Test infrastructure: a small plonky2-style circuit over the gates the device quotient knows (NoopGate, ConstantGate,
PublicInputGate + the reference's U32InterleaveGate / UninterleaveToU32Gate / UninterleaveToB32Gate), with a satisfying
witness produced the way the reference's generators do (src/u32/gates/interleave_u32.rs:268-331,
uninterleave_to_u32.rs / uninterleave_to_b32.rs run_once), copy constraints between gate rows, the sigma polynomials of
the resulting permutation, and the selector / constant columns.  standard_recursion_config geometry: 135 wires, 80 routed,
2 challenges, quotient_degree_factor 8."""
import numpy as np

P = 0xFFFFFFFF00000001
NUM_WIRES, NUM_ROUTED, NUM_CHALLENGES, QDF = 135, 80, 2, 8
NOOP, CONSTANT, PUBLIC_INPUT, U32_INTERLEAVE, UNINTERLEAVE_TO_U32, UNINTERLEAVE_TO_B32 = range(6)
UNUSED_SELECTOR = 0xFFFFFFFF
# (kind, num_ops, selector_index, group_start, group_end): the gate's own index is its position
GATES = [(NOOP, 0, 0, 0, 3), (CONSTANT, 2, 0, 0, 3), (PUBLIC_INPUT, 0, 0, 0, 3),
         (U32_INTERLEAVE, 3, 1, 3, 6), (UNINTERLEAVE_TO_U32, 2, 1, 3, 6), (UNINTERLEAVE_TO_B32, 2, 1, 3, 6)]
NUM_SELECTORS, NUM_GATE_CONSTANTS = 2, 2


def k_is():
    return np.array([pow(7, j, P) for j in range(NUM_ROUTED)], dtype=np.uint64)   # get_unique_coset_shifts: g^j


def interleave(x):
    r = 0
    for b in range(32):
        r |= ((x >> b) & 1) << (2 * b)
    return r


def build(lg_n, seed=1, corrupt=None):
    """Returns dict(circuit, gates, k_is, constants [4][n], sigmas [80][n], wires [135][n], pih [4]).
    corrupt: None | "bit" (a decomposition bit flipped: a gate constraint fails) | "copy" (one side of a copy constraint
    changed, its own gate still satisfied where possible: the permutation argument fails)."""
    rng = np.random.default_rng(seed)
    n = 1 << lg_n
    wires = np.zeros((NUM_WIRES, n), dtype=np.uint64)
    consts = np.zeros((NUM_SELECTORS + NUM_GATE_CONSTANTS, n), dtype=np.uint64)
    gate_of_row = np.zeros(n, dtype=np.int64)
    pih = rng.integers(0, P, 4, dtype=np.uint64)
    copies = []                       # ((col, row), (col, row)) pairs that must hold equal values
    u32_outputs, interleaved_outputs = [], []   # positions holding a u32 / an interleaved 64-bit value
    gate_of_row[0] = 2
    wires[0:4, 0] = pih
    kinds = rng.choice([0, 1, 3, 3, 4, 5], size=n)
    for row in range(1, n):
        g = int(kinds[row])
        gate_of_row[row] = g
        if g == 1:
            c = rng.integers(0, 1 << 32, 2, dtype=np.uint64)
            consts[NUM_SELECTORS:NUM_SELECTORS + 2, row] = c
            wires[0:2, row] = c
            u32_outputs += [(0, row), (1, row)]
        elif g == 3:
            for op in range(3):
                if u32_outputs and rng.random() < 0.6:
                    src = u32_outputs[int(rng.integers(0, len(u32_outputs)))]
                    x = int(wires[src])
                    copies.append((src, (2 * op, row)))
                else:
                    x = int(rng.integers(0, 1 << 32))
                wires[2 * op, row] = x
                wires[2 * op + 1, row] = interleave(x)
                for b in range(32):
                    wires[6 + 32 * op + b, row] = (x >> (31 - b)) & 1       # big-endian decomposition
                interleaved_outputs.append((2 * op + 1, row))
                u32_outputs.append((2 * op, row))
        elif g in (4, 5):
            for op in range(2):
                if interleaved_outputs and rng.random() < 0.6:
                    src = interleaved_outputs[int(rng.integers(0, len(interleaved_outputs)))]
                    xi = int(wires[src])
                    copies.append((src, (3 * op, row)))
                else:
                    xi = int(rng.integers(0, 1 << 63))
                ev = od = 0
                for j in range(32):
                    shift = 2 * (32 - j - 1)
                    e, o = (xi >> (shift + 1)) & 1, (xi >> shift) & 1
                    wires[6 + 64 * op + 2 * j, row] = e
                    wires[6 + 64 * op + 2 * j + 1, row] = o
                    coeff = (1 << (32 - j - 1)) if g == 4 else (1 << shift)
                    ev += e * coeff
                    od += o * coeff
                wires[3 * op, row], wires[3 * op + 1, row], wires[3 * op + 2, row] = xi, ev, od
                if g == 4:
                    u32_outputs += [(3 * op + 1, row), (3 * op + 2, row)]
                else:
                    interleaved_outputs += [(3 * op + 1, row), (3 * op + 2, row)]
    for s in range(NUM_SELECTORS):
        for row in range(n):
            g = int(gate_of_row[row])
            consts[s, row] = g if GATES[g][2] == s else UNUSED_SELECTOR
    # sigma: union-find over routed positions, each class becomes one cycle
    parent = {}

    def find(a):
        while parent.setdefault(a, a) != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in copies:
        parent[find(a)] = find(b)
    classes = {}
    for a in list(parent):
        classes.setdefault(find(a), []).append(a)
    lg = lg_n
    w = pow(pow(7, (P - 1) >> 32, P), 1 << (32 - lg), P)
    sub = np.empty(n, dtype=object)
    cur = 1
    for i in range(n):
        sub[i] = cur
        cur = cur * w % P
    ks = [int(v) for v in k_is()]
    sig = np.empty((NUM_ROUTED, n), dtype=object)
    for j in range(NUM_ROUTED):
        sig[j, :] = [ks[j] * int(sub[i]) % P for i in range(n)]
    for members in classes.values():
        members = sorted(members)
        for t, (col, row) in enumerate(members):
            tc, tr = members[(t + 1) % len(members)]
            sig[col, row] = ks[tc] * int(sub[tr]) % P
    sigmas = sig.astype(np.uint64)
    if corrupt == "bit":
        row = int(np.nonzero(gate_of_row == 3)[0][0])
        wires[6 + 5, row] ^= np.uint64(1)
    elif corrupt == "copy":
        (src, dst) = copies[len(copies) // 2]
        wires[src] = (int(wires[src]) + 1) % P          # the producer row now breaks its gate AND the copy
    circuit = (lg_n, NUM_WIRES, NUM_ROUTED, NUM_SELECTORS + NUM_GATE_CONSTANTS, NUM_SELECTORS, NUM_CHALLENGES, QDF, len(GATES))
    gates = [(k, ops, s, g0, g1, 0) for (k, ops, s, g0, g1) in GATES]
    return {"circuit": circuit, "gates": gates, "k_is": k_is(), "constants": consts, "sigmas": sigmas, "wires": wires, "pih": pih,
            "num_copies": len(copies)}
