#!/usr/bin/env python3
"""Search the trie order of engine 4 (tools/gen_sym.py): a permutation of the K-orbits of the 36 variables.

The order decides which variables become the dynamic row pair / the static child / the static leaves of every monomial, hence
the number of FP64 instructions per trimer (children and rows cost 4 |K| each) and the size of the generated code (distinct
row shapes).  Measured on B200 (experiments/README.md section 9): k_trimers time ~ 0.15 ms per 1000 FP64 instructions + 0.043 ms
per 1000 leaf pairs of code, so the objective is J = instructions + 0.28 * code.  Random-restart hill climbing over swaps and
moves; prints the best orders found.  usage: search_sym_order.py [seconds] [seed]

The objective only ranks candidates: the instruction cache makes the landscape bumpy (an order with 70.4 k instructions but 15.5 k leaf
pairs of code runs at 17.6 ms, one with 6.7 k leaf pairs of code and 77.6 k instructions at 32.8 ms with the split row order), so the
two or three best local optima must be BUILT AND MEASURED (csrc/Makefile: SYMFLAGS="--order ... --repsel ..." OUT=variants/...,
scripts/ab_phases.py) before one becomes DEFAULT_ORDER in gen_sym.py."""
import sys
import time

import numpy as np

import gen_poly as G
import gen_sym as S


W_CODE = float(__import__("os").environ.get("SYM_W_CODE", "0.6"))      # refitted after the first searched order was measured: 0.32 ms per 1000 instructions, 0.19 ms per 1000 leaf pairs of code


def J(mono, order, repsel):
    rows, shapes, vps = S.build(mono, "z", True, repsel, order=list(order))
    if S.PAIRS:
        rows, shapes = S.merge_pairs(rows)
    nfp, code, nrows, nshapes = S.stats(rows, shapes, len(vps))
    return nfp + W_CODE * code, (nfp, code, nrows, nshapes)


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 300.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    mono = np.load(G.DATA)["mono3b"]
    n = 12
    t0 = time.time()
    best = None
    while time.time() - t0 < budget:
        order = list(rng.permutation(n))
        repsel = "min" if rng.random() < 0.5 else "max"
        cur, info = J(mono, order, repsel)
        improved = True
        while improved and time.time() - t0 < budget:
            improved = False
            for i in range(n):
                for j in range(n):
                    if i == j:
                        continue
                    cand = order[:]
                    cand.insert(j, cand.pop(i))
                    c, inf = J(mono, cand, repsel)
                    if c < cur - 1e-9:
                        order, cur, info, improved = cand, c, inf, True
            other = "min" if repsel == "max" else "max"
            c, inf = J(mono, order, other)
            if c < cur - 1e-9:
                repsel, cur, info, improved = other, c, inf, True
        print("local optimum J = %.0f  instr %d code %d rows %d shapes %d  order %s repsel %s" % ((cur,) + info + (",".join(map(str, order)), repsel)), flush=True)
        if best is None or cur < best[0]:
            best = (cur, info, order, repsel)
    print("BEST J = %.0f  instr %d code %d rows %d shapes %d  --order %s --repsel %s" % ((best[0],) + best[1] + (",".join(map(str, best[2])), best[3])))


if __name__ == "__main__":
    main()
