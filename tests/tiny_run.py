#!/usr/bin/env python
"""Smallest end-to-end run of the CUDA path (for compute-sanitizer / debugging): 40 x 24 tripole
grid, two evp calls of ndte subcycles, compared with the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from cice4_b200 import evp as E, synth  # noqa: E402
from helpers import STATE, cuda_steps, oracle_steps  # noqa: E402
from oracle import oracle as O  # noqa: E402

ndte = int(sys.argv[1]) if len(sys.argv) > 1 else 6
case = synth.make_case("om1deg", nx=40, ny=24, realistic=True)
st, f, strengths, _ = oracle_steps(O, case, nsteps=2, ndte=ndte)
lay = E.BlockLayout.single_block(40, 24)
dyn, out = cuda_steps(case, nsteps=2, strengths=strengths, ndte=ndte, use_graph=0)
ok = all(np.array_equal(E.merge_blocks(dyn.state[n], lay), st[n]) for n in STATE)
print("tiny_run:", "BIT-EXACT" if ok else "MISMATCH")
dyn.finalize()
sys.exit(0 if ok else 1)
