import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import binding as O
from outfit_b200 import IODParams, OutfitB200, SolverType, synth
ctx = OutfitB200(0)
for kind in (0, 1):
    rv, t0, t1 = synth.make_propagation_states(200_000, seed=7 + kind)
    st = SolverType(kind=kind)
    out, status = ctx.propagate_universal(rv, t0, t1, st)
    want, wst = O.propagate_universal_batch(rv, t0, t1, kind, st.convergency, 0)
    bad = np.where(status != wst)[0]
    print("kind", kind, "mismatch", len(bad), "gpu hist", np.unique(status, return_counts=True), "cpu hist", np.unique(wst, return_counts=True))
    for i in bad[:5]:
        r, v = rv[0:3, i], rv[3:6, i]
        mu = 0.01720209895 ** 2
        alpha = (v @ v - 2 * mu / np.linalg.norm(r)) / mu
        print("   i", i, "gpu", status[i], "cpu", wst[i], "alpha", alpha, "dt", t1[i] - t0[i], "r0", np.linalg.norm(r))
