"""production vs oracle on a two-phase case, step by step (where does the difference come from?)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import cases2p
from oracle.cref import RefTwoPhaseC
name = sys.argv[1] if len(sys.argv) > 1 else "case_bcs"
case = getattr(cases2p, name)()
fl = case.solid == 0
o = case.make_oracle(RefTwoPhaseC)
lb = case.make_solver()
done = 0
for upto in (1, 2, 5, 10, 20, 35, 50):
    o.run(upto - done); lb.run(upto - done); done = upto
    out = []
    for n in ("F", "rho", "psi", "rho_r", "rho_b", "v"):
        a = getattr(lb, n).to_numpy(); b = getattr(o, n)
        d = np.abs(a.astype(np.float64) - b)
        d[~fl] = 0
        k = np.unravel_index(np.argmax(d), d.shape)
        out.append("%s %.2e@%s" % (n, d.max() / np.abs(b[fl]).max(), k[:3]))
    print(upto, " | ".join(out), flush=True)
