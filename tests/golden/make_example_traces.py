#!/usr/bin/env python
"""Record what the reference's case scripts do with the solver class (tests/refscripts.py) and
commit it as tests/golden/ref_example_traces.json.  Needs /root/reference; run here, offline."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests import refscripts  # noqa: E402

REF = os.environ.get("LBM3D_REFERENCE", "/root/reference")
# script -> iterations of its main loop that are run (the scripts loop 2001 .. 150001 times)
SCRIPTS = {"Single_phase/example_cavity.py": 60, "Single_phase/example_poiseuille_flow.py": 60,
           "Single_phase/example_porous_medium.py": 3}


def traces():
    return {rel: {"max_iter": n, "calls": refscripts.record(os.path.join(REF, rel), n)}
            for rel, n in SCRIPTS.items()}


if __name__ == "__main__":
    with open(os.path.join(HERE, "ref_example_traces.json"), "w") as fh:
        json.dump(traces(), fh, indent=1)
    print("wrote ref_example_traces.json")
