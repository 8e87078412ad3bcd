#!/usr/bin/env python
"""
Converts the reference's 49 input-only fixtures /root/reference/max_iter_examples/{28..76}.pickle (python-2
protocol-0 dumps of QP relaxations on which 2017-era OSQP hit max_iter; produced by the commented code in
/root/reference/miosqp/solver.py:93-109, read by /root/reference/extra/run_maxiter_problem.py:15-30) into one
compressed npz (tests/golden/max_iter_examples.npz): dense P (n x n), dense A (m x n), q, l, u, i_idx and the
stored settings (JSON).  The files are untrusted input, so a whitelisting Unpickler rebuilds only numpy arrays,
dtypes and scipy CSC containers.

    python tests/golden/make_pickle_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
SRC = "/root/reference/max_iter_examples"


from miosqp_b200.maxiter_problems import load_pickle as load   # whitelisting Unpickler lives in the package


def main():
    names = sorted(int(f.split(".")[0]) for f in os.listdir(SRC) if f.endswith(".pickle"))
    arrays, settings = {}, {}
    for k in names:
        p = load(os.path.join(SRC, "%d.pickle" % k))
        arrays["P_%d" % k] = p["P"].toarray()
        arrays["A_%d" % k] = p["A"].toarray()
        for v in ("q", "l", "u"):
            arrays["%s_%d" % (v, k)] = np.asarray(p[v], dtype=np.float64)
        arrays["i_idx_%d" % k] = np.asarray(p["i_idx"], dtype=np.int64)
        settings[str(k)] = p["settings"]
    arrays["names"] = np.array(names)
    arrays["settings_json"] = np.array(json.dumps(settings))
    out = os.path.join(HERE, "max_iter_examples.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out, os.path.getsize(out), "bytes;", len(names), "problems; settings of #28:", settings["28"], "#76:", settings["76"])


if __name__ == "__main__":
    main()
