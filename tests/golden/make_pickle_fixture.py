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
import pickle

import numpy as np
import scipy.sparse as spa

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/max_iter_examples"


class _Obj(object):
    """Placeholder for copy_reg._reconstructor targets (scipy.sparse.csc.csc_matrix instances)."""


def _reconstructor(cls, base, state):
    return _Obj()


class SafeUnpickler(pickle.Unpickler):
    ALLOWED = {
        ("copy_reg", "_reconstructor"): _reconstructor,
        ("__builtin__", "object"): object,
        ("scipy.sparse.csc", "csc_matrix"): _Obj,
        ("numpy.core.multiarray", "_reconstruct"): None,     # resolved lazily in find_class (numpy._core on numpy 2)
        ("numpy", "ndarray"): np.ndarray,
        ("numpy", "dtype"): np.dtype,
        ("numpy.core.multiarray", "scalar"): None,
    }

    def find_class(self, module, name):
        key = (module, name)
        if key == ("numpy.core.multiarray", "_reconstruct"):
            from numpy._core import multiarray
            return multiarray._reconstruct
        if key == ("numpy.core.multiarray", "scalar"):
            from numpy._core import multiarray
            return multiarray.scalar
        if key in self.ALLOWED and self.ALLOWED[key] is not None:
            return self.ALLOWED[key]
        raise pickle.UnpicklingError("blocked global %s.%s" % key)


def load(path):
    with open(path, "rb") as f:
        d = SafeUnpickler(f, encoding="latin1").load()
    out = {}
    for k, v in d.items():
        if isinstance(v, _Obj):
            st = v.__dict__
            shape = tuple(int(s) for s in st["_shape"])
            v = spa.csc_matrix((st["data"], st["indices"], st["indptr"]), shape=shape)
        out[k] = v
    return out


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
        settings[str(k)] = {kk: (vv.item() if hasattr(vv, "item") else vv) for kk, vv in p["settings"].items()}
    arrays["names"] = np.array(names)
    arrays["settings_json"] = np.array(json.dumps(settings))
    out = os.path.join(HERE, "max_iter_examples.npz")
    np.savez_compressed(out, **arrays)
    print("wrote", out, os.path.getsize(out), "bytes;", len(names), "problems; settings of #28:", settings["28"], "#76:", settings["76"])


if __name__ == "__main__":
    main()
