"""
Readers for the reference's "max_iter" QP relaxations (BASELINE config 5).

/root/reference/max_iter_examples/{28..76}.pickle are python-2 protocol-0 dumps with keys P, q, A, l, u, i_idx,
settings (written by the commented code in /root/reference/miosqp/solver.py:93-109, read by
/root/reference/extra/run_maxiter_problem.py:15-30).  They are untrusted files, so `load_pickle` uses a
whitelisting Unpickler that can only rebuild numpy arrays, dtypes and scipy CSC containers.  `load_npz` reads the
converted bundle the tests ship (tests/golden/max_iter_examples.npz, made by tests/golden/make_pickle_fixture.py).
"""
import json
import pickle

import numpy as np
import scipy.sparse as spa


class _Obj(object):
    """Placeholder for copy_reg._reconstructor targets (scipy.sparse.csc.csc_matrix instances)."""


def _reconstructor(cls, base, state):
    return _Obj()


class SafeUnpickler(pickle.Unpickler):
    ALLOWED = {
        ("copy_reg", "_reconstructor"): _reconstructor,
        ("__builtin__", "object"): object,
        ("scipy.sparse.csc", "csc_matrix"): _Obj,
        ("numpy", "ndarray"): np.ndarray,
        ("numpy", "dtype"): np.dtype,
    }
    NUMPY_CORE = ("_reconstruct", "scalar")        # numpy.core.multiarray on numpy 1, numpy._core.multiarray on numpy 2

    def find_class(self, module, name):
        if module == "numpy.core.multiarray" and name in self.NUMPY_CORE:
            from numpy._core import multiarray
            return getattr(multiarray, name)
        if (module, name) in self.ALLOWED:
            return self.ALLOWED[(module, name)]
        raise pickle.UnpicklingError("blocked global %s.%s" % (module, name))


def load_pickle(path):
    """One reference pickle -> dict(P csc, q, A csc, l, u, i_idx, settings)."""
    with open(path, "rb") as f:
        d = SafeUnpickler(f, encoding="latin1").load()
    out = {}
    for k, v in d.items():
        if isinstance(v, _Obj):
            st = v.__dict__
            shape = tuple(int(s) for s in st["_shape"])
            v = spa.csc_matrix((st["data"], st["indices"], st["indptr"]), shape=shape)
        out[k] = v
    out["settings"] = {kk: (vv.item() if hasattr(vv, "item") else vv) for kk, vv in out["settings"].items()}
    return out


def load_npz(path):
    """The converted bundle -> list of dict(name, P csc, q, A csc, l, u, i_idx, settings), in file order."""
    z = np.load(path)
    settings = json.loads(str(z["settings_json"]))
    probs = []
    for k in z["names"]:
        k = int(k)
        probs.append(dict(name=k, P=spa.csc_matrix(z["P_%d" % k]), A=spa.csc_matrix(z["A_%d" % k]), q=z["q_%d" % k],
                          l=z["l_%d" % k], u=z["u_%d" % k], i_idx=z["i_idx_%d" % k], settings=settings[str(k)]))
    return probs
