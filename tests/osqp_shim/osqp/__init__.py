"""
TEST-ONLY stand-in for the PyPI `osqp` module (absent from this image), so the
UNMODIFIED reference package at /root/reference/miosqp can be imported and run.
It forwards the six-call surface the reference uses (workspace.py:63-68,
node.py:102-125, solver.py:185, osqp.constant) to a selectable backend:

    osqp.set_backend("oracle")  -> oracle/oracle.py      (CPU oracle, default)
    osqp.set_backend("b200")    -> miosqp_b200.osqp_compat (the CUDA engine, B=1 batches)

Put this directory on sys.path BEFORE importing the reference `miosqp`.
"""
import importlib

_backend = "oracle"


def set_backend(name):
    global _backend
    assert name in ("oracle", "b200")
    _backend = name


def _mod():
    if _backend == "oracle":
        return importlib.import_module("oracle.oracle")
    return importlib.import_module("miosqp_b200.osqp_compat")


def constant(name):
    return _mod().constant(name)


def OSQP():
    return _mod().OSQP()
