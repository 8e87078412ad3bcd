"""
bqp_setup_many (include/bqp.h): the host halves of many set-ups on several host threads must give
exactly the layouts of one-by-one set-ups, and fail as a whole.  The device half is the same to_device() as bqp_setup
(GPU test at the end, collected late: written after the round's last GPU run).
"""
import numpy as np
import pytest

from miosqp_b200 import engine, problems

QP = dict(problems.RANDOM_MIQP_QP_SETTINGS)


def _items(shapes):
    return [problems.extend(problems.random_miqp(n, m, p, d, seed=seed)[0]) for (n, m, p, d, seed) in shapes]


def test_host_halves_in_parallel_are_bit_identical():
    items = _items([(130, 200, 10, 0.7, 4), (50, 100, 5, 0.7, 1), (200, 300, 10, 0.05, 3), (130, 200, 10, 0.7, 5), (40, 40, 20, 0.7, 3)])
    seq = [engine.BatchedQP().setup(P, q, A, l, u, i_idx=i, host_only=True, **QP) for (P, q, A, l, u, i) in items]
    for threads in (1, 3, 0):
        par = engine.setup_many(items, host_only=True, threads=threads, **QP)
        rng = np.random.default_rng(1)
        for a, b in zip(seq, par):
            assert a.dims() == b.dims() and (a.n, a.m, a.n_int) == (b.n, b.m, b.n_int)
            rhs = rng.standard_normal(a.n + a.m)
            assert np.array_equal(a.debug_kkt_solve(rhs), b.debug_kkt_solve(rhs))
            Da, Ea, ca = a.scaling(); Db, Eb, cb = b.scaling()
            assert np.array_equal(Da, Db) and np.array_equal(Ea, Eb) and ca == cb
        for qp in par:
            qp.free()


def test_all_or_nothing():
    items = _items([(50, 100, 5, 0.7, 1), (50, 100, 5, 0.7, 2)])
    P, q, A, l, u, i = items[1]
    items[1] = (-P, q, A, l, u, i)                          # concave objective: the reduced KKT matrix is not positive definite
    with pytest.raises(ValueError):
        engine.setup_many(items, host_only=True, **QP)
    assert engine.setup_many([], host_only=True, **QP) == []
    with pytest.raises(ValueError):
        engine.setup_many(items[:1], host_only=True, adaptive_rho=True)


def test_no_device_no_setup():
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.BqpError):
        engine.setup_many(_items([(50, 100, 5, 0.7, 1)]), **QP)
