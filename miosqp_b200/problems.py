"""
Workload generators for the BASELINE.json configs (synthetic inputs).

random_miqp follows the reference generator statement by statement
(/root/reference/examples/random_miqp/run_example.py:71-83); `sp.randn/sp.rand`
no longer exist in scipy 1.18, `np.random.randn/rand` draw from the same global
stream.
"""
import numpy as np
import scipy.sparse as spa

# reference settings, /root/reference/examples/random_miqp/run_example.py:98-116
RANDOM_MIQP_SETTINGS = {'eps_int_feas': 1e-03, 'max_iter_bb': 1000, 'tree_explor_rule': 1,
                        'branching_rule': 0, 'verbose': False, 'print_interval': 1}
RANDOM_MIQP_QP_SETTINGS = {'eps_abs': 1e-03, 'eps_rel': 1e-03, 'eps_prim_inf': 1e-04, 'verbose': False}


def random_miqp_draw(n, m, p, density):
    """One draw from the reference generator, consuming the global numpy RNG state."""
    i_idx = np.random.choice(np.arange(0, n), p, replace=False)
    Pt = spa.random(n, n, density=density)
    P = spa.csc_matrix(Pt.dot(Pt.T))
    q = np.random.randn(n)
    A = spa.random(m, n, density=density)
    u = 2 + np.random.rand(m)
    l = -2 + np.random.rand(m)
    i_l = np.zeros(p)
    i_u = np.ones(p)
    return dict(P=P, q=q, A=spa.csc_matrix(A), l=l, u=u, i_idx=i_idx, i_l=i_l, i_u=i_u)


def random_miqp(n, m, p, density, seed=1, count=1):
    """`count` consecutive draws after one np.random.seed(seed) (cfg 1: count=1; cfg 2: count=100)."""
    state = np.random.get_state()
    try:
        np.random.seed(seed)
        probs = [random_miqp_draw(n, m, p, density) for _ in range(count)]
    finally:
        np.random.set_state(state)
    return probs


CONFIGS = {
    "cfg1": dict(n=50, m=100, p=5, density=0.7, seed=1, count=1),
    "cfg2": dict(n=500, m=1000, p=50, density=0.7, seed=1, count=100),
    "cfg4": dict(n=2000, m=4000, p=200, density=0.05, seed=1, count=1),
}


def extend(pr):
    """(P, q, A_ext, l_ext, u_ext, i_idx) exactly as Data/add_bounds builds them
    (/root/reference/miosqp/data.py:5-33): integer-variable bound rows I[i_idx,:] appended below A."""
    n = pr['A'].shape[1]
    I_int = spa.identity(n, format='csc')[pr['i_idx'], :]
    A = spa.vstack([pr['A'], I_int]).tocsc()
    l = np.append(pr['l'], pr['i_l'])
    u = np.append(pr['u'], pr['i_u'])
    return pr['P'], pr['q'], A, l, u, np.asarray(pr['i_idx'])


def branched_nodes(l_ext, u_ext, n_int, count, rng, depth=3):
    """`count` synthetic B&B leaves of one instance: the root bounds with up to `depth` integer rows
    fixed to one side, as add_left/add_right would (/root/reference/miosqp/workspace.py:157-203).
    Node 0 is the root itself."""
    m = l_ext.shape[0]
    ls, us = [l_ext.copy()], [u_ext.copy()]
    for _ in range(count - 1):
        l = l_ext.copy(); u = u_ext.copy()
        rows = m - n_int + rng.choice(n_int, size=min(depth, n_int), replace=False)
        for r in rows:
            if rng.random() < 0.5:
                u[r] = np.floor(0.5 * (l_ext[r] + u_ext[r]))      # left child: u = floor(x)
            else:
                l[r] = np.ceil(0.5 * (l_ext[r] + u_ext[r]))       # right child: l = ceil(x)
        ls.append(l); us.append(u)
    return np.array(ls), np.array(us)
