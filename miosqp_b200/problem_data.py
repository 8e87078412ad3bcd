"""
MIQP problem data in the layout the engine receives.

Mirrors what /root/reference/miosqp/data.py builds (add_bounds :5-33, Data :78-126): the integer-variable
bounds i_l <= x[i_idx] <= i_u become |i_idx| identity rows appended BELOW A (in i_idx order, unsorted), so a
B&B node only ever differs from the root in the last n_int entries of l and u.
"""
import numpy as np
import scipy.sparse as spa

try:    # the C routine behind csc_matrix.dot(vector): same arithmetic, none of the ~15 us of Python dispatch per call
    from scipy.sparse._sparsetools import csc_matvec as _csc_matvec
except ImportError:     # pragma: no cover
    _csc_matvec = None


class CscOperator(object):
    """y = M x for a fixed CSC matrix, bit-identical to `M.dot(x)` (it IS scipy's kernel, called directly).  The B&B
    replay evaluates P x and A x two to three times per node; at config-3 sizes the dispatch cost dominated."""

    def __init__(self, M):
        M = M.tocsc()
        self.M = M
        self.shape = M.shape
        self.fast = _csc_matvec is not None and M.dtype == np.float64 and M.indices.dtype == M.indptr.dtype

    def dot(self, x):
        if not (self.fast and x.dtype == np.float64 and x.ndim == 1 and x.flags.c_contiguous and x.shape[0] == self.shape[1]):
            return self.M.dot(x)
        y = np.zeros(self.shape[0])
        _csc_matvec(self.shape[0], self.shape[1], self.M.indptr, self.M.indices, self.M.data, x, y)
        return y


def add_bounds(i_idx, l_new, u_new, A, l, u):
    """Append the rows I[i_idx, :] with bounds (l_new, u_new) to l <= A x <= u."""
    n = A.shape[1]
    rows = spa.identity(n, format="csc")[i_idx, :]
    return spa.vstack([A, rows]).tocsc(), np.append(l, l_new), np.append(u, u_new)


class Data(object):
    """P, q and the EXTENDED A, l, u of one MIQP; n, m (original rows) and n_int."""

    def __init__(self, P, q, A, l, u, i_idx, i_l, i_u):
        self.m, self.n = A.shape
        self.n_int = len(i_idx)
        self.A, self.l, self.u = add_bounds(i_idx, i_l, i_u, A, l, u)
        self.P = P.tocsc()
        self.q = q
        self.P_op, self.A_op = CscOperator(self.P), CscOperator(self.A)
        self.i_idx, self.i_l, self.i_u = i_idx, i_l, i_u

    def compute_obj_val(self, x):
        """1/2 x'Px + q'x with the full symmetric P (data.py:99-103)."""
        return .5 * np.dot(x, self.P_op.dot(x)) + np.dot(self.q, x)

    def update_vectors(self, q=None, l=None, u=None):
        """Replace q and the ORIGINAL rows of l, u in place; dimension errors as data.py:105-126."""
        for name, vec, size in (("q", q, self.n), ("l", l, self.m), ("u", u, self.m)):
            if vec is not None and len(vec) != size:
                raise ValueError('Wrong %s dimension!' % name)
        if q is not None:
            self.q = q
        if l is not None:
            self.l[:self.m] = l
        if u is not None:
            self.u[:self.m] = u
