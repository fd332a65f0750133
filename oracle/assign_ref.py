"""Exact balanced assignment on the CPU.  TEST INFRASTRUCTURE ONLY.

The reference's arithmetic for this step lives in cvxopt.glpk.ilp -> GLPK (third-party, not
vendored, version unpinned; see oracle/assign_oracle.c).  Three independent checkers:

  solve_exact          our C shortest-augmenting-path restatement (int64, O(n^3))
  solve_scipy          scipy.optimize.linear_sum_assignment
  lp_relaxation        scipy.optimize.linprog(HiGHS) on the reference's own 2n x n^2 equality
                       layout (solver.py:15-25: arr[i][n*i+j] = 1, arr[n+i][n*j+i] = 1, b = 1),
                       built sparse; this is glpk.mod:9's `var x >= 0` model and anchors the
                       "LP-relaxation objective within 1e-9 relative" clause of north_star.

The optimal objective is unique; the optimal x is not (KAT A1 has two optima), so parity is on
the objective plus feasibility of x.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _clib


def solve_exact(cost):
    c = np.ascontiguousarray(np.asarray(cost, dtype=np.int32))
    n = c.shape[0]
    assert c.shape == (n, n)
    col = np.zeros(max(n, 1), dtype=np.int32)
    obj = ctypes.c_int64(0)
    rc = _clib.lib().assign_oracle(c.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(n),
                                   col.ctypes.data_as(ctypes.c_void_p), ctypes.byref(obj))
    if rc != 0:
        raise RuntimeError("assign_oracle rc=%d" % rc)
    return obj.value, col[:n].copy()


def solve_scipy(cost):
    from scipy.optimize import linear_sum_assignment
    c = np.asarray(cost, dtype=np.int64)
    r, col = linear_sum_assignment(c)
    return int(c[r, col].sum()), col.astype(np.int32)


def x_from_cols(col_of_row):
    """The reference solution-vector layout: x[n*cab+cust] in {0,1} (procedure.py:56, split.py:23)."""
    n = len(col_of_row)
    x = np.zeros(n * n, dtype=np.uint8)
    x[np.arange(n) * n + np.asarray(col_of_row, dtype=np.int64)] = 1
    return x


def lp_relaxation(cost):
    """LP optimum of the reference model with x >= 0 instead of binary (glpk.mod:9)."""
    from scipy.optimize import linprog
    from scipy.sparse import coo_matrix
    c = np.asarray(cost, dtype=np.float64)
    n = c.shape[0]
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    i = i.ravel()
    j = j.ravel()
    rows = np.concatenate([i, n + i])                   # solver.py:18-19
    cols = np.concatenate([n * i + j, n * j + i])
    a = coo_matrix((np.ones(2 * n * n), (rows, cols)), shape=(2 * n, n * n)).tocsr()
    res = linprog(c.ravel(), A_eq=a, b_eq=np.ones(2 * n), bounds=(0, None), method="highs")
    if res.status != 0:
        raise RuntimeError(res.message)
    return float(res.fun), res.x


def check_x(x, n):
    """Row sums = column sums = 1 and entries in {0,1}."""
    m = np.asarray(x).reshape(n, n)
    return bool(np.isin(m, (0, 1)).all() and (m.sum(0) == 1).all() and (m.sum(1) == 1).all())
