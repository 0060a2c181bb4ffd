"""CPU oracle for the adaptive grid refinement hand-off -- TEST INFRASTRUCTURE ONLY (never imported by the product).

numpy / pure-Python restatement of the reference's `new_grid` dispatch (src/numerical/BVP_Damp/grid_api.rs:154-189) and of
the algorithms in src/numerical/BVP_Damp/adaptive_grid_basic.rs it dispatches to.  Loops follow the reference statement by
statement (including its index quirks) so marks, meshes and interpolated guesses are reproduced bit for bit; sizes are
test sizes.  Pinned by the reference's own unit tests (adaptive_grid_basic.rs:992-1241, restated in
tests/test_adaptive_grid.py): fixtures `boundary_layer_fixture` / `flat_fixture`, the refinement invariants, the uniform
mid-point test and the SciPy threshold known-answer test.

Conventions: `y` is the reference's y_DMatrix, shape (n_vars, n_points); `x` the mesh, shape (n_points,).
Every function returns (new_grid, new_initial_guess, n_added, marks).
"""
from __future__ import annotations

import numpy as np

DOUBLE_POINTS, EASY, PEARSON, GRCAR_SMOOKE, SCI = range(5)


def _as_i32(v: float) -> int:
    """Rust `f64 as i32`: truncation toward zero, saturating, NaN -> 0."""
    if v != v:
        return 0
    if v >= 2147483647.0:
        return 2147483647
    if v <= -2147483648.0:
        return -2147483648
    return int(v)


def construct_refined_grid_from_marks(y, x, mark):
    """adaptive_grid_basic.rs:56-105."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    assert len(mark) == len(x)
    assert all(m >= 0 for m in mark)
    total = int(sum(int(m) for m in mark))
    n_rows = y.shape[0]
    new_grid = []
    cols = []
    for i in range(len(x)):
        yi = y[:, i]
        new_grid.append(x[i])
        cols.append(yi.copy())
        n = int(mark[i])
        if n > 0 and i < len(x) - 1:
            h_i = x[i + 1] - x[i]
            dy = y[:, i + 1] - yi
            for k in range(1, n + 1):
                theta = np.float64(k) / (np.float64(n) + 1.0)
                new_grid.append(x[i] + h_i * theta)
                cols.append(yi + dy * theta)
    guess = np.stack(cols, axis=1) if cols else np.zeros((n_rows, 0))
    return np.array(new_grid), guess, total, np.array(mark, dtype=np.int32)


def refine_all_grid(y, x):
    """adaptive_grid_basic.rs:187-203 (DoublePoints)."""
    mark = [1] * len(x)
    if mark:
        mark[-1] = 0
    return construct_refined_grid_from_marks(y, x, mark)


def easy_grid_refinement(y, x, tolerance):
    """adaptive_grid_basic.rs:257-313 (parallel variant; marks are 0/1 so the merge order is irrelevant)."""
    y = np.asarray(y, dtype=np.float64)
    mark = [0] * len(x)
    for j in range(y.shape[0]):
        row = y[j]
        delta = tolerance * (row.max() - row.min())
        for i in range(1, len(x) - 1):
            if abs(row[i] - row[i - 1]) > delta:
                mark[i - 1] = 1
    return construct_refined_grid_from_marks(y, x, mark)


def _merge_legacy_shifted_mark(mark, trigger_idx, value):
    """adaptive_grid_basic.rs:50-54."""
    if 0 < trigger_idx < len(mark) and value > mark[trigger_idx]:
        mark[trigger_idx - 1] = value


def _buffer_pass(mark, x, C):
    """adaptive_grid_basic.rs:517-533 (same statements at :703-733)."""
    h = [x[i + 1] - x[i] for i in range(len(x) - 1)]
    for i in range(1, len(x) - 1):
        ratio = h[i] / h[i - 1]
        c1 = ratio <= C
        c2 = ratio >= 1.0 / C
        if (not c1) and (i - 1) != 0:
            if 1 > mark[i - 1]:
                mark[i - 1] = 1
        if not c2:
            if 1 > mark[i]:
                mark[i - 1] = 1


def pearson_grid_refinement(y, x, d, C):
    """adaptive_grid_basic.rs:469-540 (the parallel variant `new_grid` calls): per-row triggers are collected first and
    merged row by row, in increasing i."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    per_row = []
    for j in range(y.shape[0]):
        row = y[j]
        delta = d * (row.max() - row.min())
        local = []
        for i in range(n):
            if n - 1 > i and i > 0:
                tau = abs(row[i] - row[i - 1])
                not_small = (abs(row[i]) > 1e-4) and (abs(row[i + 1]) > 1e-4)
                if tau > delta and not_small:
                    with np.errstate(divide="ignore", invalid="ignore"):
                        q = _as_i32(float(np.float64(tau) / np.float64(delta)))
                    local.append((i, q if q >= 1 else 1))
        per_row.append(local)
    mark = [0] * n
    for local in per_row:
        for trigger_idx, N in local:
            _merge_legacy_shifted_mark(mark, trigger_idx, N)
    _buffer_pass(mark, x, C)
    return construct_refined_grid_from_marks(y, x, mark)


def grcar_smooke_grid_refinement(y, x, d, g, C):
    """adaptive_grid_basic.rs:565-741 (the sequential variant `new_grid` calls)."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    h = [x[i + 1] - x[i] for i in range(n - 1)]
    mark = [0] * n
    for j in range(y.shape[0]):
        row = y[j]
        delta = d * (row.max() - row.min())
        slopes = [(row[i + 1] - row[i]) / h[i] for i in range(n - 1)]
        gamma = g * (max(slopes) - min(slopes))
        for i in range(n):
            if n - 1 > i and i > 0:
                s1 = (row[i + 1] - row[i]) / h[i]
                s0 = (row[i] - row[i - 1]) / h[i - 1]
                eta = abs(s1 - s0)
                tau = abs(row[i] - row[i - 1])
                not_small = (abs(row[i]) > 1e-4) and (abs(row[i + 1]) > 1e-4)
                if (tau > delta and not_small) or (eta > gamma and not_small):
                    with np.errstate(divide="ignore", invalid="ignore"):
                        q = _as_i32(float(np.float64(tau) / np.float64(delta)))
                    N = q if q >= 1 else 1
                    if N > mark[i]:
                        mark[i - 1] = N
    _buffer_pass(mark, x, C)
    return construct_refined_grid_from_marks(y, x, mark)


def scipy_grid_refinement(y, x, tolerance, residuals):
    """adaptive_grid_basic.rs:885-924 with modify_mesh (:926-945) and interpolate_solution (:947-989).  The residual index is
    used as the interval index, as in the reference; an index past the mesh raises like the reference's slice panic."""
    y = np.asarray(y, dtype=np.float64)
    x = np.asarray(x, dtype=np.float64)
    res = np.asarray(residuals, dtype=np.float64)
    insert_1, insert_2 = [], []
    for j in range(len(res) - 1):
        if res[j] > tolerance and res[j] < 100.0 * tolerance:
            insert_1.append(j)
        elif res[j] >= 100.0 * tolerance:
            insert_2.append(j)
    added = len(insert_1) + 2 * len(insert_2)
    mark = [0] * len(x)
    if added == 0:
        return x.copy(), y.copy(), 0, np.array(mark, dtype=np.int32)
    pts = list(x)
    for i in insert_1:
        if i + 1 >= len(x):
            raise IndexError("modify_mesh: index out of bounds (reference panics)")
        pts.append(0.5 * (x[i] + x[i + 1]))
        mark[i] = 1
    for i in insert_2:
        if i + 1 >= len(x):
            raise IndexError("modify_mesh: index out of bounds (reference panics)")
        pts.append((2.0 * x[i] + x[i + 1]) / 3.0)
        pts.append((x[i] + 2.0 * x[i + 1]) / 3.0)
        mark[i] = 2
    new_x = np.array(sorted(pts))
    cols = []
    for i in range(y.shape[1]):
        cols.append(y[:, i].copy())
        N = mark[i]
        if N and i < y.shape[1] - 1:
            dy = y[:, i + 1] - y[:, i]
            for k in range(1, N + 1):
                t = np.float64(k) / (np.float64(N) + 1.0)
                cols.append(y[:, i] + dy * t)
    return new_x, np.stack(cols, axis=1), added, np.array(mark, dtype=np.int32)


def new_grid(method, y, x, abs_tolerance=1e-6, residuals=None, params=()):
    """grid_api.rs:154-189."""
    if method == DOUBLE_POINTS:
        return refine_all_grid(y, x)
    if method == EASY:
        return easy_grid_refinement(y, x, params[0])
    if method == PEARSON:
        return pearson_grid_refinement(y, x, params[0], params[1])
    if method == GRCAR_SMOOKE:
        return grcar_smooke_grid_refinement(y, x, params[0], params[1], params[2])
    if method == SCI:
        if residuals is None:
            raise ValueError("residual vector is required for scipy grid refinement")
        return scipy_grid_refinement(y, x, abs_tolerance, residuals)
    raise ValueError("TwoPoint refinement is outside the hot-path scope")
