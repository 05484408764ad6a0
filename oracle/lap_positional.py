"""TEST INFRASTRUCTURE — not imported by the product package.

Numpy / pure-Python model of the data structure `lap_kernel_v3` (pleas_merging_b200/csrc/lap.cu) uses in place of
SciPy's `remaining` list: every column carries its list POSITION, the list itself is never materialised.  The
algorithm it models is SciPy's `linear_sum_assignment` (rectangular_lsap.cpp, Crouse's shortest augmenting path),
which the reference calls through `scipy_solve_lsa` (/root/reference/pleas/core/solvers.py:18-33) on the float32
cost matrix widened to float64 and negated for `maximize`.

What the model pins (tests/test_oracle.py checks it against the SciPy goldens, ties included):
  * `remaining` is filled in reverse, so column j starts at position n-1-j;
  * removing position t moves the column at the LAST position into t (its owner sees `pos == last`);
  * among columns at the minimum tentative distance an unassigned one wins — the last such position —,
    otherwise the first position: rank = n-1-pos (unassigned) / n+pos (assigned), smallest (distance, rank) wins;
  * the candidate word rank << 13 | (row4col + 1) orders like the rank and carries the next row;
  * the sink is the only scanned unassigned column; duals are updated by the owners of the scanned columns.
Small sizes only (O(n^3) Python loops).
"""
import numpy as np

INF = float("inf")


def solve_lsa_positional(cost, maximize=False):
    """Returns col4row (int64) computed with lap_kernel_v3's bookkeeping."""
    C = np.asarray(cost, dtype=np.float32)
    n = C.shape[0]
    assert C.shape == (n, n) and n <= 4096
    Cs = (-C if maximize else C).astype(np.float64)
    u, v = np.zeros(n), np.zeros(n)
    pred = -np.ones(n, dtype=np.int64)
    row4col = -np.ones(n, dtype=np.int64)
    col4row = -np.ones(n, dtype=np.int64)
    for cur in range(n):
        r4c = row4col.copy()  # the owners' register copies, reloaded after every flip
        kb = np.where(r4c < 0, (n - 1) << 13, (n << 13) | (r4c + 1))
        ks = np.where(r4c < 0, -8192, 8192)
        dist = np.full(n, INF)
        pos = n - 1 - np.arange(n)
        row, ntodo, min_val = cur, n, 0.0
        while True:
            best = (INF, 0xFFFFFFFF)
            for j in range(n):
                if pos[j] < 0:
                    continue
                r = ((min_val + Cs[row, j]) - u[row]) - v[j]
                if r < dist[j]:
                    dist[j], pred[j] = r, row
                pk = int(kb[j] + ks[j] * pos[j])
                if (dist[j], pk) < best:
                    best = (dist[j], pk)
            if best[0] == INF:
                raise ValueError("cost matrix is infeasible")
            min_val, word = best
            grank, next_row = word >> 13, (word & 0x1FFF) - 1
            t = (n - 1 - grank) if grank < n else grank - n
            last = ntodo - 1
            pos = np.where(pos == t, -1, np.where(pos == last, t, pos))
            ntodo = last
            if next_row < 0:
                break
            row = next_row
        scanned = np.nonzero(pos == -1)[0]
        sink = [j for j in scanned if r4c[j] < 0]
        assert len(sink) == 1
        for j in scanned:
            dv = min_val - dist[j]
            v[j] = v[j] - dv
            if r4c[j] >= 0:
                u[r4c[j]] = u[r4c[j]] + dv
        u[cur] = u[cur] + min_val
        j = sink[0]
        while True:
            i = pred[j]
            row4col[j] = i
            col4row[i], j = j, col4row[i]
            if i == cur:
                break
    return col4row
