"""Exact permanent of a sparse integer matrix in Python integers, independent of the library: rows are taken
one at a time; a state is the set of already used columns among the "open" columns (columns that occur both in
a processed and in an unprocessed row); its value is the exact weighted number of ways to match the processed
rows into exactly those columns plus all closed ones.  Columns whose last row has been processed must be used
(otherwise the state dies).  The row order is chosen greedily to keep the set of open columns small.
Cost: sum over rows of (#states x row degree); fine while the open set stays below ~25 columns.

usage: exact_permanent_dp.py <name in tests/golden/known_perman.json> [max_states]"""
import json, os, sys, time


def exact_permanent(A, max_states=50_000_000, verbose=False):
    n = len(A)
    rows = [[(j, A[i][j]) for j in range(n) if A[i][j] != 0] for i in range(n)]
    col_rows = [set(i for i in range(n) if A[i][j] != 0) for j in range(n)]
    if any(not r for r in rows) or any(not c for c in col_rows):
        return 0
    # greedy order: next row = the one that opens the fewest new columns (ties: closes the most)
    done, order, open_cols = set(), [], set()
    while len(order) < n:
        best = None
        for i in range(n):
            if i in done:
                continue
            new = sum(1 for j, _ in rows[i] if j not in open_cols)
            closes = sum(1 for j, _ in rows[i] if col_rows[j] <= done | {i})
            key = (new - closes, new, i)
            if best is None or key < best[0]:
                best = (key, i)
        i = best[1]
        order.append(i); done.add(i)
        open_cols |= {j for j, _ in rows[i]}
        open_cols = {j for j in open_cols if not col_rows[j] <= done}
    states = {0: 1}                       # bitmask over ALL columns (only open ones can be set)
    done = set()
    peak = 1
    for step, i in enumerate(order):
        done.add(i)
        closing = 0
        for j, _ in rows[i]:
            if col_rows[j] <= done:
                closing |= 1 << j
        nxt = {}
        for used, cnt in states.items():
            for j, a in rows[i]:
                b = 1 << j
                if used & b:
                    continue
                u2 = used | b
                if (u2 & closing) != closing:
                    continue              # a column whose last row this is stays unused: dead
                u2 &= ~closing            # closed columns are used by construction: drop them from the key
                nxt[u2] = nxt.get(u2, 0) + cnt * a
        states = nxt
        peak = max(peak, len(states))
        if verbose:
            print("row %2d (#%d): %d states" % (i, step, len(states)), flush=True)
        if len(states) > max_states:
            raise MemoryError("more than %d states at step %d" % (max_states, step))
        if not states:
            return 0
    assert list(states.keys()) == [0], states.keys()
    return states[0]


def load(name):
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = json.load(open(os.path.join(here, "tests", "golden", "known_perman.json")))[name]
    n = d["n"]
    A = [[0] * n for _ in range(n)]
    for i, j, v in d["triples"]:
        A[i][j] = int(v)
    return A


if __name__ == "__main__":
    A = load(sys.argv[1])
    t = time.time()
    p = exact_permanent(A, int(sys.argv[2]) if len(sys.argv) > 2 else 50_000_000, verbose=True)
    print(sys.argv[1], "exact permanent =", p, "(%.1f s)" % (time.time() - t))
