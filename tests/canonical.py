"""Independent canonical SVO-DAG builder (pure numpy/python) — SURVEY.md §7.0.

Derived from the *definition* of the fresh-tree result, not from the reference's
control flow, so it cross-checks the oracle restatement (oracle/voxelis_oracle.hpp):

    node(voxel v)   = EMPTY if v == 0 else Leaf(v)
    node(cube)      = EMPTY            if all eight children EMPTY
                    = that Leaf        if all eight children are the same Leaf
                    = Branch(c0..c7)   otherwise, interned on the children tuple

Node encoding here: 0 = EMPTY, -v = Leaf(v), k > 0 = k-th distinct branch.
"""
import numpy as np

from voxelis_b200.workloads import lane_coords


class CanonicalDag:
    def __init__(self):
        self.branches = {}   # tuple(children) -> id
        self.children = [None]  # id -> tuple

    def intern(self, key):
        i = self.branches.get(key)
        if i is None:
            i = len(self.children)
            self.branches[key] = i
            self.children.append(key)
        return i

    def build(self, dense_yzx: np.ndarray) -> int:
        """dense in to_vec layout [y][z][x]; returns the root node."""
        n = dense_yzx.shape[0]
        depth = int(np.log2(n))
        x, y, z = lane_coords(depth)
        level = -dense_yzx[y, z, x].astype(np.int64)  # Morton order, leaf encoding
        for _ in range(depth):
            rows = level.reshape(-1, 8)
            uniq, inv = np.unique(rows, axis=0, return_inverse=True)
            ids = np.empty(len(uniq), np.int64)
            for k, r in enumerate(uniq):
                if (r == r[0]).all() and r[0] <= 0:
                    ids[k] = r[0]            # all EMPTY -> EMPTY ; same leaf -> that leaf
                else:
                    ids[k] = self.intern(tuple(int(v) for v in r))
            level = ids[inv.reshape(-1)]
        return int(level[0])

    def per_depth(self, roots, depth):
        """[(#distinct branches, #distinct leaves)] reachable at each depth 0..depth."""
        out = []
        level = {r for r in roots if r != 0}
        for _ in range(depth + 1):
            b = [i for i in level if i > 0]
            out.append((len(b), len(level) - len(b)))
            nxt = set()
            for i in b:
                nxt.update(c for c in self.children[i] if c != 0)
            level = nxt
        return out

    def reachable(self, roots):
        seen_b, seen_l = set(), set()
        stack = [r for r in roots if r != 0]
        while stack:
            i = stack.pop()
            if i < 0:
                seen_l.add(i)
            elif i not in seen_b:
                seen_b.add(i)
                stack.extend(c for c in self.children[i] if c != 0)
        return len(seen_b), len(seen_l)

    def stream(self, roots):
        """Same record stream as oracle_capi.cpp:orc_dag_signature (post-order numbering)."""
        number, words = {}, []

        def visit(i):
            if i == 0 or i in number:
                return
            if i < 0:
                number[i] = len(number) + 1
                words.extend([1, -i, 0, 0, 0, 0, 0, 0, 0])
                return
            for c in self.children[i]:
                visit(c)
            number[i] = len(number) + 1
            words.append(0)
            words.extend(0 if c == 0 else number[c] for c in self.children[i])

        for r in roots:
            visit(r)
        words.extend(0 if r == 0 else number[r] for r in roots)
        return np.array(words, np.uint64)

    def indegree(self, roots):
        """in-degree from distinct reachable branches + root handles, keyed by node."""
        deg = {}
        seen = set()
        stack = []
        for r in roots:
            if r != 0:
                deg[r] = deg.get(r, 0) + 1
                stack.append(r)
        while stack:
            i = stack.pop()
            if i <= 0 or i in seen:
                continue
            seen.add(i)
            for c in self.children[i]:
                if c != 0:
                    deg[c] = deg.get(c, 0) + 1
                    stack.append(c)
        return deg
