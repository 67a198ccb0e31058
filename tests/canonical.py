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


def lod_pyramid(dense_yzx: np.ndarray):
    """Independent restatement of what to_vec returns at every level of detail (pure numpy).

    A branch's value is calc_average of its eight children's values (core/voxel.rs:96-141): the most
    frequent value; on a tie a non-default value wins, then the earliest first occurrence in child order
    i = x | y<<1 | z<<2.  EMPTY children count as the default (0); a uniform-collapsed Leaf keeps its value,
    which is what the same rule gives.  Returns [lod 0 (input), lod 1, ...] down to one voxel."""
    out = [dense_yzx]
    cur = dense_yzx.astype(np.int64)
    while cur.shape[0] > 1:
        m = cur.shape[0] // 2
        # children of cell (Y, Z, X) in child order i = x | y<<1 | z<<2
        kids = np.stack([cur[(i >> 1 & 1)::2, (i >> 2 & 1)::2, (i & 1)::2] for i in range(8)], axis=-1)  # [Y][Z][X][8]
        flat = kids.reshape(-1, 8)
        best = flat[:, 0].copy()
        best_score = np.full(len(flat), -1, np.int64)
        for i in range(8):
            v = flat[:, i]
            cnt = (flat == v[:, None]).sum(1)
            first = np.ones(len(flat), bool)
            for j in range(i):
                first &= flat[:, j] != v
            score = np.where(first, cnt * 2 + (v != 0), -1)   # count, then non-default; earlier i wins ties
            take = score > best_score
            best = np.where(take, v, best)
            best_score = np.where(take, score, best_score)
        cur = best.reshape(m, m, m)
        out.append(cur.astype(dense_yzx.dtype))
    return out
