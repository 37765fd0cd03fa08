"""
TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's per-cell topology bookkeeping. Never imported by the
product package.

Restates, on plain Python lists, what the reference keeps on its ``Cell`` objects (paths relative to the reference
repository JanisGeise/sparseSpatialSampling v1.0.0):

* ``Cell.nb`` and ``_assign_neighbors`` (sparseSpatialSampling/s_cube.py:904-1186) with ``parent_or_child`` (:1758-1775),
* ``Cell.node_idx`` and ``_assign_indices`` (:1188-1536) with ``check_nb_node`` (:1739-1755),
* ``_check_nb`` (:447-464), the neighbour reset of ``_remove_invalid_cells`` (:721-731),
* ``_resort_nodes_and_indices_of_grid`` (:734-772) with ``renumber_node_indices_parallel`` (:1695-1736).

The reference writes both assignment procedures as hand-unrolled case ladders; here they are two small tables:
the neighbour rule is the geometric one (child offset + direction, leaving the parent through the matching neighbour
slot) and the node rule is the priority table of SURVEY.md appendix C. Pinning: ``tests/golden/make_golden.py``
(i) extracts the reference's own neighbour table by instrumenting ``_assign_neighbors`` and asserts equality with
``neighbour_table`` below for all 4*8 + 8*26 entries, and (ii) asserts that faces and vertices produced by this class
for the golden cases are bit-identical to the reference's.
"""
import numpy as np

CHILD_DIRS = np.array([[-1, -1, 1], [-1, 1, 1], [1, 1, 1], [1, -1, 1], [-1, -1, -1], [-1, 1, -1], [1, 1, -1],
                       [1, -1, -1]], dtype=np.int64)                       # CH order, s_cube.py:29, :188-194
PLANE = [(-1, 0), (-1, 1), (0, 1), (1, 1), (1, 0), (1, -1), (0, -1), (-1, -1)]   # w nw n ne e se s sw, s_cube.py:22-26
SLOT = {"w": 0, "nw": 1, "n": 2, "ne": 3, "e": 4, "se": 5, "s": 6, "sw": 7, "wl": 8, "nwl": 9, "nl": 10, "nel": 11,
        "el": 12, "sel": 13, "sl": 14, "swl": 15, "cl": 16, "wu": 17, "nwu": 18, "nu": 19, "neu": 20, "eu": 21, "seu": 22,
        "su": 23, "swu": 24, "cu": 25}

# node rule (s_cube.py:1215-1536): per child, in evaluation order, (node, [(neighbour slot, node of that neighbour), ...])
# -- the first neighbour that exists, is a leaf and has the child's level lends its node, otherwise a new node is made
NEW_2D = {0: [(1, [("w", 2)]), (2, []), (3, [("s", 2)])], 1: [(2, [("n", 3)])], 2: [(3, [("e", 0)])], 3: []}
COPY_2D = {0: [], 1: [(0, 0, 1), (3, 0, 2)], 2: [(0, 0, 2), (1, 1, 2)], 3: [(0, 0, 3), (1, 0, 2), (2, 2, 3)]}
NEW_3D = {
    0: [(1, [("w", 2), ("wu", 6), ("cu", 5)]), (2, [("cu", 6)]), (3, [("s", 2), ("su", 6), ("cu", 7)]),
        (4, [("w", 7), ("sw", 6), ("s", 5)]), (5, [("w", 6)]), (6, []), (7, [("s", 6)])],
    1: [(2, [("n", 3), ("nu", 7), ("cu", 6)]), (5, [("w", 6), ("nw", 7), ("n", 4)]), (6, [("n", 7)])],
    2: [(3, [("e", 0), ("eu", 4), ("cu", 7)]), (6, [("e", 5), ("ne", 4), ("n", 7)]), (7, [("e", 4)])],
    3: [(7, [("e", 4), ("se", 5), ("s", 6)])],
    4: [(5, [("w", 6), ("wl", 2), ("cl", 1)]), (6, [("cl", 2)]), (7, [("s", 6), ("sl", 2), ("cl", 3)])],
    5: [(6, [("n", 7), ("nl", 3), ("cl", 2)])],
    6: [(7, [("e", 4), ("el", 0), ("cl", 3)])],
    7: [],
}
# (node, sibling, node of that sibling)
COPY_3D = {
    0: [],
    1: [(0, 0, 1), (3, 0, 2), (4, 0, 5), (7, 0, 6)],
    2: [(0, 0, 2), (1, 1, 2), (4, 0, 6), (5, 1, 6)],
    3: [(0, 0, 3), (1, 0, 2), (2, 2, 3), (4, 0, 7), (5, 0, 6), (6, 2, 7)],
    4: [(0, 0, 4), (1, 0, 5), (2, 0, 6), (3, 0, 7)],
    5: [(0, 1, 4), (1, 1, 5), (2, 1, 6), (3, 1, 7), (4, 4, 5), (7, 4, 6)],
    6: [(0, 2, 4), (1, 2, 5), (2, 2, 6), (3, 2, 7), (4, 5, 7), (5, 5, 6)],
    7: [(0, 3, 4), (1, 3, 5), (2, 3, 6), (3, 3, 7), (4, 4, 7), (5, 4, 6), (6, 6, 7)],
}


def slot_directions(d: int) -> list:
    """Direction vector of every neighbour slot (NB order)."""
    if d == 2:
        return [p for p in PLANE]
    return [p + (0,) for p in PLANE] + [p + (-1,) for p in PLANE + [(0, 0)]] + [p + (1,) for p in PLANE + [(0, 0)]]


def neighbour_table(d: int) -> list:
    """
    table[c][s] = ("sibling", j) or ("poc", parent slot, child j): the entry the reference assigns to
    ``children[c].nb[s]`` in ``_assign_neighbors`` -- ``children[j]`` or ``parent_or_child(cell.nb, check[P], P, j)``.
    """
    dirs = slot_directions(d)
    nch = 2 ** d
    table = []
    for c in range(nch):
        row = []
        for s, delta in enumerate(dirs):
            pdir, off = [0] * d, [0] * d
            for a in range(d):
                t = (1 if CHILD_DIRS[c][a] > 0 else 0) + delta[a]
                if t < 0:
                    pdir[a], off[a] = -1, 1
                elif t > 1:
                    pdir[a], off[a] = 1, 0
                else:
                    off[a] = t
            j = [k for k in range(nch) if all((CHILD_DIRS[k][a] > 0) == (off[a] == 1) for a in range(d))][0]
            if not any(pdir):
                row.append(("sibling", j))
            else:
                row.append(("poc", dirs.index(tuple(pdir)), j))
        table.append(row)
    return table


class OracleTopology:
    LEAF, EMPTY = -1, -2          # children is None / children == []

    def __init__(self, d: int, root_center, width: float):
        self.d, self.nch, self.nnb = d, 2 ** d, 8 if d == 2 else 26
        self.width = float(width)
        self.table = neighbour_table(d)
        self.new_rules = NEW_2D if d == 2 else NEW_3D
        self.copy_rules = COPY_2D if d == 2 else COPY_3D
        self.dirs = CHILD_DIRS[:self.nch, :d].astype(np.float64)
        c = np.asarray(root_center, dtype=np.float64)
        # _create_first_cell, s_cube.py:338-397
        self.parent, self.children, self.level = [-1], [self.LEAF], [0]
        self.nb = [[-1] * self.nnb]
        self.node = [list(range(self.nch))]
        self.center = [c.copy()]
        self.nodes = [c + self.dirs[j] * 0.5 * self.width for j in range(self.nch)]

    # ---- _assign_neighbors(cell, children=cell.children)
    def assign_neighbors(self, p: int) -> None:
        first = self.children[p]
        if first < 0:
            return
        pnb = self.nb[p]
        for c in range(self.nch):
            cnb = self.nb[first + c]
            for s, e in enumerate(self.table[c]):
                if e[0] == "sibling":
                    cnb[s] = first + e[1]
                else:
                    q = pnb[e[1]]
                    if q < 0:
                        cnb[s] = -1
                    elif self.children[q] >= 0:          # bool(n and n.children)
                        cnb[s] = self.children[q] + e[2]
                    else:
                        cnb[s] = q

    def _shares(self, cell: int, slot: int) -> bool:      # check_nb_node, s_cube.py:1739-1755
        q = self.nb[cell][slot]
        return q >= 0 and self.children[q] == self.LEAF and self.level[q] == self.level[cell]

    def assign_indices(self, p: int) -> None:
        first = self.children[p]
        for c in range(self.nch):
            cell = first + c
            nd = self.node[cell]
            nd[c] = self.node[p][c]
            for node, sources in self.new_rules[c]:
                for name, j in sources:
                    if self._shares(cell, SLOT[name]):
                        nd[node] = self.node[self.nb[cell][SLOT[name]]][j]
                        break
                else:
                    # _compute_cell_centers(_factor=0.5, _cell=cell), s_cube.py:441
                    self.nodes.append(self.center[cell] + self.dirs[node] * 0.5 * self.width / (2 ** self.level[cell]))
                    nd[node] = len(self.nodes) - 1
            for node, sib, j in self.copy_rules[c]:
                nd[node] = self.node[first + sib][j]

    # ---- interface used by OracleTree
    def refine(self, parents) -> None:
        for p in parents:
            first = len(self.parent)
            for c in range(self.nch):
                self.parent.append(p)
                self.children.append(self.LEAF)
                self.level.append(self.level[p] + 1)
                self.nb.append([-1] * self.nnb)
                self.node.append([0] * self.nch)
                self.center.append(self.center[p] + self.dirs[c] * 0.25 * self.width / (2 ** self.level[p]))
            self.children[p] = first
            self.assign_neighbors(p)
            self.assign_indices(p)

    def refresh_siblings(self, cells) -> None:
        for c in cells:
            if self.parent[c] >= 0:
                self.assign_neighbors(self.parent[c])

    def refresh_children(self, parents) -> None:
        for p in parents:
            self.assign_neighbors(p)

    def mark_invalid(self, cells) -> None:                # s_cube.py:721-731
        for cell in cells:
            self.children[cell] = self.EMPTY
            for q in self.nb[cell]:
                if q >= 0:
                    self.nb[q] = [-1 if x == cell else x for x in self.nb[q]]

    def check_nb(self, cell: int) -> list:                # s_cube.py:447-464
        return [q for q in self.nb[cell] if q >= 0 and self.children[q] == self.LEAF and self.level[q] < self.level[cell]]

    def final(self):                                      # s_cube.py:734-772, 1695-1736
        leaves = [c for c in range(len(self.parent)) if self.children[c] == self.LEAF]
        all_idx = np.array([self.node[c] for c in leaves], dtype=np.int64)
        used = set(all_idx.reshape(-1).tolist())
        available = set(range(self.nch)) | set(range(int(all_idx.min()), int(all_idx.max()) + 1))
        unused = available - used
        mapping, counter = {}, 0
        for i in range(len(self.nodes)):
            if i not in unused:
                mapping[i] = counter
                counter += 1
        vertices = np.stack([self.nodes[i] for i in range(len(self.nodes)) if i in mapping])
        faces = np.vectorize(mapping.get)(all_idx).astype(np.int32)
        return faces, vertices, np.stack(self.center)
