"""ORACLE (test infrastructure, never imported by the product): pure-Python restatement of the
reference's algebraic partitioner for small graphs.

Follows ``src/partitioners/mod.rs`` (``new_ls_strength_graph`` :337-393, ``extract_local_subgraph``
:695-718, ``pairwise_merge`` / ``map_indices`` / ``merge_pair`` :439-463, :505-586,
``Partition::pairwise_merge`` :110-129) and ``src/partitioners/modularity.rs`` (``Partitioner::new``
:28-137, ``initialize_partition`` :179-192, ``generate_modularity_triplets`` :305-337,
``greedy_matching`` :339-383, ``size_cost`` / ``delta_q`` :385-435, ``improve_partition`` :437-510).

**Parity unpinned**: the reference ships no tests for the partitioner and is itself not
deterministic (unstable float sorts with ties, ``max_by`` over a randomly seeded
``HashSet``).  Every unspecified order is fixed by the TIE-BREAK rules 1-5 listed in
``faer_amg_b200/csrc/partition.cu``; this file applies the same rules with different data
structures, so the C++ product and this restatement must agree bit for bit.  F10c
(``pairwise_merge_rowsums`` missing ``+ pairs.len()``) is fixed in both, as SURVEY 7 prescribes.

Python's ``sum()`` uses compensated summation for floats since 3.12 -- every reduction below is an
explicit left-to-right loop, like the Rust iterators it restates.
"""
from __future__ import annotations

import math
from collections import deque
from typing import List, Sequence, Tuple

import numpy as np

Edge = Tuple[int, float]


def extract_local_subgraph(row_ptr, col_idx, center: int, max_depth: int) -> List[int]:
    """mod.rs:695-718 -- the visited set in ascending order (BTreeSet iteration)."""
    visited = {center}
    todo = deque([(center, 0)])
    while todo:
        j, depth = todo.popleft()
        if depth < max_depth:
            for q in range(int(row_ptr[j]), int(row_ptr[j + 1])):
                nb = int(col_idx[q])
                if nb not in visited:
                    visited.add(nb)
                    todo.append((nb, depth + 1))
    return sorted(visited)


def _wdot(vi, w, vj) -> float:
    s = 0.0
    for c in range(len(w)):
        s += float(vi[c]) * float(w[c]) * float(vj[c])
    return s


def new_ls_strength_graph(row_ptr, col_idx, near_null, weights: Sequence[float], max_depth: int = 3) -> List[List[Edge]]:
    """mod.rs:337-393."""
    near_null = np.asarray(near_null, dtype=np.float64)
    if near_null.ndim == 1:
        near_null = near_null.reshape(-1, 1)
    n = len(row_ptr) - 1
    w = [float(x) for x in list(weights)[: near_null.shape[1]]]
    theta, eps = 0.5, 1e-30
    nodes: List[List[Edge]] = [[] for _ in range(n)]
    for i in range(n):
        local = extract_local_subgraph(row_ptr, col_idx, i, max_depth)
        vi = near_null[i]
        vi_norm = max(_wdot(vi, w, vi), eps)
        for j in local:
            if j <= i:
                continue
            vj = near_null[j]
            vj_norm = max(_wdot(vj, w, vj), eps)
            x = _wdot(vi, w, vj)
            rho2 = (x * x) / (vi_norm * vj_norm)
            d = 2.0 * math.sqrt(max(1.0 - rho2, 0.0))
            nodes[i].append((j, d))
            nodes[j].append((i, d))
    eps, alpha = 1e-12, 4.0
    out: List[List[Edge]] = []
    for nb in nodes:
        if not nb:
            raise ValueError("graph is disconnected")
        nb = sorted(nb, key=lambda e: e[1])  # TIE-BREAK 1 (list.sort is stable)
        keep = max(int(math.floor(len(nb) * theta)), 1)
        nb = nb[:keep]
        d_min, d_max = nb[0][1], nb[-1][1]
        if abs(d_max - d_min) < eps:
            nb = [(j, 1.0) for j, _ in nb]
        else:
            nb = [(j, math.pow((d_max - d) / (d_max - d_min + eps), alpha)) for j, d in nb]
        out.append(sorted(nb, key=lambda e: e[0]))
    return out


def _merge_pair(a: List[Edge], b: List[Edge]) -> List[Edge]:
    """mod.rs:518-586."""
    merged: List[List] = []

    def add(idx, wt):
        if merged and merged[-1][0] == idx:
            merged[-1][1] += wt
        else:
            merged.append([idx, wt])

    ia = ib = 0
    while ia < len(a) and ib < len(b):
        if a[ia][0] == b[ib][0]:
            add(a[ia][0], a[ia][1] + b[ib][1])
            ia += 1
            ib += 1
        elif a[ia][0] < b[ib][0]:
            add(*a[ia])
            ia += 1
        else:
            add(*b[ib])
            ib += 1
    for e in a[ia:]:
        add(*e)
    for e in b[ib:]:
        add(*e)
    return [(i, wt) for i, wt in merged]


class Partitioner:
    """modularity.rs:15-137 with ``starting_partition = None``, unit node weights."""

    def __init__(self, strength: List[List[Edge]], coarsening_factor: float = 8.0, agg_size_penalty: float = 1.0,
                 max_improvement_iters: int = 100):
        self.cf, self.pen, self.max_iters = float(coarsening_factor), float(agg_size_penalty), max_improvement_iters
        self.base = [list(nb) for nb in strength]
        self.strength = [list(nb) for nb in strength]
        n = len(strength)
        self.row_sums = []
        for i, nb in enumerate(strength):
            s = 0.0
            for j, wt in nb:
                assert j != i
                s += wt
            self.row_sums.append(0.0 if s < 0.0 else s)
        total = 0.0
        for s in self.row_sums:
            total += s
        self.inverse_total = 1.0 / total
        self.node_to_agg = list(range(n))
        self.agg_to_node = [[i] for i in range(n)]
        self.agg_sizes = [1] * n

    def part_cf(self) -> float:
        return len(self.node_to_agg) / len(self.agg_to_node)

    def modularity_triplets(self):
        out = []
        for i, nb in enumerate(self.strength):
            for j, s in nb:
                if not i > j:
                    continue
                expected = self.inverse_total * self.row_sums[i] * self.row_sums[j]
                wt = s - expected
                new_weight = float(self.agg_sizes[i] + self.agg_sizes[j])
                sq = math.pow(new_weight - self.cf, 2.0)
                if new_weight > self.cf:
                    wt -= self.pen * sq
                else:
                    wt += self.pen * sq
                out.append((i, j, wt))
        return out

    def greedy_matching(self, step_cf: float):
        vertex_count = len(self.row_sums)
        t = math.ceil(vertex_count - len(self.node_to_agg) / step_cf)
        target = max(int(t), 0) + 1
        wants = self.modularity_triplets()
        pairs, unmatched = [], []
        if not wants:
            return pairs, unmatched
        wants.sort(key=lambda tr: tr[2])  # TIE-BREAK 3
        alive = [True] * vertex_count
        while wants:
            i, j, _ = wants.pop()
            if alive[i] and alive[j]:
                alive[i] = alive[j] = False
                pairs.append((i, j))
            if len(pairs) > target:
                break
        unmatched = [i for i, a in enumerate(alive) if a]
        return pairs, unmatched

    def pairwise_merge(self, pairs, unmatched):
        np_ = len(pairs)
        ids = {}
        for a, (i, j) in enumerate(pairs):
            ids[i] = a
            ids[j] = a
        for a, i in enumerate(unmatched):
            ids[i] = a + np_
        mapped = [sorted(((ids[j], wt) for j, wt in nb), key=lambda e: e[0]) for nb in self.strength]  # TIE-BREAK 2
        self.strength = [_merge_pair(mapped[i], mapped[j]) for i, j in pairs] + [mapped[i] for i in unmatched]
        self.agg_to_node = [sorted(self.agg_to_node[i] + self.agg_to_node[j]) for i, j in pairs] + \
                           [self.agg_to_node[i] for i in unmatched]
        # F10c fixed: unmatched row sums land behind the pairs
        self.row_sums = [self.row_sums[i] + self.row_sums[j] for i, j in pairs] + [self.row_sums[i] for i in unmatched]
        for a, agg in enumerate(self.agg_to_node):
            for node in agg:
                self.node_to_agg[node] = a
        self.agg_sizes = [len(agg) for agg in self.agg_to_node]

    def initialize_partition(self):
        while self.part_cf() < self.cf:
            pairs, unmatched = self.greedy_matching(self.cf)
            if not pairs:
                break
            self.pairwise_merge(pairs, unmatched)

    def size_cost(self, size: int) -> float:
        rel = abs(float(size) - self.cf) / self.cf
        return math.pow(4.0 * rel, 4.0) * self.pen

    def delta_q(self, node: int, src: int, dst: int) -> float:
        ind = outd = 0.0
        for j, s in self.base[node]:
            a = self.node_to_agg[j]
            if a == src:
                ind += s
            elif a == dst:
                outd += s
        old_cost = self.size_cost(self.agg_sizes[dst]) + self.size_cost(self.agg_sizes[src])
        new_cost = self.size_cost(self.agg_sizes[dst] + 1) + self.size_cost(self.agg_sizes[src] - 1)
        return (outd - ind) + self.pen * (old_cost - new_cost)

    def improve_partition(self):
        n = len(self.node_to_agg)
        for _ in range(self.max_iters):
            swaps = []
            for i in range(n):
                agg_i = self.node_to_agg[i]
                if self.agg_sizes[agg_i] == 1:
                    continue
                best = None
                for agg_j in sorted({self.node_to_agg[j] for j, _ in self.base[i]} - {agg_i}):  # TIE-BREAK 4
                    dq = self.delta_q(i, agg_i, agg_j)
                    if dq > 0.0 and (best is None or dq >= best[1]):
                        best = (agg_j, dq)
                if best is not None:
                    swaps.append((i, best[0], best[1]))
            if not swaps:
                break
            swaps.sort(key=lambda s: -s[2])  # TIE-BREAK 5
            alive_nodes = [True] * n
            alive_aggs = [True] * len(self.agg_to_node)
            for node, new_agg, _ in swaps:
                old_agg = self.node_to_agg[node]
                if alive_nodes[node] and alive_aggs[new_agg] and alive_aggs[old_agg]:
                    self.node_to_agg[node] = new_agg
                    self.agg_sizes[old_agg] -= 1
                    self.agg_sizes[new_agg] += 1
                    self.agg_to_node[old_agg].remove(node)
                    self.agg_to_node[new_agg] = sorted(self.agg_to_node[new_agg] + [node])
                    alive_aggs[new_agg] = alive_aggs[old_agg] = False
                    alive_nodes[node] = False
                    for j, _ in self.base[node]:
                        alive_nodes[j] = False
                        alive_aggs[self.node_to_agg[j]] = False


def block_reduce(nodes: List[List[Edge]], block_size: int) -> List[List[Edge]]:
    """``strength.aggregate(&block_reduce); strength.filter_diag()`` (mod.rs:293-300, 465-502, 588-656):
    neighbour ids -> node ids, the lists of a node's dofs merged (equal ids summed in member / list order,
    TIE-BREAK 6), weights divided by the largest merged weight (self loops included), self loops dropped."""
    nb = len(nodes) // block_size
    merged, gmax = [], None
    for b in range(nb):
        allv = []
        for o in range(block_size):
            nbrs = nodes[b * block_size + o]
            if not nbrs:
                raise ValueError("empty neighborhood means graph is disconnected")
            allv.extend((j // block_size, wt) for j, wt in nbrs)
        allv.sort(key=lambda e: e[0])
        out: List[List] = []
        for j, wt in allv:
            if out and out[-1][0] == j:
                out[-1][1] += wt
            else:
                out.append([j, wt])
        for _, wt in out:
            gmax = wt if gmax is None or wt > gmax else gmax
        merged.append(out)
    return [[(j, wt / gmax) for j, wt in row if j != b] for b, row in enumerate(merged)]


def build_partition(row_ptr, col_idx, near_null, weights, coarsening_factor: float = 8.0, agg_size_penalty: float = 1.0,
                    max_improvement_iters: int = 100, max_depth: int = 3, block_size: int = 1):
    """``PartitionerConfig::build_partition`` (mod.rs:273-329) -> (node_to_agg, strength)."""
    strength = new_ls_strength_graph(row_ptr, col_idx, near_null, weights, max_depth)
    if block_size > 1:
        strength = block_reduce(strength, block_size)
    p = Partitioner(strength, coarsening_factor, agg_size_penalty, max_improvement_iters)
    p.initialize_partition()
    p.improve_partition()
    return np.asarray(p.node_to_agg, dtype=np.int64), strength
