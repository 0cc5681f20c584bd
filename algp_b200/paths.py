"""Path enumeration for Agent.best_path: FieldEnv.get_all_paths (reference env.py:197-310) with the
expansion-tree search in native code (csrc/paths.cu, C ABI algp_paths_*).

The planning graph, the waypoint insertion (`_pre_search`, env.py:138-150), the heuristic bound
(`get_heuristic_cost`, env.py:312-382) and the graph restore (`_post_search`) stay the reference's own
methods; what is replaced is the breadth-first expansion with its O(tree) merge scan per child
(graph_utils.py:128-134) and the per-target `nx.all_shortest_paths` calls.  Results come back in the
reference's order: same paths, same sample-index lists, same costs.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import call


class PathSet(object):
    """Ragged result of one enumeration: node sequences, sampled field locations and costs per path."""

    def __init__(self, nodes, path_ptr, path_nodes, idx_ptr, idx, cost, stats):
        self.nodes = nodes                  # graph node objects (the (row, col) tuples of the planning graph)
        self.path_ptr, self.path_nodes = path_ptr, path_nodes
        self.idx_ptr, self.idx = idx_ptr, idx
        self.cost = cost
        self.stats = stats                  # tree_nodes, merged, least_cost

    def __len__(self):
        return len(self.cost)

    def locations(self):
        """list of paths, each a list of graph nodes (all_paths of the reference)"""
        pn = self.path_nodes
        return [[self.nodes[v] for v in pn[self.path_ptr[p]:self.path_ptr[p + 1]]] for p in range(len(self))]

    def indices(self):
        """list of per-path sample-index lists (all_paths_indices of the reference)"""
        return [self.idx[self.idx_ptr[p]:self.idx_ptr[p + 1]].tolist() for p in range(len(self))]

    def costs(self):
        return self.cost.tolist()

    def slots(self, k=None):
        """[P, k] int32 matrix (-1 = empty slot): the array form Agent.best_path scores without host lists."""
        lens = np.diff(self.idx_ptr)
        k = int(lens.max()) if k is None and len(lens) else (k or 1)
        out = np.full((len(self), max(1, k)), -1, dtype=np.int32)
        if len(self.idx):
            rows = np.repeat(np.arange(len(self)), lens)
            cols = np.arange(len(self.idx)) - np.repeat(self.idx_ptr[:-1], lens)
            out[rows, cols] = self.idx
        return out


def graph_arrays(graph):
    """CSR view of a networkx planning graph: nodes and neighbours in the graph's own iteration order
    (what `graph.neighbors(pose)` yields, env.py:224), per-edge `indices` lists (env.py:292)."""
    nodes = list(graph.nodes())
    pos = {n: i for i, n in enumerate(nodes)}
    rc = np.array([[int(n[0]), int(n[1])] for n in nodes], dtype=np.int32).reshape(-1, 2)
    adj_ptr = np.zeros(len(nodes) + 1, dtype=np.int64)
    adj, eptr, eidx = [], [0], []
    for i, n in enumerate(nodes):
        for m, data in graph[n].items():
            adj.append(pos[m])
            eidx.extend(int(v) for v in data.get('indices', ()))
            eptr.append(len(eidx))
        adj_ptr[i + 1] = len(adj)
    return (nodes, pos, rc, adj_ptr, np.asarray(adj, dtype=np.int32), np.asarray(eptr, dtype=np.int64),
            np.asarray(eidx, dtype=np.int32))


def enumerate_paths(graph, start, heading, waypoints, least_cost, slack=0, max_tree_nodes=50_000_000):
    """The expansion-tree search of env.py:197-300 on `graph` (start and waypoints already inserted)."""
    nodes, pos, rc, adj_ptr, adj, eptr, eidx = graph_arrays(graph)
    wp = [pos[tuple(w)] for w in waypoints]
    return enumerate_paths_arrays(nodes, rc, adj_ptr, adj, eptr, eidx, pos[tuple(start)], heading, wp, least_cost, slack,
                                  max_tree_nodes)


def enumerate_paths_arrays(nodes, rc, adj_ptr, adj, eptr, eidx, start_idx, heading, waypoint_idx, least_cost, slack=0,
                           max_tree_nodes=50_000_000):
    """Same search on a graph already in CSR form (see graph_arrays / include/algp_b200.h)."""
    rc = np.ascontiguousarray(rc, dtype=np.int32)
    adj_ptr = np.ascontiguousarray(adj_ptr, dtype=np.int64)
    adj = np.ascontiguousarray(adj, dtype=np.int32)
    eptr = np.ascontiguousarray(eptr, dtype=np.int64)
    eidx = np.ascontiguousarray(eidx, dtype=np.int32)
    wp = np.ascontiguousarray(waypoint_idx, dtype=np.int32)
    handle = C.c_void_p()
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    call("algp_paths_enumerate", len(rc), p(rc), p(adj_ptr), p(adj), p(eptr), p(eidx), int(start_idx),
         int(heading[0]), int(heading[1]), p(wp), len(wp), float(least_cost), float(slack), int(max_tree_nodes),
         C.byref(handle))
    try:
        sizes = np.zeros(6, dtype=np.int64)
        lc = C.c_double()
        call("algp_paths_sizes", handle, p(sizes), C.byref(lc))
        P, npn, nidx = int(sizes[0]), int(sizes[1]), int(sizes[2])
        path_ptr = np.zeros(P + 1, dtype=np.int64)
        path_nodes = np.zeros(max(1, npn), dtype=np.int32)
        idx_ptr = np.zeros(P + 1, dtype=np.int64)
        idx = np.zeros(max(1, nidx), dtype=np.int32)
        cost = np.zeros(max(1, P), dtype=np.float64)
        call("algp_paths_fetch", handle, p(path_ptr), p(path_nodes), p(idx_ptr), p(idx), p(cost))
    finally:
        call("algp_paths_free", handle)
    stats = {"tree_nodes": int(sizes[3]), "merged": int(sizes[4]), "least_cost": float(lc.value)}
    return PathSet(nodes, path_ptr, path_nodes[:npn], idx_ptr, idx[:nidx], cost[:P], stats)


def get_all_paths(env, start, heading, waypoints, heuristic_cost=None, slack=0, return_set=False):
    """Drop-in for FieldEnv.get_all_paths(start, heading, waypoints, heuristic_cost=None, slack=0)
    (env.py:197): returns (all_paths, all_paths_indices, all_paths_cost); `return_set=True` returns the
    PathSet instead (use `.slots()` to hand the whole batch to Agent.best_path as one array)."""
    env._pre_search(start, waypoints)
    try:
        least_cost = env.get_heuristic_cost(start, heading, waypoints) if heuristic_cost is None else heuristic_cost
        ps = enumerate_paths(env.graph, start, heading, waypoints, least_cost, slack)
    finally:
        env._post_search()
    if return_set:
        return ps
    cost = ps.cost
    costs = [int(c) if float(c).is_integer() else float(c) for c in cost]
    return ps.locations(), ps.indices(), costs


def patch_env(env_cls):
    """Install the native search on the reference's FieldEnv class (env.py:15)."""
    def _get_all_paths(self, start, heading, waypoints, heuristic_cost=None, slack=0):
        return get_all_paths(self, start, heading, waypoints, heuristic_cost, slack)
    env_cls.get_all_paths = _get_all_paths
    return env_cls
