"""Path enumeration (algp_b200/paths.py, csrc/paths.cu) against what the reference's FieldEnv.get_all_paths
returned on the same planning graphs (tests/golden/ref_paths.npz, made by tests/golden/make_golden_paths.py
from the unmodified env.py / map.py / graph_utils.py).  Host code only: no GPU needed."""
import ctypes as C
import os

import numpy as np
import pytest

from algp_b200 import _lib, paths as P

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_paths.npz")


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


def _case(g, k):
    return {name: g["c%d_%s" % (k, name)] for name in ("rc", "adj_ptr", "adj", "eptr", "eidx", "start", "heading", "waypoints",
                                                        "least_cost", "slack", "path_ptr", "path_nodes", "idx_ptr", "idx", "cost")}


def _run(c, **kw):
    nodes = [tuple(r) for r in c["rc"].tolist()]
    return P.enumerate_paths_arrays(nodes, c["rc"], c["adj_ptr"], c["adj"], c["eptr"], c["eidx"], int(c["start"]),
                                    tuple(c["heading"].tolist()), c["waypoints"], float(c["least_cost"]), float(c["slack"]), **kw)


def test_paths_match_reference_in_order(golden):
    """Same paths, same sample-index lists, same costs, in the order the reference returns them."""
    n = int(golden["n_cases"])
    assert n >= 7
    total = 0
    for k in range(n):
        c = _case(golden, k)
        ps = _run(c)
        np.testing.assert_array_equal(ps.path_ptr, c["path_ptr"])
        np.testing.assert_array_equal(ps.path_nodes, c["path_nodes"])
        np.testing.assert_array_equal(ps.idx_ptr, c["idx_ptr"])
        np.testing.assert_array_equal(ps.idx, c["idx"])
        np.testing.assert_array_equal(ps.cost, c["cost"])
        assert ps.stats["least_cost"] <= float(c["least_cost"])
        assert (ps.cost <= ps.stats["least_cost"] + float(c["slack"])).all()          # env.py:284
        total += len(ps)
    assert total > 1000


def test_paths_are_walks_that_visit_every_waypoint(golden):
    """Size-independent properties: consecutive nodes are graph neighbours, no U-turns, every waypoint is on the
    path, the cost is the Manhattan length, the index list is the concatenation of the edge lists."""
    c = _case(golden, 4)
    ps = _run(c)
    rc, adj_ptr, adj, eptr, eidx = c["rc"], c["adj_ptr"], c["adj"], c["eptr"], c["eidx"]
    for p in range(0, len(ps), 37):
        nodes = ps.path_nodes[ps.path_ptr[p]:ps.path_ptr[p + 1]]
        assert nodes[0] == int(c["start"]) and set(c["waypoints"].tolist()) <= set(nodes.tolist())
        length, want_idx, prev_h = 0, [], tuple(c["heading"].tolist())
        for u, v in zip(nodes[:-1], nodes[1:]):
            nb = adj[adj_ptr[u]:adj_ptr[u + 1]].tolist()
            assert v in nb
            e = adj_ptr[u] + nb.index(v)
            want_idx += eidx[eptr[e]:eptr[e + 1]].tolist()
            d = rc[v] - rc[u]
            h = (0, int(np.sign(d[1]))) if d[0] == 0 else (int(np.sign(d[0])), 0)
            assert h[0] * prev_h[0] + h[1] * prev_h[1] != -1
            prev_h = h
            length += abs(int(d[0])) + abs(int(d[1]))
        assert length == ps.cost[p]
        assert want_idx == ps.idx[ps.idx_ptr[p]:ps.idx_ptr[p + 1]].tolist()


def test_slot_matrix_forms_agree(golden):
    c = _case(golden, 2)
    ps = _run(c)
    lists = ps.indices()
    slots = ps.slots()
    assert slots.shape == (len(ps), max(len(l) for l in lists)) and slots.dtype == np.int32
    for p, l in enumerate(lists):
        assert slots[p, :len(l)].tolist() == l and (slots[p, len(l):] == -1).all()
    wide = ps.slots(k=slots.shape[1] + 5)
    assert (wide[:, :slots.shape[1]] == slots).all() and (wide[:, slots.shape[1]:] == -1).all()
    # the C-ABI filler writes the same matrix
    nodes = [tuple(r) for r in c["rc"].tolist()]
    p_ = lambda a: a.ctypes.data_as(C.c_void_p)
    h = C.c_void_p()
    arrs = [np.ascontiguousarray(c[k], dtype=t) for k, t in (("rc", np.int32), ("adj_ptr", np.int64), ("adj", np.int32),
                                                              ("eptr", np.int64), ("eidx", np.int32), ("waypoints", np.int32))]
    _lib.call("algp_paths_enumerate", len(nodes), p_(arrs[0]), p_(arrs[1]), p_(arrs[2]), p_(arrs[3]), p_(arrs[4]),
              int(c["start"]), int(c["heading"][0]), int(c["heading"][1]), p_(arrs[5]), len(arrs[5]), float(c["least_cost"]),
              float(c["slack"]), 1 << 30, C.byref(h))
    try:
        sizes = np.zeros(6, dtype=np.int64)
        _lib.call("algp_paths_sizes", h, p_(sizes), None)
        assert sizes[0] == len(ps) and sizes[5] == slots.shape[1] and sizes[3] == ps.stats["tree_nodes"]
        out = np.zeros((len(ps), int(sizes[5])), dtype=np.int32)
        _lib.call("algp_paths_fill_slots", h, p_(out), int(sizes[5]))
        np.testing.assert_array_equal(out, slots)
        with pytest.raises(_lib.AlgpError):
            _lib.call("algp_paths_fill_slots", h, p_(out), int(sizes[5]) - 1)        # too narrow
    finally:
        _lib.call("algp_paths_free", h)


def _grid_graph(R, Cc):
    """R x Cc lattice with unit spacing 2 between columns, 3 between rows; every edge samples one location."""
    nodes = [(3 * r, 2 * c) for r in range(R) for c in range(Cc)]
    pos = {n: i for i, n in enumerate(nodes)}
    adj_ptr, adj, eptr, eidx = [0], [], [0], []
    for (r, c) in nodes:
        for dr, dc in ((0, 2), (0, -2), (3, 0), (-3, 0)):
            m = (r + dr, c + dc)
            if m in pos:
                adj.append(pos[m])
                eidx.append(min(pos[m], pos[(r, c)]) * 4 + (0 if dr == 0 else 1))      # same list both directions
                eptr.append(len(eidx))
        adj_ptr.append(len(adj))
    return nodes, np.array(nodes, np.int32), np.array(adj_ptr, np.int64), np.array(adj, np.int32), np.array(eptr, np.int64), \
        np.array(eidx, np.int32), pos


def test_edge_cases():
    nodes, rc, adj_ptr, adj, eptr, eidx, pos = _grid_graph(4, 5)
    run = lambda start, heading, wps, least, slack=0, **kw: P.enumerate_paths_arrays(
        nodes, rc, adj_ptr, adj, eptr, eidx, pos[start], heading, [pos[w] for w in wps], least, slack, **kw)
    # no waypoints: every first move closes at once (sum([]) == 0 == nw, env.py:270); the bound prunes the rest
    ps = run((0, 0), (1, 0), [], 0, 3)
    assert sorted(ps.costs()) == [2.0, 3.0] and all(len(l) == 2 for l in ps.locations())
    # nothing reachable within the bound: empty result, not an error
    ps = run((0, 0), (1, 0), [(9, 8)], 3, 0)
    assert len(ps) == 0 and ps.slots().shape == (0, 1) and ps.indices() == []
    # the only way to the waypoint is a U-turn at the start: forbidden
    ps = run((3, 0), (1, 0), [(0, 0)], 3, 0)
    assert len(ps) == 0
    # a straight run: one path, cost = Manhattan length, one sample per edge
    ps = run((0, 0), (1, 0), [(9, 0)], 9, 0)
    assert ps.costs() == [9.0] and ps.locations() == [[(0, 0), (3, 0), (6, 0), (9, 0)]] and len(ps.indices()[0]) == 3
    # two equal-cost routes to a diagonal neighbour
    ps = run((0, 0), (1, 0), [(3, 2)], 5, 0)
    assert sorted(map(tuple, ps.locations())) == [((0, 0), (0, 2), (3, 2)), ((0, 0), (3, 0), (3, 2))]
    # a duplicated waypoint is only ever marked at its first position (waypoints.index, env.py:237): never closes
    ps = run((0, 0), (1, 0), [(3, 0), (3, 0)], 3, 0)
    assert len(ps) == 0
    # search budget and waypoint-count limits are reported, not silently truncated
    with pytest.raises(_lib.AlgpError):
        run((0, 0), (1, 0), [(9, 8)], 17, 12, max_tree_nodes=10)
    with pytest.raises(_lib.AlgpError):
        run((0, 0), (1, 0), [(3, 2)] * 65, 5, 0)


def test_slack_only_adds_paths(golden):
    """Monotone in the slack: a larger budget keeps every path of the smaller one (same least cost)."""
    c = _case(golden, 0)
    ps0 = _run(c)
    c2 = dict(c); c2["slack"] = np.float64(float(c["slack"]) + 4)
    ps4 = _run(c2)
    key = lambda ps: {tuple(ps.path_nodes[ps.path_ptr[p]:ps.path_ptr[p + 1]].tolist()) for p in range(len(ps))}
    assert key(ps0) <= key(ps4) and len(ps4) > len(ps0)
