"""Golden vectors for the path enumeration (algp_b200/paths.py, csrc/paths.cu).

Runs the reference's UNMODIFIED env.py / map.py / graph_utils.py from /root/reference in this container and
freezes what FieldEnv.get_all_paths returns, together with the planning graph it searched (after
_pre_search inserted the start and the waypoints) so that the test can replay the search without the
reference.  Only import plumbing is shimmed: plotting / debugger modules that are absent here become empty
modules, `from networkx import nx` and `graph.node` (networkx < 2.4 spellings, env.py:5,219) are aliased onto
the installed networkx 3.x, and gpytorch is the stand-in used by make_golden.py (never called on this path).

    python tests/golden/make_golden_paths.py        # writes tests/golden/ref_paths.npz
"""
import os
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
for name in ["seaborn", "ipdb", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors", "matplotlib.cm"]:
    sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
import networkx  # noqa: E402
networkx.nx = networkx
networkx.Graph.node = property(lambda self: self.nodes)
networkx.DiGraph.node = property(lambda self: self.nodes)
sys.path.insert(0, HERE)
import gpytorch_standin  # noqa: E402
sys.modules["gpytorch"] = gpytorch_standin
sys.path.insert(0, "/root/reference")
import env as E  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from algp_b200 import paths as P  # noqa: E402


def run_case(e, start, heading, waypoints, slack):
    """The reference call, plus a replay of its own preamble to capture the searched graph and the bound."""
    t0 = time.perf_counter()
    ref_paths, ref_idx, ref_cost = e.get_all_paths(start, heading, waypoints, slack=slack)
    t_ref = time.perf_counter() - t0
    e._pre_search(start, waypoints)
    least = e.get_heuristic_cost(start, heading, waypoints)
    nodes, pos, rc, adj_ptr, adj, eptr, eidx = P.graph_arrays(e.graph)
    e._post_search()
    case = {
        "rc": rc, "adj_ptr": adj_ptr, "adj": adj, "eptr": eptr, "eidx": eidx,
        "start": np.int32(pos[tuple(start)]), "heading": np.array(heading, dtype=np.int32),
        "waypoints": np.array([pos[tuple(w)] for w in waypoints], dtype=np.int32),
        "least_cost": np.float64(least), "slack": np.float64(slack),
        "path_ptr": np.cumsum([0] + [len(p) for p in ref_paths]).astype(np.int64),
        "path_nodes": np.array([pos[tuple(n)] for p in ref_paths for n in p], dtype=np.int32),
        "idx_ptr": np.cumsum([0] + [len(p) for p in ref_idx]).astype(np.int64),
        "idx": np.array([int(i) for p in ref_idx for i in p], dtype=np.int32),
        "cost": np.array([float(c) for c in ref_cost], dtype=np.float64),
    }
    # the drop-in against the live reference environment, list for list
    t0 = time.perf_counter()
    got_paths, got_idx, got_cost = P.get_all_paths(e, start, heading, waypoints, slack=slack)
    t_new = time.perf_counter() - t0
    t0 = time.perf_counter()
    ps = P.get_all_paths(e, start, heading, waypoints, slack=slack, return_set=True)
    slots = ps.slots()
    t_set = time.perf_counter() - t0
    print("    reference %.1f ms, drop-in (lists) %.1f ms, drop-in (slot matrix %s) %.1f ms" % (
        t_ref * 1e3, t_new * 1e3, slots.shape, t_set * 1e3))
    assert [[tuple(map(int, n)) for n in p] for p in got_paths] == [[tuple(map(int, n)) for n in p] for p in ref_paths]
    assert [[int(i) for i in p] for p in got_idx] == [[int(i) for i in p] for p in ref_idx]
    assert [float(c) for c in got_cost] == [float(c) for c in ref_cost]
    return case


def main():
    np.random.seed(3)
    e = E.FieldEnv()                       # the reference's default synthetic 30 x 30 field (env.py:20-26)
    rng = np.random.default_rng(7)
    cases = {}
    specs = [((0, 0), (1, 0), 3, 0), ((0, 0), (1, 0), 3, 2), ((0, 0), (0, 1), 4, 4), ((0, 0), (1, 0), 1, 6),
             ((0, 0), (1, 0), 5, 0), ((0, 0), (1, 0), 2, 8)]
    for k, (start, heading, nw, slack) in enumerate(specs):
        idx = rng.choice(e.num_samples, nw, replace=False)
        wps = [tuple(int(v) for v in e.gp_index_to_map_pose(i)) for i in idx]
        case = run_case(e, start, heading, wps, slack)
        print("case %d: %d waypoints, slack %d -> %d paths, %d sample indices" % (k, nw, slack, len(case["cost"]), len(case["idx"])))
        for name, v in case.items():
            cases["c%d_%s" % (k, name)] = v
    # a start in the middle of a corridor (not a junction), heading down the row
    wp0 = tuple(int(v) for v in e.gp_index_to_map_pose(int(rng.integers(e.num_samples))))
    mid = (wp0[0], wp0[1])
    others = [tuple(int(v) for v in e.gp_index_to_map_pose(i)) for i in rng.choice(e.num_samples, 3, replace=False)]
    others = [w for w in others if w != mid]
    case = run_case(e, mid, (1, 0), others, 8)
    k = len(specs)
    print("case %d: corridor start %s -> %d paths" % (k, mid, len(case["cost"])))
    for name, v in case.items():
        cases["c%d_%s" % (k, name)] = v
    cases["n_cases"] = np.int32(k + 1)
    np.savez_compressed(os.path.join(HERE, "ref_paths.npz"), **cases)


if __name__ == "__main__":
    main()
