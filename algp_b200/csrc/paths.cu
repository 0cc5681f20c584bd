// Path enumeration feeding Agent.best_path: the expansion-tree search of FieldEnv.get_all_paths
// (reference env.py:197-310) as host C++ behind the C ABI.  No GPU work: this is the caller-side stage that
// dominates a planning step once the scoring itself takes a millisecond.
//
// Same algorithm, same order of results:
//   * breadth-first expansion over the planning graph (FIFO open list, neighbours in the graph's adjacency
//     order), U-turns forbidden (graph_utils.py:8-13), children pruned when
//         g + lower_bound_path_cost(pose, unvisited waypoints) > least_cost + slack   (env.py:241-244,
//         graph_utils.py:49-63), least_cost tightened whenever a node visits every waypoint;
//   * a child equal to an existing tree node in (pose, heading, visited, g) is merged into it
//     (graph_utils.py:128-134): the reference scans the whole tree per child, here it is one hash lookup;
//   * every closed node is expanded into ALL root paths of the expansion DAG, in the order
//     networkx.all_shortest_paths yields them (Dijkstra predecessor lists in heap-pop order, then the
//     stack walk of _build_paths_from_predecessors), paths costlier than the final least_cost + slack dropped
//     (env.py:278-300);
//   * a path's sample set is the concatenation of the `indices` lists of its graph edges (env.py:292).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <queue>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "common.cuh"

namespace {

struct TNode {
  int32_t pose;        // graph node
  int32_t hr, hc;      // heading
  uint64_t visited;    // bit i = waypoint i reached
  double g;
};

struct Key {
  int32_t pose, hr, hc;
  uint64_t visited;
  double g;
  bool operator==(const Key& o) const { return pose == o.pose && hr == o.hr && hc == o.hc && visited == o.visited && g == o.g; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    uint64_t h = 1469598103934665603ull;
    auto mix = [&h](uint64_t v) { h ^= v; h *= 1099511628211ull; h ^= h >> 29; };
    mix((uint32_t)k.pose); mix((uint32_t)(k.hr + 2) * 5u + (uint32_t)(k.hc + 2)); mix(k.visited);
    uint64_t gb; memcpy(&gb, &k.g, 8); mix(gb);
    return (size_t)h;
  }
};

struct PathsResult {
  std::vector<int64_t> path_ptr{0};   // [n_paths + 1] into path_nodes
  std::vector<int32_t> path_nodes;    // graph nodes along each path
  std::vector<int64_t> idx_ptr{0};    // [n_paths + 1] into idx
  std::vector<int32_t> idx;           // field-location indices sampled along each path
  std::vector<double> cost;           // [n_paths]
  int64_t tree_nodes = 0, merged = 0, closed = 0;
  double least_cost = 0.0;
};

inline int sgn(int v) { return (v > 0) - (v < 0); }

}  // namespace

extern "C" int algp_paths_enumerate(int32_t n_nodes, const int32_t* node_rc, const int64_t* adj_ptr, const int32_t* adj,
                                    const int64_t* eidx_ptr, const int32_t* eidx, int32_t start_node, int32_t heading_r,
                                    int32_t heading_c, const int32_t* waypoint_nodes, int32_t n_waypoints, double least_cost,
                                    double slack, int64_t max_tree_nodes, void** handle_out) {
  if (!node_rc || !adj_ptr || !adj || !eidx_ptr || !handle_out || n_nodes <= 0 || start_node < 0 || start_node >= n_nodes ||
      n_waypoints < 0 || (n_waypoints > 0 && !waypoint_nodes))
    return ALGP_ERR_INVALID;
  if (n_waypoints > 64) return ALGP_ERR_UNSUPPORTED;
  for (int i = 0; i < n_waypoints; ++i)
    if (waypoint_nodes[i] < 0 || waypoint_nodes[i] >= n_nodes) return ALGP_ERR_INVALID;
  *handle_out = nullptr;
  const int nw = n_waypoints;
  const uint64_t all_visited = nw == 64 ? ~0ull : ((1ull << nw) - 1);
  // `new_pose in waypoints` + `waypoints.index(new_pose)`: the first waypoint at that node (env.py:236-237)
  std::vector<int32_t> first_wp(n_nodes, -1);
  for (int i = nw - 1; i >= 0; --i) first_wp[waypoint_nodes[i]] = i;

  std::vector<TNode> tree;
  std::vector<std::vector<std::pair<int32_t, double>>> succ;   // tree edges, insertion order (networkx _adj order)
  std::unordered_map<Key, int32_t, KeyHash> index;
  std::vector<int32_t> open, closed;
  tree.push_back({start_node, heading_r, heading_c, 0ull, 0.0});
  succ.emplace_back();
  index.emplace(Key{start_node, heading_r, heading_c, 0ull, 0.0}, 0);
  open.push_back(0);
  PathsResult* res = new PathsResult();
  size_t head = 0;
  while (head < open.size()) {
    const int32_t parent = open[head++];
    const TNode pn = tree[parent];
    const int pr = node_rc[2 * pn.pose], pc = node_rc[2 * pn.pose + 1];
    for (int64_t e = adj_ptr[pn.pose]; e < adj_ptr[pn.pose + 1]; ++e) {
      const int32_t np = adj[e];
      const int nr = node_rc[2 * np], nc = node_rc[2 * np + 1];
      const int dr = nr - pr, dc = nc - pc;
      if (dr == 0 && dc == 0) continue;                       // get_heading -> None: not an edge of a planning graph
      // get_heading (graph_utils.py:16-27): along columns when the rows agree, else along rows
      const int nhr = dr == 0 ? 0 : sgn(dr), nhc = dr == 0 ? sgn(dc) : 0;
      if (pn.hr * nhr + pn.hc * nhc == -1) continue;          // U-turn: cost inf (graph_utils.py:8-13, 66-71)
      const double cost = (double)(abs(dr) + abs(dc));        // manhattan_distance
      const double ng = pn.g + cost;
      uint64_t nv = pn.visited;
      if (first_wp[np] >= 0) nv |= 1ull << first_wp[np];
      // lower_bound_path_cost (graph_utils.py:49-63): bounding box of the unvisited waypoints and the pose
      int minr = nr, maxr = nr, minc = nc, maxc = nc;
      for (int i = 0; i < nw; ++i)
        if (!((nv >> i) & 1)) {
          const int wr = node_rc[2 * waypoint_nodes[i]], wc = node_rc[2 * waypoint_nodes[i] + 1];
          minr = std::min(minr, wr); maxr = std::max(maxr, wr);
          minc = std::min(minc, wc); maxc = std::max(maxc, wc);
        }
      const int a0 = nr - minr, a1 = nc - minc, b0 = maxr - nr, b1 = maxc - nc;
      const double togo = (double)(a0 + a1 + b0 + b1 + std::min(a0, b0) + std::min(a1, b1));
      if (ng + togo > least_cost + slack) continue;
      const Key key{np, nhr, nhc, nv, ng};
      auto it = index.find(key);
      if (it != index.end()) {                                // find_merge_to_node
        succ[parent].emplace_back(it->second, cost);
        ++res->merged;
        continue;
      }
      if ((int64_t)tree.size() >= max_tree_nodes) {
        delete res;
        return ALGP_ERR_UNSUPPORTED;                          // search budget exhausted
      }
      const int32_t id = (int32_t)tree.size();
      tree.push_back({np, nhr, nhc, nv, ng});
      succ.emplace_back();
      index.emplace(key, id);
      succ[parent].emplace_back(id, cost);
      if (nv == all_visited) {
        least_cost = std::min(ng, least_cost);
        closed.push_back(id);
      } else {
        open.push_back(id);
      }
    }
  }
  res->tree_nodes = (int64_t)tree.size();
  res->closed = (int64_t)closed.size();
  res->least_cost = least_cost;

  // networkx.dijkstra_predecessor_and_distance(tree, root, weight='weight'): pred[u] in heap-pop order of the
  // predecessors; ties in distance are broken by the push counter
  const int T = (int)tree.size();
  std::vector<std::vector<int32_t>> pred(T);
  {
    std::vector<double> dist(T, -1.0), seen(T, -1.0);
    typedef std::tuple<double, int64_t, int32_t> Item;
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> heap;
    int64_t counter = 0;
    seen[0] = 0.0;
    heap.emplace(0.0, counter++, 0);
    while (!heap.empty()) {
      const Item top = heap.top();
      heap.pop();
      const double d = std::get<0>(top);
      const int32_t v = std::get<2>(top);
      if (dist[v] >= 0.0) continue;
      dist[v] = d;
      for (const auto& ed : succ[v]) {
        const int32_t u = ed.first;
        const double vu = d + ed.second;
        if (dist[u] >= 0.0) {
          if (vu == dist[u]) pred[u].push_back(v);
        } else if (seen[u] < 0.0 || vu < seen[u]) {
          seen[u] = vu;
          heap.emplace(vu, counter++, u);
          pred[u].assign(1, v);
        } else if (vu == seen[u]) {
          pred[u].push_back(v);
        }
      }
    }
  }

  // networkx _build_paths_from_predecessors({root}, target, pred), target by target in closing order
  std::vector<std::pair<int32_t, int32_t>> stack;
  std::vector<uint8_t> in_seen(T, 0);
  for (const int32_t target : closed) {
    if (tree[target].g > least_cost + slack) continue;         // env.py:284-285 (every path to it costs g)
    stack.clear();
    stack.emplace_back(target, 0);
    in_seen[target] = 1;
    int top = 0;
    while (top >= 0) {
      const int32_t node = stack[top].first;
      const int32_t i = stack[top].second;
      if (node == 0) {
        // a root path: stack[0..top] reversed
        for (int s = top; s >= 0; --s) res->path_nodes.push_back(tree[stack[s].first].pose);
        res->path_ptr.push_back((int64_t)res->path_nodes.size());
        for (int s = top; s > 0; --s) {
          const int32_t u = tree[stack[s].first].pose, w = tree[stack[s - 1].first].pose;
          for (int64_t e = adj_ptr[u]; e < adj_ptr[u + 1]; ++e)
            if (adj[e] == w) {
              if (eidx) res->idx.insert(res->idx.end(), eidx + eidx_ptr[e], eidx + eidx_ptr[e + 1]);
              break;
            }
        }
        res->idx_ptr.push_back((int64_t)res->idx.size());
        res->cost.push_back(tree[target].g);
      }
      if ((int32_t)pred[node].size() > i) {
        stack[top].second = i + 1;
        const int32_t next = pred[node][i];
        if (in_seen[next]) continue;
        in_seen[next] = 1;
        ++top;
        if (top == (int)stack.size()) stack.emplace_back(next, 0);
        else stack[top] = std::make_pair(next, 0);
      } else {
        in_seen[node] = 0;
        --top;
      }
    }
  }
  *handle_out = res;
  return ALGP_OK;
}

// sizes[0..5] = {paths, total path nodes, total indices, tree nodes, merged children, longest index list}
extern "C" int algp_paths_sizes(const void* handle, int64_t* sizes, double* least_cost) {
  if (!handle || !sizes) return ALGP_ERR_INVALID;
  const PathsResult* r = (const PathsResult*)handle;
  sizes[0] = (int64_t)r->cost.size();
  sizes[1] = (int64_t)r->path_nodes.size();
  sizes[2] = (int64_t)r->idx.size();
  sizes[3] = r->tree_nodes;
  sizes[4] = r->merged;
  int64_t longest = 0;
  for (size_t p = 0; p + 1 < r->idx_ptr.size(); ++p) longest = std::max(longest, r->idx_ptr[p + 1] - r->idx_ptr[p]);
  sizes[5] = longest;
  if (least_cost) *least_cost = r->least_cost;
  return ALGP_OK;
}

// copies the ragged results into caller buffers sized from algp_paths_sizes (any pointer may be NULL)
extern "C" int algp_paths_fetch(const void* handle, int64_t* path_ptr, int32_t* path_nodes, int64_t* idx_ptr, int32_t* idx,
                                double* cost) {
  if (!handle) return ALGP_ERR_INVALID;
  const PathsResult* r = (const PathsResult*)handle;
  if (path_ptr) std::copy(r->path_ptr.begin(), r->path_ptr.end(), path_ptr);
  if (path_nodes) std::copy(r->path_nodes.begin(), r->path_nodes.end(), path_nodes);
  if (idx_ptr) std::copy(r->idx_ptr.begin(), r->idx_ptr.end(), idx_ptr);
  if (idx) std::copy(r->idx.begin(), r->idx.end(), idx);
  if (cost) std::copy(r->cost.begin(), r->cost.end(), cost);
  return ALGP_OK;
}

// the [paths x k] slot matrix Agent.best_path / algp_score_sets take (-1 = empty slot), k >= the longest list
extern "C" int algp_paths_fill_slots(const void* handle, int32_t* slots, int64_t k) {
  if (!handle || !slots) return ALGP_ERR_INVALID;
  const PathsResult* r = (const PathsResult*)handle;
  const int64_t P = (int64_t)r->cost.size();
  for (int64_t p = 0; p < P; ++p) {
    const int64_t len = r->idx_ptr[p + 1] - r->idx_ptr[p];
    if (len > k) return ALGP_ERR_INVALID;
    int32_t* row = slots + p * k;
    std::copy(r->idx.begin() + r->idx_ptr[p], r->idx.begin() + r->idx_ptr[p + 1], row);
    std::fill(row + len, row + k, -1);
  }
  return ALGP_OK;
}

extern "C" int algp_paths_free(void* handle) {
  delete (PathsResult*)handle;
  return ALGP_OK;
}
