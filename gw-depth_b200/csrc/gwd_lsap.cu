// gwd_lsap.cu -- host side of the Hungarian matcher: a batch of rectangular linear-sum-assignment problems solved on
// the host cores, one problem per worker at a time (HOST code; nothing here touches the GPU).
//
// The reference solves them one by one with scipy.optimize.linear_sum_assignment (src/models/matcher.py:74), i.e. the
// shortest-augmenting-path algorithm with dual variables of D. F. Crouse, "On implementing 2D rectangular assignment
// algorithms", IEEE T-AES 52(4), 2016.  This is the same algorithm restated from the paper, including the details that
// decide between EQUAL-cost optima (columns scanned from the last to the first, ties resolved towards unassigned
// columns, the more numerous side transposed to the columns), so that the assignment -- not just its cost -- is the one
// scipy returns; tests/test_lsap_cpu.py checks index equality against scipy on random, tied and degenerate problems.
// A training step has 6 stages x B images of these (100 x T, T ~ 10..50): ~35 us each, 3.5 ms in sequence for B = 16.
#include <math.h>
#include <thread>
#include <vector>
#include <algorithm>
#include <atomic>
#include "gwd_common.cuh"

namespace {

struct Workspace {
  std::vector<double> cost, u, v, spc;
  std::vector<int> path, col4row, row4col, remaining;
  std::vector<char> SR, SC;
};

// rows <= cols.  Returns false when no feasible assignment exists (an infinite / NaN row).
bool solve(int nr, int nc, const double* cost, Workspace& w) {
  w.u.assign(nr, 0.0); w.v.assign(nc, 0.0); w.spc.resize(nc);
  w.path.assign(nc, -1); w.col4row.assign(nr, -1); w.row4col.assign(nc, -1);
  w.SR.resize(nr); w.SC.resize(nc); w.remaining.resize(nc);
  for (int cur = 0; cur < nr; ++cur) {
    // ---- shortest augmenting path from row `cur`
    double min_val = 0.0;
    int i = cur, num_remaining = nc, sink = -1;
    for (int it = 0; it < nc; ++it) w.remaining[it] = nc - it - 1;
    std::fill(w.SR.begin(), w.SR.end(), 0);
    std::fill(w.SC.begin(), w.SC.end(), 0);
    std::fill(w.spc.begin(), w.spc.end(), INFINITY);
    while (sink == -1) {
      int index = -1;
      double lowest = INFINITY;
      w.SR[i] = 1;
      const double* row = cost + static_cast<size_t>(i) * nc;
      for (int it = 0; it < num_remaining; ++it) {
        const int j = w.remaining[it];
        const double r = min_val + row[j] - w.u[i] - w.v[j];
        if (r < w.spc[j]) { w.path[j] = i; w.spc[j] = r; }
        if (w.spc[j] < lowest || (w.spc[j] == lowest && w.row4col[j] == -1)) { lowest = w.spc[j]; index = it; }
      }
      min_val = lowest;
      if (!(min_val < INFINITY)) return false;
      const int j = w.remaining[index];
      if (w.row4col[j] == -1) sink = j; else i = w.row4col[j];
      w.SC[j] = 1;
      w.remaining[index] = w.remaining[--num_remaining];
    }
    // ---- dual update
    w.u[cur] += min_val;
    for (int r = 0; r < nr; ++r)
      if (w.SR[r] && r != cur) w.u[r] += min_val - w.spc[w.col4row[r]];
    for (int j = 0; j < nc; ++j)
      if (w.SC[j]) w.v[j] -= min_val - w.spc[j];
    // ---- augment
    int j = sink;
    while (true) {
      const int r = w.path[j];
      w.row4col[j] = r;
      std::swap(w.col4row[r], j);
      if (r == cur) break;
    }
  }
  return true;
}

// one [Q, T] float32 problem -> pairs sorted by the query index, as scipy returns them
int solve_problem(const float* cost, int Q, int T, int32_t* qi, int32_t* ti, Workspace& w) {
  if (Q == 0 || T == 0) return 0;
  for (size_t e = 0; e < static_cast<size_t>(Q) * T; ++e)
    if (isnan(cost[e]) || cost[e] == -INFINITY) return -1;        // scipy raises "matrix contains invalid numeric entries"
  const bool transpose = T < Q;            // more rows than columns: solve the transposed problem
  const int nr = transpose ? T : Q, nc = transpose ? Q : T;
  w.cost.resize(static_cast<size_t>(nr) * nc);
  if (transpose) {
    for (int q = 0; q < Q; ++q)
      for (int t = 0; t < T; ++t) w.cost[static_cast<size_t>(t) * Q + q] = cost[static_cast<size_t>(q) * T + t];
  } else {
    for (size_t e = 0; e < static_cast<size_t>(Q) * T; ++e) w.cost[e] = cost[e];
  }
  if (!solve(nr, nc, w.cost.data(), w)) return -1;
  if (!transpose) {
    for (int q = 0; q < Q; ++q) { qi[q] = q; ti[q] = w.col4row[q]; }
    return Q;
  }
  // col4row[t] = query of target t; order the pairs by query (a stable argsort of distinct keys)
  std::vector<int> order(T);
  for (int t = 0; t < T; ++t) order[t] = t;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return w.col4row[a] < w.col4row[b]; });
  for (int k = 0; k < T; ++k) { qi[k] = w.col4row[order[k]]; ti[k] = order[k]; }
  return T;
}

}  // namespace

extern "C" int gwd_lsap_batch(const float* cost, const int64_t* cost_offsets, const int32_t* T, int32_t Q, int32_t n_problems,
                              int32_t* query_idx, int32_t* target_idx, int32_t* counts, int32_t n_threads) {
  GWD_CHECK_ARG(cost && cost_offsets && T && query_idx && target_idx && counts && Q >= 0 && n_problems >= 0,
                "gwd_lsap_batch: bad argument");
  if (n_problems == 0) return GWD_OK;
  int workers = n_threads > 0 ? n_threads : static_cast<int>(std::thread::hardware_concurrency());
  workers = std::max(1, std::min(workers, std::min(n_problems, 16)));
  std::atomic<int> next{0}, failed{0};
  auto work = [&]() {
    Workspace w;
    for (int p = next.fetch_add(1); p < n_problems; p = next.fetch_add(1)) {
      const int stride = std::max(Q, 1);
      const int c = solve_problem(cost + cost_offsets[p], Q, T[p], query_idx + static_cast<size_t>(p) * stride,
                                  target_idx + static_cast<size_t>(p) * stride, w);
      counts[p] = c;
      if (c < 0) failed.store(1);
    }
  };
  if (workers == 1) {
    work();
  } else {
    std::vector<std::thread> pool;
    for (int t = 0; t < workers; ++t) pool.emplace_back(work);
    for (auto& t : pool) t.join();
  }
  if (failed.load()) {
    gwd_set_error("gwd_lsap_batch: a cost matrix is infeasible or holds NaN / -inf");
    return GWD_ERR_ARG;
  }
  return GWD_OK;
}
