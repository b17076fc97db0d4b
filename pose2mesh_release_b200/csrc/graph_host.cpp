// Host-side (CPU) native pieces of the graph-hierarchy builder: the sequential greedy heavy-edge
// matching that the reference runs as pure-Python loops (lib/coarsening.py:153-211, ~1 s for SMPL).
// Exported through the same C ABI; pose2mesh_release_b200/graph.py drives it.
#include <cstdint>
#include <vector>

extern "C" {

// One matching level.  Entries (rows[e], cols[e], vals[e]) are sorted by (row, col).
// Semantics follow the reference exactly, including
//   * its check-after-increment row-length bookkeeping (row 0 scans one entry too many, the last row
//     one too few; lib/coarsening.py:166-171),
//   * W_ii / W_jj taken from the FIRST stored entry of row i / j (:191-192),
//   * strict '>' when choosing the best neighbour (:197).
// Returns the number of clusters, or -1 on bad input.
int32_t p2m_graph_match_level(int64_t nnz, const int32_t* rows, const int32_t* cols, const double* vals,
                              const int64_t* visit_order, const double* weights, int32_t* cluster_out) {
  if (nnz <= 0 || !rows || !cols || !vals || !visit_order || !weights || !cluster_out) return -1;
  const int64_t n = (int64_t)rows[nnz - 1] + 1;
  std::vector<int64_t> start(n, 0), length(n, 0);
  std::vector<char> taken(n, 0);
  int32_t prev = rows[0];
  int64_t slot = 0;
  for (int64_t e = 0; e < nnz; ++e) {
    length[slot] += 1;
    if (rows[e] > prev) {
      prev = rows[e];
      if (slot + 1 >= n) return -1;
      start[slot + 1] = e;
      slot += 1;
    }
  }
  for (int64_t i = 0; i < n; ++i) cluster_out[i] = 0;
  int32_t n_cluster = 0;
  for (int64_t t = 0; t < n; ++t) {
    const int64_t u = visit_order[t];
    if (u < 0 || u >= n) return -1;
    if (taken[u]) continue;
    taken[u] = 1;
    int64_t best = -1;
    double best_score = 0.0;
    const int64_t s = start[u];
    for (int64_t q = 0; q < length[u]; ++q) {
      if (s + q >= nnz) break;
      const int64_t v = cols[s + q];
      double score = 0.0;
      if (!taken[v]) {
        score = (2.0 * vals[s + q] + vals[start[u]] + vals[start[v]]) * 1.0 / (weights[u] + weights[v] + 1e-9);
      }
      if (score > best_score) {
        best_score = score;
        best = v;
      }
    }
    cluster_out[u] = n_cluster;
    if (best >= 0) {
      cluster_out[best] = n_cluster;
      taken[best] = 1;
    }
    n_cluster += 1;
  }
  return n_cluster;
}

}  // extern "C"
