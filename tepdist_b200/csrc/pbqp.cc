#include "pbqp.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>
#include <limits>

namespace tepdist {
namespace {
constexpr double kInf = 1e30;
inline double sat(double a, double b) { return (a >= kInf || b >= kInf) ? kInf : a + b; }

struct State {
  std::vector<PBQP::Vec> cost;
  std::vector<std::map<int, PBQP::Mat>> adj;
  std::vector<char> alive;
  double constant = 0;  // cost already committed by R0 eliminations and fixed nodes
};

struct Elim {  // how to recover the choice of an eliminated node
  int node;
  int kind;  // 0: independent, 1: depends on a, 2: depends on (a, b), 3: fixed
  int a = -1, b = -1;
  std::vector<int> best1;               // [opt_a] -> opt
  std::vector<std::vector<int>> best2;  // [opt_a][opt_b] -> opt
  int fixed = 0;
};

void RemoveNode(State& s, int u) {
  for (auto& kv : s.adj[u]) s.adj[kv.first].erase(u);
  s.adj[u].clear();
  s.alive[u] = 0;
}

void AddEdgeTo(State& s, int u, int v, const PBQP::Mat& m) {
  auto it = s.adj[u].find(v);
  if (it == s.adj[u].end()) {
    s.adj[u][v] = m;
    PBQP::Mat t(m[0].size(), PBQP::Vec(m.size()));
    for (size_t i = 0; i < m.size(); ++i)
      for (size_t j = 0; j < m[0].size(); ++j) t[j][i] = m[i][j];
    s.adj[v][u] = t;
  } else {
    auto& a = it->second;
    auto& b = s.adj[v][u];
    for (size_t i = 0; i < m.size(); ++i)
      for (size_t j = 0; j < m[0].size(); ++j) {
        a[i][j] = sat(a[i][j], m[i][j]);
        b[j][i] = a[i][j];
      }
  }
}

// Apply R0/R1/R2 until none applies.
void Reduce(State& s, std::vector<Elim>& stack) {
  bool progress = true;
  while (progress) {
    progress = false;
    for (int u = 0; u < (int)s.cost.size(); ++u) {
      if (!s.alive[u]) continue;
      const int deg = (int)s.adj[u].size();
      if (deg == 0) {
        Elim e{u, 0};
        int best = 0;
        for (int i = 1; i < (int)s.cost[u].size(); ++i)
          if (s.cost[u][i] < s.cost[u][best]) best = i;
        e.fixed = best;
        stack.push_back(e);
        s.constant = sat(s.constant, s.cost[u][best]);
        s.alive[u] = 0;
        progress = true;
      } else if (deg == 1) {
        const int a = s.adj[u].begin()->first;
        const PBQP::Mat& m = s.adj[u].begin()->second;  // [opt_u][opt_a]
        Elim e{u, 1, a};
        e.best1.resize(s.cost[a].size());
        for (int j = 0; j < (int)s.cost[a].size(); ++j) {
          double bv = kInf * 2;
          int bi = 0;
          for (int i = 0; i < (int)s.cost[u].size(); ++i) {
            double v = sat(s.cost[u][i], m[i][j]);
            if (v < bv) { bv = v; bi = i; }
          }
          e.best1[j] = bi;
          s.cost[a][j] = sat(s.cost[a][j], std::min(bv, kInf));
        }
        stack.push_back(e);
        RemoveNode(s, u);
        progress = true;
      } else if (deg == 2) {
        auto it = s.adj[u].begin();
        const int a = it->first;
        const PBQP::Mat ma = it->second;
        ++it;
        const int b = it->first;
        const PBQP::Mat mb = it->second;
        Elim e{u, 2, a, b};
        const int na = (int)s.cost[a].size(), nb = (int)s.cost[b].size();
        e.best2.assign(na, std::vector<int>(nb, 0));
        PBQP::Mat nm(na, PBQP::Vec(nb, 0));
        for (int j = 0; j < na; ++j)
          for (int k = 0; k < nb; ++k) {
            double bv = kInf * 2;
            int bi = 0;
            for (int i = 0; i < (int)s.cost[u].size(); ++i) {
              double v = sat(sat(s.cost[u][i], ma[i][j]), mb[i][k]);
              if (v < bv) { bv = v; bi = i; }
            }
            e.best2[j][k] = bi;
            nm[j][k] = std::min(bv, kInf);
          }
        stack.push_back(e);
        RemoveNode(s, u);
        AddEdgeTo(s, a, b, nm);
        progress = true;
      }
    }
  }
}

double LowerBound(const State& s) {
  double lb = 0;
  for (int u = 0; u < (int)s.cost.size(); ++u) {
    if (!s.alive[u]) continue;
    lb += *std::min_element(s.cost[u].begin(), s.cost[u].end());
    for (auto& kv : s.adj[u]) {
      if (kv.first < u) continue;
      double m = kInf;
      for (auto& row : kv.second)
        for (double x : row) m = std::min(m, x);
      lb += m;
    }
    if (lb >= kInf) return kInf;
  }
  return lb;
}

void FixNode(State& s, int u, int opt) {
  s.constant = sat(s.constant, s.cost[u][opt]);
  for (auto& kv : s.adj[u]) {
    const int v = kv.first;
    for (int j = 0; j < (int)s.cost[v].size(); ++j) s.cost[v][j] = sat(s.cost[v][j], kv.second[opt][j]);
  }
  RemoveNode(s, u);
}

}  // namespace

int PBQP::AddNode(const Vec& costs) {
  cost_.push_back(costs);
  adj_.emplace_back();
  return (int)cost_.size() - 1;
}

void PBQP::AddEdge(int u, int v, const Mat& m) {
  if (u == v) {  // self edge: fold the diagonal into the node cost
    for (size_t i = 0; i < cost_[u].size(); ++i) cost_[u][i] = sat(cost_[u][i], m[i][i]);
    return;
  }
  State s{cost_, adj_, {}};
  AddEdgeTo(s, u, v, m);
  adj_ = std::move(s.adj);
}

double PBQP::Evaluate(const std::vector<int>& choice) const {
  double c = 0;
  for (int u = 0; u < (int)cost_.size(); ++u) {
    c = sat(c, cost_[u][choice[u]]);
    for (auto& kv : adj_[u])
      if (kv.first > u) c = sat(c, kv.second[choice[u]][choice[kv.first]]);
  }
  return c;
}

PBQP::Result PBQP::Solve(double time_limit_s) {
  Result res;
  const int n = num_nodes();
  res.choice.assign(n, 0);
  if (n == 0) return res;
  State root{cost_, adj_, std::vector<char>(n, 1), 0.0};
  {
    State probe = root;
    std::vector<Elim> st;
    Reduce(probe, st);
    res.reduced_nodes = (int)st.size();
    for (int u = 0; u < n; ++u) res.core_nodes += probe.alive[u] ? 1 : 0;
  }
  auto t0 = std::chrono::steady_clock::now();
  auto timed_out = [&] {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > time_limit_s;
  };
  bool complete = true;

  // Exact search: reduce (DP), split into connected components (independent sub-problems: this is what keeps a deep
  // stack of identical layers linear instead of exponential), branch on the highest-degree node of a component.
  // Returns the optimal cost of `s` if it is < ub (and fills `stack`), else >= ub.
  std::function<double(State&, std::vector<Elim>&, double)> solve = [&](State& s, std::vector<Elim>& stack, double ub) -> double {
    ++res.bb_nodes;
    Reduce(s, stack);
    if (s.constant >= ub) return kInf * 8;
    std::vector<int> alive;
    for (int u = 0; u < n; ++u)
      if (s.alive[u]) alive.push_back(u);
    if (alive.empty()) return s.constant;
    if (s.constant + LowerBound(s) >= ub) return kInf * 8;
    // connected components
    std::vector<int> comp(n, -1);
    int nc = 0;
    for (int u : alive) {
      if (comp[u] >= 0) continue;
      std::vector<int> q{u};
      comp[u] = nc;
      while (!q.empty()) {
        int x = q.back();
        q.pop_back();
        for (auto& kv : s.adj[x])
          if (comp[kv.first] < 0) { comp[kv.first] = nc; q.push_back(kv.first); }
      }
      ++nc;
    }
    if (nc > 1) {
      double total = s.constant;
      std::vector<double> lbs(nc, 0.0);
      std::vector<State> subs(nc);
      for (int c = 0; c < nc; ++c) {
        subs[c] = s;
        subs[c].constant = 0;
        for (int u : alive)
          if (comp[u] != c) subs[c].alive[u] = 0;
        lbs[c] = LowerBound(subs[c]);
      }
      double rest = 0;
      for (int c = 0; c < nc; ++c) rest += lbs[c];
      for (int c = 0; c < nc; ++c) {
        rest -= lbs[c];
        std::vector<Elim> st;
        double r = solve(subs[c], st, ub - total - rest);
        if (r >= kInf) return kInf * 8;
        total += r;
        stack.insert(stack.end(), st.begin(), st.end());
        if (total + rest >= ub) return kInf * 8;
      }
      return total;
    }
    int pick = alive[0];
    for (int u : alive)
      if (s.adj[u].size() > s.adj[pick].size()) pick = u;
    std::vector<std::pair<double, int>> order;
    for (int i = 0; i < (int)s.cost[pick].size(); ++i) {
      double est = s.cost[pick][i];
      for (auto& kv : s.adj[pick]) {
        double m = kInf;
        for (int j = 0; j < (int)s.cost[kv.first].size(); ++j) m = std::min(m, sat(kv.second[i][j], s.cost[kv.first][j]));
        est = sat(est, m);
      }
      order.push_back({est, i});
    }
    std::sort(order.begin(), order.end());
    double best = ub;
    std::vector<Elim> best_stack;
    bool found = false, first = true;
    for (auto& oi : order) {
      if (oi.first >= kInf) continue;
      if (!first && timed_out()) { complete = false; break; }  // past the time limit only the greedy branch is taken
      first = false;
      State c = s;
      std::vector<Elim> st;
      Elim e{pick, 3};
      e.fixed = oi.second;
      st.push_back(e);
      FixNode(c, pick, oi.second);
      double r = solve(c, st, best);
      if (r < best) { best = r; best_stack = std::move(st); found = true; }
    }
    if (!found) return kInf * 8;
    stack.insert(stack.end(), best_stack.begin(), best_stack.end());
    return best;
  };

  std::vector<Elim> stack;
  double total = solve(root, stack, kInf * 4);
  res.optimal = complete;
  std::vector<int> ch(n, 0);
  if (total < kInf * 4) {
    for (int i = (int)stack.size() - 1; i >= 0; --i) {
      const Elim& e = stack[i];
      if (e.kind == 0 || e.kind == 3) ch[e.node] = e.fixed;
      else if (e.kind == 1) ch[e.node] = e.best1[ch[e.a]];
      else ch[e.node] = e.best2[ch[e.a]][ch[e.b]];
    }
  }
  res.choice = ch;
  res.cost = Evaluate(ch);
  return res;
}

}  // namespace tepdist
