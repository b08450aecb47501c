#include "ilp.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <functional>

namespace tepdist {
namespace {

constexpr double kEps = 1e-9;

struct LpOut {
  int status = 2;  // 0 optimal, 1 unbounded, 2 infeasible
  std::vector<double> x;
  double obj = 0;
};

// Dense two-phase simplex on:  min c^T y,  rows (a, sense, rhs),  y >= 0.   sense: -1 (<=), 0 (=), +1 (>=)
struct Dense {
  int n;
  std::vector<std::vector<double>> a;
  std::vector<int> sense;
  std::vector<double> rhs;
  std::vector<double> c;
};

LpOut SolveDense(const Dense& p) {
  LpOut out;
  const int n = p.n, m = (int)p.a.size();
  // normalise rhs >= 0
  std::vector<std::vector<double>> A = p.a;
  std::vector<int> sense = p.sense;
  std::vector<double> b = p.rhs;
  for (int i = 0; i < m; ++i)
    if (b[i] < 0) {
      for (auto& v : A[i]) v = -v;
      b[i] = -b[i];
      sense[i] = -sense[i];
    }
  int n_slack = 0, n_art = 0;
  for (int i = 0; i < m; ++i) {
    if (sense[i] != 0) ++n_slack;
    if (sense[i] >= 0) ++n_art;
  }
  const int cols = n + n_slack + n_art;
  std::vector<std::vector<double>> T(m + 1, std::vector<double>(cols + 1, 0.0));
  std::vector<int> basis(m, -1);
  int sc = n, ac = n + n_slack;
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < n; ++j) T[i][j] = A[i][j];
    T[i][cols] = b[i];
    if (sense[i] < 0) { T[i][sc] = 1.0; basis[i] = sc++; }
    else if (sense[i] > 0) { T[i][sc++] = -1.0; T[i][ac] = 1.0; basis[i] = ac++; }
    else { T[i][ac] = 1.0; basis[i] = ac++; }
  }
  auto pivot = [&](int r, int col) {
    const double pv = T[r][col];
    for (auto& v : T[r]) v /= pv;
    for (int i = 0; i <= m; ++i) {
      if (i == r) continue;
      const double f = T[i][col];
      if (std::fabs(f) < 1e-13) continue;
      for (int j = 0; j <= cols; ++j) T[i][j] -= f * T[r][j];
    }
    basis[r] = col;
  };
  auto run = [&](int ncols_active) -> int {  // minimise objective row T[m]; returns 0 ok, 1 unbounded
    long it = 0;
    while (true) {
      const bool bland = it > 20000;
      int col = -1;
      double best = -1e-9;
      for (int j = 0; j < ncols_active; ++j) {
        if (T[m][j] < best) {
          best = T[m][j];
          col = j;
          if (bland) break;
        }
      }
      if (col < 0) return 0;
      int row = -1;
      double ratio = 0;
      for (int i = 0; i < m; ++i)
        if (T[i][col] > 1e-9) {
          const double r = T[i][cols] / T[i][col];
          if (row < 0 || r < ratio - 1e-12 || (std::fabs(r - ratio) <= 1e-12 && basis[i] < basis[row])) { row = i; ratio = r; }
        }
      if (row < 0) return 1;
      pivot(row, col);
      if (++it > 200000) return 0;
    }
  };
  if (n_art > 0) {
    // phase 1: minimise the sum of artificials
    for (int j = 0; j <= cols; ++j) T[m][j] = 0;
    for (int j = n + n_slack; j < cols; ++j) T[m][j] = 1.0;
    for (int i = 0; i < m; ++i)
      if (basis[i] >= n + n_slack)
        for (int j = 0; j <= cols; ++j) T[m][j] -= T[i][j];
    run(cols);
    if (-T[m][cols] > 1e-7) { out.status = 2; return out; }
    // drive remaining artificials out of the basis
    for (int i = 0; i < m; ++i)
      if (basis[i] >= n + n_slack) {
        int col = -1;
        for (int j = 0; j < n + n_slack; ++j)
          if (std::fabs(T[i][j]) > 1e-9) { col = j; break; }
        if (col >= 0) pivot(i, col);
      }
  }
  // phase 2
  for (int j = 0; j <= cols; ++j) T[m][j] = 0;
  for (int j = 0; j < n; ++j) T[m][j] = p.c[j];
  for (int i = 0; i < m; ++i)
    if (basis[i] < n && std::fabs(p.c[basis[i]]) > 0) {
      const double f = T[m][basis[i]];
      for (int j = 0; j <= cols; ++j) T[m][j] -= f * T[i][j];
    }
  if (run(n + n_slack) == 1) { out.status = 1; return out; }
  out.status = 0;
  out.x.assign(n, 0.0);
  for (int i = 0; i < m; ++i)
    if (basis[i] < n) out.x[basis[i]] = T[i][cols];
  out.obj = 0;
  for (int j = 0; j < n; ++j) out.obj += p.c[j] * out.x[j];
  return out;
}

LpOut SolveRelaxation(const IlpModel& m, const std::vector<double>& lo, const std::vector<double>& hi) {
  // y = x - lo >= 0 ; upper bounds become rows
  Dense d;
  d.n = m.num_vars;
  d.c = m.obj;
  double shift = 0;
  for (int j = 0; j < m.num_vars; ++j) {
    if (lo[j] > hi[j] + 1e-9) return LpOut();
    shift += m.obj[j] * lo[j];
    if (hi[j] < IlpModel::kInf / 2) {
      std::vector<double> r(m.num_vars, 0.0);
      r[j] = 1.0;
      d.a.push_back(r);
      d.sense.push_back(-1);
      d.rhs.push_back(hi[j] - lo[j]);
    }
  }
  for (auto& row : m.rows) {
    std::vector<double> r(m.num_vars, 0.0);
    double off = 0;
    for (size_t k = 0; k < row.idx.size(); ++k) {
      r[row.idx[k]] += row.val[k];
      off += row.val[k] * lo[row.idx[k]];
    }
    const bool has_lo = row.lo > -IlpModel::kInf / 2, has_hi = row.hi < IlpModel::kInf / 2;
    if (has_lo && has_hi && std::fabs(row.lo - row.hi) < 1e-12) {
      d.a.push_back(r); d.sense.push_back(0); d.rhs.push_back(row.lo - off);
    } else {
      if (has_hi) { d.a.push_back(r); d.sense.push_back(-1); d.rhs.push_back(row.hi - off); }
      if (has_lo) { d.a.push_back(r); d.sense.push_back(+1); d.rhs.push_back(row.lo - off); }
    }
  }
  LpOut o = SolveDense(d);
  if (o.status == 0) {
    for (int j = 0; j < m.num_vars; ++j) o.x[j] += lo[j];
    o.obj += shift;
  }
  return o;
}

}  // namespace

std::string IlpResult::StatusName() const {
  switch (status) {
    case kOptimal: return "optimal";
    case kFeasible: return "feasible(time-limit)";
    case kInfeasible: return "infeasible";
    default: return "unbounded";
  }
}

IlpResult SolveLp(const IlpModel& m) {
  IlpResult r;
  LpOut o = SolveRelaxation(m, m.lo, m.hi);
  r.status = o.status == 0 ? IlpResult::kOptimal : (o.status == 1 ? IlpResult::kUnbounded : IlpResult::kInfeasible);
  r.x = o.x;
  r.objective = o.obj;
  return r;
}

IlpResult SolveIlp(const IlpModel& m, double time_limit_s, double cutoff) {
  IlpResult res;
  auto t0 = std::chrono::steady_clock::now();
  auto elapsed = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
  double incumbent = cutoff;
  std::vector<double> best_x;
  bool timed_out = false;

  std::function<void(std::vector<double>&, std::vector<double>&)> bb = [&](std::vector<double>& lo, std::vector<double>& hi) {
    if (timed_out) return;
    if (elapsed() > time_limit_s) { timed_out = true; return; }
    ++res.nodes;
    LpOut o = SolveRelaxation(m, lo, hi);
    if (o.status == 1 && best_x.empty()) { res.status = IlpResult::kUnbounded; return; }
    if (o.status != 0) return;
    if (o.obj >= incumbent - 1e-9 * std::max(1.0, std::fabs(incumbent))) return;
    int br = -1;
    double frac = 1e-6;
    for (int j = 0; j < m.num_vars; ++j)
      if (m.is_int[j]) {
        const double f = std::fabs(o.x[j] - std::round(o.x[j]));
        if (f > frac) { frac = f; br = j; }
      }
    if (br < 0) {
      incumbent = o.obj;
      best_x = o.x;
      for (int j = 0; j < m.num_vars; ++j)
        if (m.is_int[j]) best_x[j] = std::round(best_x[j]);
      return;
    }
    const double v = o.x[br];
    const double fl = std::floor(v), ce = std::ceil(v);
    const bool down_first = (v - fl) <= (ce - v);
    for (int side = 0; side < 2; ++side) {
      const bool down = (side == 0) == down_first;
      const double old_lo = lo[br], old_hi = hi[br];
      if (down) hi[br] = fl; else lo[br] = ce;
      bb(lo, hi);
      lo[br] = old_lo;
      hi[br] = old_hi;
    }
  };
  std::vector<double> lo = m.lo, hi = m.hi;
  bb(lo, hi);
  res.seconds = elapsed();
  if (!best_x.empty()) {
    res.status = timed_out ? IlpResult::kFeasible : IlpResult::kOptimal;
    res.x = best_x;
    res.objective = incumbent;
  } else if (res.status != IlpResult::kUnbounded) {
    res.status = IlpResult::kInfeasible;
  }
  return res;
}

}  // namespace tepdist
