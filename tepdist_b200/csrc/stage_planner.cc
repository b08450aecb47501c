#include "stage_planner.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <map>
#include <set>
#include <sstream>

#include "ilp.h"
#include "spmd_planner.h"

namespace tepdist {
namespace {
bool IsFwd(const Node& n) { return !n.backward && !IsSource(n.op) && n.op.rfind("apply_", 0) != 0; }
bool IsHeavy(const Node& n) { return IsComputeIntensive(n.op) || n.op == "attention"; }
}  // namespace

bool GraphSketch::IsChain() const {
  for (auto& e : edges)
    if (e.dst != e.src + 1) return false;
  return true;
}
double GraphSketch::TotalFlops() const {
  double f = 0;
  for (auto& n : nodes) f += n.fwd_flops + n.bwd_flops;
  return f;
}
std::string GraphSketch::ToDot() const {
  std::ostringstream o;
  o << "digraph sketch {\n";
  for (auto& n : nodes)
    o << "  s" << n.id << " [label=\"" << n.name << "\\n" << (n.fwd_flops + n.bwd_flops) / 1e9 << " GF\"];\n";
  for (auto& e : edges) o << "  s" << e.src << " -> s" << e.dst << " [label=\"" << e.bytes / 1e6 << " MB\"];\n";
  o << "}\n";
  return o.str();
}

GraphSketch BuildSketch(const Graph& g, bool fine_grained) {
  GraphSketch sk;
  const int N = (int)g.nodes.size();
  sk.node_of.assign(N, -1);
  // 1. representative of every forward node: heavy ops are cores; light ops are absorbed into the core that
  //    produced their first forward operand (or, for leading light ops, into the next core)
  std::vector<int> rep(N, -1);
  int last_core = -1;
  std::vector<int> pending;  // light nodes seen before any core
  for (auto& n : g.nodes) {
    if (!IsFwd(n)) continue;
    if (IsHeavy(n)) {
      rep[n.id] = n.id;
      for (int p : pending) rep[p] = n.id;
      pending.clear();
      last_core = n.id;
      continue;
    }
    int r = -1;
    for (auto& v : n.inputs)
      if (rep[v.node] >= 0) { r = rep[v.node]; }
    // light ops fed only by sources (re-layouts of variables / inputs) belong with their consumer: the next core
    if (r < 0) pending.push_back(n.id);
    else rep[n.id] = r;
  }
  for (int p : pending) rep[p] = last_core;
  // 2. coarse sketch: merge every core between two consecutive critical nodes
  if (!fine_grained) {
    std::vector<int> seps = FindCriticalNodes(g);
    auto seg = [&](int id) { return (int)(std::lower_bound(seps.begin(), seps.end(), id) - seps.begin()); };
    std::map<int, int> first_of_seg;
    for (auto& n : g.nodes)
      if (rep[n.id] >= 0) {
        int s = seg(n.id);
        if (!first_of_seg.count(s)) first_of_seg[s] = rep[n.id];
      }
    for (auto& n : g.nodes)
      if (rep[n.id] >= 0) rep[n.id] = first_of_seg[seg(n.id)];
  }
  // 3. materialise sketch nodes in topological (id) order
  std::map<int, int> idx;
  for (auto& n : g.nodes)
    if (rep[n.id] >= 0 && !idx.count(rep[n.id])) {
      int k = (int)sk.nodes.size();
      idx[rep[n.id]] = k;
      SketchNode s;
      s.id = k;
      s.name = g.nodes[rep[n.id]].name;
      sk.nodes.push_back(s);
    }
  std::map<int, int> group_sk;
  for (auto& n : g.nodes)
    if (rep[n.id] >= 0) {
      int k = idx[rep[n.id]];
      sk.node_of[n.id] = k;
      sk.nodes[k].members.push_back(n.id);
      sk.nodes[k].fwd_flops += NodeFlops(g, n);
      if (n.group >= 0) group_sk[n.group] = k;
    }
  for (auto& n : g.nodes) {
    if (n.backward && n.op.rfind("apply_", 0) != 0) {
      auto it = group_sk.find(n.group);
      if (it != group_sk.end()) sk.nodes[it->second].bwd_flops += NodeFlops(g, n);
    }
    if (n.op == "parameter") {
      for (auto& u : g.users(ValueRef{n.id, 0}))
        if (sk.node_of[u.node] >= 0) { sk.nodes[sk.node_of[u.node]].param_bytes += (double)n.outputs[0].numel() * 18.0; break; }
    }
  }
  std::map<std::pair<int, int>, double> em;
  std::set<std::pair<ValueRef, int>> counted;
  for (auto& n : g.nodes) {
    if (sk.node_of[n.id] < 0) continue;
    for (auto& v : n.inputs) {
      int a = sk.node_of[v.node], b = sk.node_of[n.id];
      if (a < 0 || a == b) continue;
      if (!counted.insert({v, b}).second) continue;  // one transfer per (value, consumer sketch node)
      em[{a, b}] += 2.0 * (double)g.type(v).bytes();
    }
  }
  for (auto& kv : em) sk.edges.push_back({kv.first.first, kv.first.second, kv.second});
  return sk;
}

StagePlanResult PlanStagesOnSketch(const GraphSketch& sk, const StagePlanOptions& opt) {
  auto t0 = std::chrono::steady_clock::now();
  StagePlanResult r;
  const int n = (int)sk.nodes.size(), S = opt.num_stages;
  r.sketch_stage.assign(n, 0);
  r.stage_flops.assign(S, 0.0);
  if (S <= 1 || n == 0) {
    for (auto& x : sk.nodes) r.stage_flops[0] += x.fwd_flops + x.bwd_flops;
    r.method = "single";
    return r;
  }
  const double total = sk.TotalFlops();
  std::vector<double> f(n);
  for (int i = 0; i < n; ++i) f[i] = sk.nodes[i].fwd_flops + sk.nodes[i].bwd_flops;
  double cap = total / S * (1.0 + opt.unbalanced_ratio);
  const double fmax = *std::max_element(f.begin(), f.end());
  cap = std::max(cap, fmax);  // a single indivisible node may exceed the nominal budget

  // ---- exact DP over contiguous partitions of the topological order (optimal for chains; a strong incumbent
  //      for general DAGs).  cutb[j] = bytes of every edge that crosses a cut placed after node j.
  std::vector<double> pre(n + 1, 0.0);
  for (int i = 0; i < n; ++i) pre[i + 1] = pre[i] + f[i];
  std::vector<double> cutb(n, 0.0);
  for (auto& e : sk.edges)
    for (int j = std::min(e.src, e.dst); j < std::max(e.src, e.dst); ++j) cutb[j] += e.bytes;
  bool dp_ok = false;
  double dp_cap = cap;
  if (n >= S) {
    for (int attempt = 0; attempt < 10 && !dp_ok; ++attempt) {
      const double INF = 1e300;
      std::vector<std::vector<double>> dp(S + 1, std::vector<double>(n + 1, INF));
      std::vector<std::vector<int>> from(S + 1, std::vector<int>(n + 1, -1));
      dp[0][0] = 0;
      for (int s = 1; s <= S; ++s)
        for (int j = s; j <= n; ++j)
          for (int i = s - 1; i < j; ++i) {
            if (dp[s - 1][i] >= INF || pre[j] - pre[i] > dp_cap * (1 + 1e-9)) continue;
            double c = dp[s - 1][i] + (j < n ? cutb[j - 1] : 0.0);
            c += 1e-9 * std::fabs((pre[j] - pre[i]) - total / S);  // tie-break towards balance
            if (c < dp[s][j]) { dp[s][j] = c; from[s][j] = i; }
          }
      if (dp[S][n] < INF) {
        int j = n;
        for (int s = S; s >= 1; --s) {
          int i = from[s][j];
          for (int k = i; k < j; ++k) r.sketch_stage[k] = s - 1;
          j = i;
        }
        dp_ok = true;
      } else {
        dp_cap *= 1.15;  // budget infeasible at this granularity: relax
      }
    }
  }
  r.method = sk.IsChain() ? "dp-chain" : "dp-topo";
  double dp_cost = 0;
  for (auto& e : sk.edges) dp_cost += e.bytes * std::abs(r.sketch_stage[e.dst] - r.sketch_stage[e.src]);

  const bool try_ilp = opt.force_ilp || (!sk.IsChain() && n * (S - 1) <= 90);
  if (try_ilp) {
    // ILP with cumulative binaries z[i][k] = 1 iff stage(i) <= k, k = 0..S-2 (reference BuildIlpStageModel; the
    // objective bytes * (stage(dst) - stage(src)) is linear because transfers are neighbour-only after B4 threading)
    IlpModel m;
    auto Z = [&](int i, int k) { return i * (S - 1) + k; };
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < S - 1; ++k) m.AddVar(0, 1, 0.0, true);
    for (int i = 0; i < n; ++i)
      for (int k = 0; k + 1 < S - 1; ++k) m.AddRow({Z(i, k), Z(i, k + 1)}, {1, -1}, -IlpModel::kInf, 0);  // monotone
    for (auto& e : sk.edges)
      for (int k = 0; k < S - 1; ++k) {
        m.AddRow({Z(e.dst, k), Z(e.src, k)}, {1, -1}, -IlpModel::kInf, 0);  // stage(dst) >= stage(src)
        m.obj[Z(e.src, k)] += e.bytes;
        m.obj[Z(e.dst, k)] -= e.bytes;
      }
    const double icap = dp_ok ? dp_cap : cap * 1.5;
    for (int k = 0; k < S; ++k) {  // FLOPs budget of stage k
      std::vector<int> idx;
      std::vector<double> val;
      for (int i = 0; i < n; ++i) {
        if (k < S - 1) { idx.push_back(Z(i, k)); val.push_back(f[i]); }
        if (k > 0) { idx.push_back(Z(i, k - 1)); val.push_back(-f[i]); }
      }
      m.AddRow(idx, val, -IlpModel::kInf, k == S - 1 ? icap - total : icap);
    }
    IlpResult ir = SolveIlp(m, opt.ilp_time_limit_s, dp_ok ? dp_cost * (1 - 1e-9) - 1e-6 : IlpModel::kInf);
    if (ir.status == IlpResult::kOptimal || ir.status == IlpResult::kFeasible) {
      for (int i = 0; i < n; ++i) {
        int st = S - 1;
        for (int k = 0; k < S - 1; ++k)
          if (ir.x[Z(i, k)] > 0.5) { st = k; break; }
        r.sketch_stage[i] = st;
      }
      r.method = "ilp";
      r.optimal = ir.status == IlpResult::kOptimal;
    } else {
      // nothing strictly better than the DP partition exists within the budget (or the time limit hit)
      r.method += ir.seconds >= opt.ilp_time_limit_s ? "+ilp-timeout" : "+ilp-confirmed";
      r.optimal = ir.seconds < opt.ilp_time_limit_s;
    }
  }
  for (int i = 0; i < n; ++i) r.stage_flops[r.sketch_stage[i]] += f[i];
  for (auto& e : sk.edges) r.cut_bytes += e.bytes * std::abs(r.sketch_stage[e.dst] - r.sketch_stage[e.src]);
  r.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return r;
}

BackwardPlanResult PlanBackwardOnSketch(const GraphSketch& sk, const std::vector<int>& sf, const std::vector<double>& act,
                                        const StagePlanOptions& opt) {
  BackwardPlanResult r;
  const int n = (int)sk.nodes.size(), S = opt.num_stages;
  r.sketch_stage = sf;
  r.method = "mirror";
  auto finish = [&]() {
    r.stage_flops.assign(std::max(1, S), 0.0);
    r.objective = 0;
    r.moved = 0;
    for (int i = 0; i < n; ++i) {
      r.stage_flops[sf[i]] += sk.nodes[i].fwd_flops;
      r.stage_flops[r.sketch_stage[i]] += sk.nodes[i].bwd_flops;
      r.objective += (i < (int)act.size() ? act[i] : 0.0) * std::abs(r.sketch_stage[i] - sf[i]);
      r.moved += r.sketch_stage[i] != sf[i];
    }
    for (auto& e : sk.edges) r.objective += 0.5 * e.bytes * std::abs(r.sketch_stage[e.dst] - r.sketch_stage[e.src]);
    return r;
  };
  if (S <= 1 || n == 0 || n * (S - 1) > 120) return finish();   // (large sketches: mirror; the ILP is for the coarse sketch)
  double total = 0, fmax = 0;
  for (auto& x : sk.nodes) { total += x.fwd_flops + x.bwd_flops; fmax = std::max(fmax, x.fwd_flops + x.bwd_flops); }
  // budget: what the forward plan already needed (the mirror placement is feasible by construction), never below the nominal one
  std::vector<double> mirror(S, 0.0);
  for (int i = 0; i < n; ++i) mirror[sf[i]] += sk.nodes[i].fwd_flops + sk.nodes[i].bwd_flops;
  const double nominal = std::max(total / S * (1.0 + opt.unbalanced_ratio), fmax);
  const double worst_mirror = *std::max_element(mirror.begin(), mirror.end());
  const double cap = std::max(nominal, 0.0) < worst_mirror ? nominal : worst_mirror;   // try to do at least as well as the nominal budget
  IlpModel m;
  auto Z = [&](int i, int k) { return i * (S - 1) + k; };
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < S - 1; ++k) m.AddVar(0, 1, 0.0, true);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k + 1 < S - 1; ++k) m.AddRow({Z(i, k), Z(i, k + 1)}, {1, -1}, -IlpModel::kInf, 0);   // z monotone in k
  double constant = 0;
  for (auto& e : sk.edges)
    for (int k = 0; k < S - 1; ++k) {
      m.AddRow({Z(e.dst, k), Z(e.src, k)}, {1, -1}, -IlpModel::kInf, 0);   // sb(dst) >= sb(src)
      m.obj[Z(e.src, k)] += 0.5 * e.bytes;     // (sketch edge bytes count both directions; the backward half is priced here)
      m.obj[Z(e.dst, k)] -= 0.5 * e.bytes;
    }
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < S - 1; ++k) {
      const double a = i < (int)act.size() ? act[i] : 0.0;
      if (sf[i] <= k) { m.obj[Z(i, k)] -= a; constant += a; }   // zf = 1: |z - 1| = 1 - z
      else m.obj[Z(i, k)] += a;                                  // zf = 0: |z - 0| = z
    }
  std::vector<double> fwd_on(S, 0.0);
  for (int i = 0; i < n; ++i) fwd_on[sf[i]] += sk.nodes[i].fwd_flops;
  double bwd_total = 0;
  for (auto& x : sk.nodes) bwd_total += x.bwd_flops;
  for (int k = 0; k < S; ++k) {   // device k: fwd_on[k] + sum_i bwd(i) * [sb(i) == k] <= cap, [sb == k] = z[k] - z[k-1]
    std::vector<int> idx;
    std::vector<double> val;
    for (int i = 0; i < n; ++i) {
      if (k < S - 1) { idx.push_back(Z(i, k)); val.push_back(sk.nodes[i].bwd_flops); }
      if (k > 0) { idx.push_back(Z(i, k - 1)); val.push_back(-sk.nodes[i].bwd_flops); }
    }
    m.AddRow(idx, val, -IlpModel::kInf, (k == S - 1 ? cap - fwd_on[k] - bwd_total : cap - fwd_on[k]));
  }
  IlpResult ir = SolveIlp(m, opt.ilp_time_limit_s);
  if (ir.status == IlpResult::kOptimal || ir.status == IlpResult::kFeasible) {
    for (int i = 0; i < n; ++i) {
      int st = S - 1;
      for (int k = 0; k < S - 1; ++k)
        if (ir.x[Z(i, k)] > 0.5) { st = k; break; }
      r.sketch_stage[i] = st;
    }
    r.method = ir.status == IlpResult::kOptimal ? "ilp" : "ilp-timeout";
  } else if (cap < worst_mirror) {
    r.method = "mirror";    // the nominal budget is infeasible for the backward groups at this granularity: keep the mirror
  }
  (void)constant;
  return finish();
}

StagePlanResult PlanStages(Graph* gp, const StagePlanOptions& opt) {
  Graph& g = *gp;
  GraphSketch sk = BuildSketch(g, /*fine_grained=*/false);
  if ((int)sk.nodes.size() < opt.num_stages) sk = BuildSketch(g, true);
  StagePlanResult r = PlanStagesOnSketch(sk, opt);
  // forward ops
  std::map<int, int> group_stage;
  for (auto& n : g.nodes)
    if (sk.node_of[n.id] >= 0) {
      n.stage = r.sketch_stage[sk.node_of[n.id]];
      // the op that OWNS the group decides where its gradient ops run (re-layout collectives inherit their
      // producer's group id but may sit on the consumer's stage)
      if (n.group >= 0 && !IsCollective(n.op) && !group_stage.count(n.group)) group_stage[n.group] = n.stage;
    }
  // backward ops: their op group is placed as a unit by the backward ILP (mirror stage of the forward group -- logical
  // 2S-1-s, same physical device s -- unless moving the group pays for shipping its activation stash)
  {
    std::vector<double> act(sk.nodes.size(), 0.0);
    for (auto& n : g.nodes) {
      const int k = sk.node_of[n.id];
      if (k < 0) continue;
      for (int o = 0; o < (int)n.outputs.size(); ++o) {
        bool saved = false;
        for (auto& u : g.users(ValueRef{n.id, o})) saved |= g.nodes[u.node].backward;
        if (saved) act[k] += (double)n.outputs[o].bytes();
      }
    }
    BackwardPlanResult bp = PlanBackwardOnSketch(sk, r.sketch_stage, act, opt);
    r.backward_stage = bp.sketch_stage;
    r.backward_method = bp.method;
    r.backward_moved = bp.moved;
    if (bp.moved > 0) {
      r.stage_flops = bp.stage_flops;
      std::map<int, int> group_sk;
      for (auto& n : g.nodes)
        if (sk.node_of[n.id] >= 0 && n.group >= 0 && !IsCollective(n.op) && !group_sk.count(n.group)) group_sk[n.group] = sk.node_of[n.id];
      for (auto& kv : group_sk) group_stage[kv.first] = bp.sketch_stage[kv.second];
    }
  }
  for (auto& n : g.nodes)
    if (n.stage < 0 && n.backward) {
      auto it = group_stage.find(n.group);
      if (it != group_stage.end()) n.stage = it->second;
    }
  // everything else by fixpoint: compute nodes follow their staged operands (apply nodes, trailing collectives),
  // sources (variables, optimizer slots, inputs, constants) follow their first staged consumer
  for (int iter = 0; iter < 8; ++iter) {
    bool changed = false;
    for (auto& n : g.nodes) {
      if (n.stage >= 0 || n.inputs.empty()) continue;
      int s = -1;
      bool all = true;
      for (auto& v : n.inputs) {
        const Node& p = g.nodes[v.node];
        if (p.stage < 0) { if (!p.inputs.empty()) all = false; continue; }
        s = std::max(s, p.stage);
      }
      if (all && s >= 0) { n.stage = s; changed = true; }
    }
    for (auto it = g.nodes.rbegin(); it != g.nodes.rend(); ++it) {
      Node& n = *it;
      if (n.stage >= 0) continue;
      int s = -1;
      for (int o = 0; o < (int)n.outputs.size(); ++o)
        for (auto& u : g.users(ValueRef{n.id, o}))
          if (g.nodes[u.node].stage >= 0) s = s < 0 ? g.nodes[u.node].stage : std::min(s, g.nodes[u.node].stage);
      if (s >= 0) { n.stage = s; changed = true; }
    }
    if (!changed) break;
  }
  for (auto& n : g.nodes)
    if (n.stage < 0) n.stage = 0;
  for (auto& n : g.nodes)
    for (auto& d : n.dist) d.stage = n.stage;
  g.stage_split_ordinal = (int)g.split_nums.size();
  g.record_split(opt.num_stages, false);
  return r;
}

}  // namespace tepdist
