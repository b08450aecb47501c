#include "spmd_planner.h"

#include <atomic>
#include <thread>
#include <tuple>

#include <algorithm>
#include <chrono>
#include <functional>
#include <set>
#include <sstream>
#include <unordered_map>

#include "pbqp.h"

namespace tepdist {
namespace {

struct Edge {
  int prod, out_idx, cons, operand;  // operand = -1: "update" edge (value -> variable storage)
  double bytes;
  int fanout = 1;                    // consumer edges of the same produced value
};

bool IsFwdCompute(const Node& n) {
  return !n.backward && !IsSource(n.op) && n.op.rfind("apply_", 0) != 0;
}

double VarStateBytes(const Graph& g, const Node& n, bool adam) {
  // fp32 master + compute copy + fp32 grad (+ two fp32 Adam moments)
  double e = (double)n.outputs[0].numel();
  return e * (4.0 + (n.outputs[0].dtype == "f32" ? 0.0 : 2.0) + 4.0 + (adam ? 8.0 : 0.0));
}

struct Problem {
  const Graph& g;
  const SpmdOptions& opt;
  std::vector<std::vector<Candidate>> cands;
  std::vector<std::vector<double>> node_cost;
  std::vector<Edge> edges;
  std::vector<std::vector<int>> edges_of;  // node -> edge indices
  std::set<int> forced;                    // nodes whose candidates were narrowed by the memory plan or a user annotation
};

void BuildProblem(Problem& p, SpmdStats* stats) {
  const Graph& g = p.g;
  const SpmdOptions& opt = p.opt;
  const int N = (int)g.nodes.size();
  p.cands.resize(N);
  p.node_cost.resize(N);
  p.edges_of.assign(N, {});
  const bool adam = std::any_of(g.nodes.begin(), g.nodes.end(), [](const Node& n) { return n.op == "apply_adamw" || n.op == "apply_lamb"; });

  // ---- memory plan (reference SplitPlanByMemCost): which variables MUST be stored sharded
  double var_bytes = 0;
  std::vector<std::pair<double, int>> vars;
  for (auto& n : g.nodes)
    if (n.op == "parameter") {
      double b = VarStateBytes(g, n, adam);
      var_bytes += b;
      if ((int)n.outputs[0].dims.size() >= opt.mem_split_min_rank) vars.push_back({b, n.id});
    }
  std::set<int> must_split;
  std::sort(vars.rbegin(), vars.rend());
  double est = var_bytes;
  for (auto& v : vars) {
    if (est <= opt.var_mem_limit) break;
    must_split.insert(v.second);
    est -= v.first * (1.0 - 1.0 / opt.num);
  }
  if (stats) {
    stats->forced_weight_splits = (int)must_split.size();
  }

  // reference split_for_mem_save_: a dot whose weight must be stored sharded loses its batch-split proposal, and
  // so do the gradient dots derived from it (same op_group) -> the weight is consumed sharded (tensor parallel)
  std::set<int> mem_save_groups;
  for (auto& n : g.nodes)
    if (IsComputeIntensive(n.op))
      for (auto& v : n.inputs)
        if (must_split.count(v.node)) mem_save_groups.insert(n.group);
  std::set<ValueRef> fetch(g.outputs.begin(), g.outputs.end());
  for (auto& n : g.nodes) {
    RuleOptions ro;
    ro.save_variable_mem = IsComputeIntensive(n.op) && mem_save_groups.count(n.group) > 0;
    ro.context_parallel = opt.context_parallel;
    ro.sequence_parallel = opt.sequence_parallel;
    if (ro.save_variable_mem) p.forced.insert(n.id);   // (candidates restricted to the ones that keep the weight split)
    auto c = EnumerateCandidates(g, n, opt.num, ro);
    // user annotations (xla_sharding.split / replicate equivalents)
    if (!opt.ignore_annotation && n.has("shard_dim")) {
      const int d = (int)n.attr_i("shard_dim");
      DimStrategy want = d < 0 ? DimStrategy::Glue() : DimStrategy::Split(d, opt.num);
      std::vector<Candidate> f;
      for (auto& x : c)
        if (!x.outs.empty() && x.outs[0] == want) f.push_back(x);
      if (!f.empty()) c = f, p.forced.insert(n.id);   // (a user annotation outranks a mirror pin as well)
      else if (stats) ++stats->ignored_annotations;     // e.g. the annotated dim is not divisible by the device count
    }
    // (slots with a reduced shape -- Adafactor row / column statistics, SM3 per-dim accumulators -- follow the apply node's
    // candidate instead: whether they can be split depends on WHICH dim of the variable is split)
    const bool full_slot = n.op == "state" && n.attr_i("slot_of", -1) >= 0 &&
                           n.outputs[0].dims == g.nodes[(int)n.attr_i("slot_of", -1)].outputs[0].dims;
    if (must_split.count(n.id) || (full_slot && must_split.count((int)n.attr_i("slot_of", -1)))) {
      std::vector<Candidate> f;
      for (auto& x : c)
        if (!x.outs[0].is_glue()) f.push_back(x);
      if (!f.empty()) c = f, p.forced.insert(n.id);
    }
    p.cands[n.id] = c;
    auto& nc = p.node_cost[n.id];
    nc.resize(c.size());
    for (size_t i = 0; i < c.size(); ++i) {
      double cost = c[i].node_cost;
      if (n.op == "linear" && c[i].tag.rfind("contract_rs", 0) == 0)   // the reduce-scatter inside the node is a launch like any other
        cost += opt.collective_latency_bytes >= 0 ? opt.collective_latency_bytes : opt.hw.coll_latency * opt.hw.link_bw;
      if (c[i].tag == "seq" && (n.op == "attention" || n.op == "attention_bwd")) {
        // the ring posts one exchange per hop (forward: K / V; backward: K / V and the dK / dV accumulator): same per-launch
        // price as every other collective, or the ring would look free next to the gathers it competes with
        const double lat = opt.collective_latency_bytes >= 0 ? opt.collective_latency_bytes : opt.hw.coll_latency * opt.hw.link_bw;
        cost += lat * (n.op == "attention" ? opt.num - 1 : 2 * opt.num - 1);
      }
      bool all_glue = true;
      for (auto& s : c[i].outs) all_glue &= s.is_glue();
      if (IsVariable(n.op)) {
        double b = n.op == "parameter" ? VarStateBytes(g, n, adam) : (double)n.outputs[0].bytes();
        cost += opt.memory_weight * (all_glue ? b : b / opt.num);
        if (n.op == "parameter" && !all_glue) cost += 32.0 + opt.shard_storage_penalty * b;  // (> the dynamic-slice epsilon)
      } else if (!IsSource(n.op) && all_glue) {
        double b = 0;
        for (auto& t : n.outputs) b += (double)t.bytes();
        cost += opt.replicate_penalty * b * (1.0 - 1.0 / opt.num);
        // attention is the one op here whose work grows with the SQUARE of a dim while its bytes grow linearly: running it
        // replicated repeats (n-1)/n of those FLOPs on every rank -- priced as the bytes the links move in that time, so that
        // past a few hundred tokens the K / V ring ("seq") beats gathering the sequence
        if (n.op == "attention" || n.op == "attention_bwd")
          cost += NodeFlops(g, n) * (1.0 - 1.0 / opt.num) * opt.hw.link_bw / opt.hw.flops;
      }
      for (int o = 0; o < (int)n.outputs.size(); ++o)
        if (fetch.count(ValueRef{n.id, o}))
          cost += ReshardCost(c[i].outs[o], DimStrategy::Glue(), (double)n.outputs[o].bytes(), opt.num, opt.cost_factor);
      nc[i] = cost;
    }
  }
  for (auto& n : g.nodes)
    for (int k = 0; k < (int)n.inputs.size(); ++k) {
      ValueRef v = n.inputs[k];
      p.edges.push_back({v.node, v.idx, n.id, k, (double)g.type(v).bytes()});
    }
  for (auto& kv : g.updates)  // In/Out affinity: the updated value has to land in the variable's storage layout
    p.edges.push_back({kv.second.node, kv.second.idx, kv.first, -1, (double)g.type(kv.second).bytes()});
  if (opt.aux_affinity)
    for (auto& n : g.nodes)  // Var/Aux affinity: slots follow their variable (zero-byte edge with infinite mismatch)
      if (n.op == "state" && n.has("slot_of") && n.outputs[0].dims == g.nodes[(int)n.attr_i("slot_of")].outputs[0].dims)
        p.edges.push_back({(int)n.attr_i("slot_of"), 0, n.id, -2, 0.0});   // (reduced-shape slots: see rules.cc AdafactorRule)
  {
    std::map<std::pair<int, int>, int> fan;     // produced value -> number of operand edges reading it
    for (auto& e : p.edges)
      if (e.operand >= 0) ++fan[{e.prod, e.out_idx}];
    for (auto& e : p.edges)
      if (e.operand >= 0) e.fanout = fan[{e.prod, e.out_idx}];
  }
  for (int e = 0; e < (int)p.edges.size(); ++e) {
    p.edges_of[p.edges[e].prod].push_back(e);
    p.edges_of[p.edges[e].cons].push_back(e);
  }
  if (stats) stats->var_bytes_per_device = est;
}

double EdgeCost(const Problem& p, const Edge& e, const Candidate& cp, const Candidate& cc) {
  const DimStrategy& from = cp.outs[e.out_idx];
  if (e.operand == -2) return from == cc.outs[0] ? 0.0 : kInfCost;  // aux affinity
  const DimStrategy& to = e.operand >= 0 ? cc.ins[e.operand] : cc.outs[0];
  // an optimizer slot that is stored split is consumed by its update in exactly that layout: gathering it (or moving it to
  // another split) every step would also leave the update's result in a layout the storage does not have
  if (e.operand >= 2 && from.is_split() && !(from == to) && p.g.nodes[e.prod].op == "state" &&
      p.g.nodes[e.cons].op.rfind("apply_", 0) == 0)
    return kInfCost;
  double cost = ReshardCost(from, to, e.bytes, p.opt.num, p.opt.cost_factor);
  // Experimental (SpmdOptions::share_relayout_cost): the rewrite re-lays a value out once per target layout however many
  // consumers want it, the pairwise objective charges every consumer edge.  Splitting the price over the value's consumers is
  // exact when they all ask for the same layout (the common case: an all-reduced activation read by a LayerNorm and a residual
  // add) and an under-estimate when they ask for different ones.
  const double share = (p.opt.share_relayout_cost && e.operand >= 0 && e.fanout > 1) ? 1.0 / e.fanout : 1.0;
  if (cost > 0 && cost < kInfCost) cost *= share;
  if (cost > 0 && cost < kInfCost) {
    // a collective is not free below its byte count: launch + cross-GPU synchronisation (~8 us on NVSwitch = ~6 MB of wire
    // time).  Without this term dozens of tiny LayerNorm-statistics / loss all-reduces look free next to one large one.
    const Reshard kind = ClassifyReshard(from, to);
    const bool launches = kind == Reshard::kAllReduce || kind == Reshard::kAllGather || kind == Reshard::kReduceScatter ||
                          kind == Reshard::kAllToAll;
    const std::string& po = p.g.nodes[e.prod].op;
    const std::string& co = p.g.nodes[e.cons].op;
    const bool bucketed = po == "parameter" || po == "state" || po.rfind("apply_", 0) == 0 || co.rfind("apply_", 0) == 0;
    if (launches && !bucketed) {
      const double lat = p.opt.collective_latency_bytes >= 0 ? p.opt.collective_latency_bytes : p.opt.hw.coll_latency * p.opt.hw.link_bw;
      cost += lat * share;
    }
  }
  return cost;
}

// Solve the sub-problem over `members`; `fixed` pins nodes (members or foreign endpoints) to one strategy of
// their output 0 (foreign nodes are represented by a single synthetic candidate).
struct SubSolution {
  std::vector<int> choice;  // per member (same order)
  double cost = kInfCost;
  bool optimal = true;
  int core = 0;
};

SubSolution SolveSub(const Problem& p, const std::vector<int>& members, const std::map<int, DimStrategy>& pin_member,
                     const std::map<int, DimStrategy>& foreign) {
  SubSolution sol;
  std::unordered_map<int, int> local;  // node -> pbqp id
  PBQP q;
  std::vector<std::vector<int>> opt_index(members.size());  // pbqp option -> candidate index
  for (size_t m = 0; m < members.size(); ++m) {
    const int nid = members[m];
    std::vector<double> c;
    auto pin = pin_member.find(nid);
    for (size_t i = 0; i < p.cands[nid].size(); ++i) {
      if (pin != pin_member.end() && p.cands[nid][i].outs[0] != pin->second) continue;
      opt_index[m].push_back((int)i);
      c.push_back(p.node_cost[nid][i]);
    }
    if (c.empty() && pin != pin_member.end() && p.forced.count(nid)) {
      // the pin cannot be honoured: a variable that MUST be stored split whose value leaves the sub-graph, for which the
      // mirror layout is Glue: leave the node free -- the consumer in the other sub-graph then pays the re-layout, which
      // the final accounting derives from the actual choices anyway.  (This used to make the whole sub-graph infeasible and
      // silently fall back to candidate 0 for all of its nodes.)
      for (size_t i = 0; i < p.cands[nid].size(); ++i) {
        opt_index[m].push_back((int)i);
        c.push_back(p.node_cost[nid][i]);
      }
    }
    if (c.empty()) return sol;
    local[nid] = q.AddNode(c);
  }
  std::set<int> seen_edges;
  for (size_t m = 0; m < members.size(); ++m) {
    const int nid = members[m];
    for (int ei : p.edges_of[nid]) {
      if (!seen_edges.insert(ei).second) continue;
      const Edge& e = p.edges[ei];
      const bool prod_in = local.count(e.prod) > 0, cons_in = local.count(e.cons) > 0;
      if (prod_in && cons_in) {
        const int a = local[e.prod], b = local[e.cons];
        const auto& oa = opt_index[a];
        const auto& ob = opt_index[b];
        PBQP::Mat mat(oa.size(), PBQP::Vec(ob.size()));
        for (size_t i = 0; i < oa.size(); ++i)
          for (size_t j = 0; j < ob.size(); ++j) mat[i][j] = EdgeCost(p, e, p.cands[e.prod][oa[i]], p.cands[e.cons][ob[j]]);
        if (a == b) {
          PBQP::Mat d = mat;
          q.AddEdge(a, a, d);
        } else {
          q.AddEdge(a, b, mat);
        }
      } else if (prod_in || cons_in) {
        // one endpoint lives in another sub-graph: it is pinned to the mirrored separator layout (or Glue)
        const int other = prod_in ? e.cons : e.prod;
        auto f = foreign.find(other);
        DimStrategy fs = f == foreign.end() ? DimStrategy::Glue() : f->second;
        Candidate fake;
        fake.outs.assign(std::max(1, e.out_idx + 1), fs);
        fake.ins.assign(std::max(1, e.operand + 1), fs);
        const int a = local[prod_in ? e.prod : e.cons];
        // we cannot add a unary term through AddEdge; fold into a 1-option node
        std::vector<double> add(opt_index[a].size());
        for (size_t i = 0; i < opt_index[a].size(); ++i) {
          const Candidate& mine = p.cands[prod_in ? e.prod : e.cons][opt_index[a][i]];
          add[i] = prod_in ? EdgeCost(p, e, mine, fake) : EdgeCost(p, e, fake, mine);
        }
        int fn = q.AddNode({0.0});
        PBQP::Mat mat(add.size(), PBQP::Vec(1));
        for (size_t i = 0; i < add.size(); ++i) mat[i][0] = add[i];
        q.AddEdge(a, fn, mat);
      }
    }
  }
  auto r = q.Solve(p.opt.ilp_time_limit_s);
  sol.cost = r.cost;
  sol.optimal = r.optimal;
  sol.core = r.core_nodes;
  sol.choice.resize(members.size());
  for (size_t m = 0; m < members.size(); ++m) sol.choice[m] = opt_index[m][r.choice[m]];
  return sol;
}

std::string SegmentSignature(const Problem& p, const std::vector<int>& members, int head, int tail) {
  std::unordered_map<int, int> pos;
  for (size_t i = 0; i < members.size(); ++i) pos[members[i]] = (int)i;
  std::ostringstream o;
  for (int nid : members) {
    const Node& n = p.g.nodes[nid];
    o << n.op << (n.backward ? "b" : "f") << (nid == tail ? "T" : "");
    for (auto& t : n.outputs) {
      o << t.dtype;
      for (auto d : t.dims) o << "," << d;
    }
    o << "(";
    for (auto& v : n.inputs) {
      auto it = pos.find(v.node);
      if (it != pos.end()) o << it->second << "." << v.idx;
      else if (v.node == head) o << "H";
      else o << "X" << p.g.type(v).numel();
      o << " ";
    }
    o << ")" << p.cands[nid].size() << ";";
  }
  return o.str();
}

}  // namespace

std::vector<int> FindCriticalNodes(const Graph& g, double min_segment_flops_frac) {
  std::vector<int> seps;
  std::map<ValueRef, int> live;  // value -> remaining forward uses
  auto fwd_uses = [&](ValueRef v) {
    int c = 0;
    for (auto& u : g.users(v))
      if (IsFwdCompute(g.nodes[u.node])) ++c;
    return c;
  };
  int last_fwd = -1;
  for (auto& n : g.nodes)
    if (IsFwdCompute(n)) last_fwd = n.id;
  for (auto& n : g.nodes) {
    if (!IsFwdCompute(n)) continue;
    for (auto& v : n.inputs) {
      auto it = live.find(v);
      if (it != live.end() && --it->second == 0) live.erase(it);
    }
    for (int o = 0; o < (int)n.outputs.size(); ++o) {
      int c = fwd_uses(ValueRef{n.id, o});
      if (c > 0) live[ValueRef{n.id, o}] = c;
    }
    if (n.id != last_fwd && live.size() == 1 && live.begin()->first.node == n.id && live.begin()->first.idx == 0 &&
        n.outputs[0].bytes() >= 1024)
      seps.push_back(n.id);
  }
  if (min_segment_flops_frac > 0 && !seps.empty()) {
    // tiny-node clustering: a separator whose sub-graph (the forward nodes since the previous kept separator) carries less than
    // the given share of the forward FLOPs is dropped, merging that sub-graph into the next one
    double total = 0;
    for (auto& n : g.nodes)
      if (IsFwdCompute(n)) total += NodeFlops(g, n);
    std::vector<int> kept;
    double acc = 0;
    size_t si = 0;
    for (auto& n : g.nodes) {
      if (!IsFwdCompute(n)) continue;
      acc += NodeFlops(g, n);
      if (si < seps.size() && n.id == seps[si]) {
        if (acc >= min_segment_flops_frac * total) { kept.push_back(n.id); acc = 0; }
        ++si;
      }
    }
    seps.swap(kept);
  }
  return seps;
}

std::vector<int> FindCriticalNodesByMainPath(const Graph& g) {
  // forward compute nodes in topological (id) order
  std::vector<int> order;
  for (auto& n : g.nodes)
    if (IsFwdCompute(n)) order.push_back(n.id);
  if (order.empty()) return {};
  std::map<int, int> pos;
  for (int i = 0; i < (int)order.size(); ++i) pos[order[i]] = i;
  const int L = (int)order.size();
  // max-FLOPs path ending at the last forward node: best[i] = heaviest path weight ending at i, pred[i]
  std::vector<double> best(L, 0.0);
  std::vector<int> pred(L, -1);
  for (int i = 0; i < L; ++i) {
    const Node& n = g.nodes[order[i]];
    double b = 0;
    int p = -1;
    for (auto& v : n.inputs) {
      auto it = pos.find(v.node);
      if (it != pos.end() && best[it->second] >= b) { b = best[it->second]; p = it->second; }
    }
    best[i] = b + NodeFlops(g, n) + 1e-3;   // (+epsilon: zero-FLOP nodes still extend a path)
    pred[i] = p;
  }
  std::vector<char> on_path(L, 0);
  for (int i = L - 1; i >= 0; i = pred[i]) {
    on_path[i] = 1;
    if (pred[i] < 0) break;
  }
  // bypass count per position: forward edges (u -> v) with pos(u) < c < pos(v); plus values of u still needed after c
  std::vector<int> bypass(L + 1, 0);
  for (int i = 0; i < L; ++i) {
    const Node& n = g.nodes[order[i]];
    for (auto& v : n.inputs) {
      auto it = pos.find(v.node);
      if (it == pos.end()) continue;
      const int a = it->second;
      if (i - a >= 2) { bypass[a + 1] += 1; bypass[i] -= 1; }   // positions a+1 .. i-1 are bypassed
    }
  }
  std::vector<int> out;
  int run = 0;
  for (int c = 0; c < L; ++c) {
    run += bypass[c];
    if (c == L - 1 || !on_path[c] || run != 0) continue;
    const Node& n = g.nodes[order[c]];
    // FreedomDegree == 0 also requires that only output 0 of the node is consumed downstream in the forward pass
    bool single = n.outputs[0].bytes() >= 1024;
    bool has_use = false;
    for (int o = 0; o < (int)n.outputs.size(); ++o)
      for (auto& u : g.users(ValueRef{n.id, o}))
        if (IsFwdCompute(g.nodes[u.node])) { has_use = true; if (o != 0) single = false; }
    if (single && has_use) out.push_back(n.id);
  }
  return out;
}

SpmdPlan PlanSpmdLevel(Graph* gp, const SpmdOptions& opt) {
  Graph& g = *gp;
  auto t0 = std::chrono::steady_clock::now();
  SpmdPlan plan;
  Problem p{g, opt, {}, {}, {}, {}};
  BuildProblem(p, &plan.stats);
  const int N = (int)g.nodes.size();
  std::vector<int> chosen(N, 0);

  std::vector<int> seps;
  if (opt.opt_level < 3) seps = FindCriticalNodes(g, opt.min_segment_flops_frac);
  if (opt.forward_sub_graph_num > 0 && (int)seps.size() > opt.forward_sub_graph_num - 1) {
    std::vector<int> pick;
    const int want = opt.forward_sub_graph_num - 1;
    for (int i = 1; i <= want; ++i) pick.push_back(seps[(size_t)i * seps.size() / (want + 1)]);
    seps = pick;
  }
  // keep only separators that give reasonably sized sub-graphs (merge tiny ones)
  const int S = (int)seps.size() + 1;

  // ---- segment membership
  std::vector<int> seg(N, -1);
  std::map<int, int> group_seg;
  auto seg_of_fwd = [&](int id) { return (int)(std::lower_bound(seps.begin(), seps.end(), id) - seps.begin()); };
  for (auto& n : g.nodes)
    if (IsFwdCompute(n)) {
      seg[n.id] = seg_of_fwd(n.id);
      if (n.group >= 0 && !group_seg.count(n.group)) group_seg[n.group] = seg[n.id];
    }
  for (auto& n : g.nodes)  // sources: first consumer's segment
    if (IsSource(n.op) && !(n.op == "state")) {
      int s = S;
      for (int o = 0; o < (int)n.outputs.size(); ++o)
        for (auto& u : g.users(ValueRef{n.id, o}))
          if (seg[u.node] >= 0) s = std::min(s, seg[u.node]);
      if (s == S) s = 0;
      seg[n.id] = s;
      if (n.group >= 0 && !group_seg.count(n.group)) group_seg[n.group] = s;
    }
  for (auto& n : g.nodes)
    if (seg[n.id] < 0) {
      auto it = group_seg.find(n.group);
      seg[n.id] = it != group_seg.end() ? it->second : S - 1;
    }
  std::vector<std::vector<int>> members(S);
  for (auto& n : g.nodes) members[seg[n.id]].push_back(n.id);

  // ---- separator strategy options
  auto sep_options = [&](int nid) {
    std::vector<DimStrategy> o;
    for (auto& c : p.cands[nid])
      if (std::find(o.begin(), o.end(), c.outs[0]) == o.end()) o.push_back(c.outs[0]);
    return o;
  };
  std::vector<std::vector<DimStrategy>> sopt(seps.size());
  for (size_t k = 0; k < seps.size(); ++k) sopt[k] = sep_options(seps[k]);

  // foreign endpoint layout: mirror of the separator between the two segments when shapes agree
  auto mirror = [](const DimStrategy* s) { return (s && s->partial) ? DimStrategy::Glue() : (s ? *s : DimStrategy::Glue()); };
  auto foreign_map = [&](int k, const DimStrategy* sh, const DimStrategy* st) {
    std::map<int, DimStrategy> f;
    for (int nid : members[k])
      for (int ei : p.edges_of[nid]) {
        const Edge& e = p.edges[ei];
        const int other = e.prod == nid ? e.cons : e.prod;
        if (seg[other] == k) continue;
        DimStrategy s = DimStrategy::Glue();
        ValueRef val{e.prod, e.out_idx};
        if (seg[other] < k && sh && k - 1 < (int)seps.size()) {  // (gradients of a separator may skip one sub-graph)
          if (other == seps[k - 1] && e.prod == other && e.out_idx == 0) s = *sh;
          else if (g.type(val).dims == g.nodes[seps[k - 1]].outputs[0].dims) s = mirror(sh);
        } else if (seg[other] > k && st && k < (int)seps.size()) {
          if (g.type(val).dims == g.nodes[seps[k]].outputs[0].dims) s = mirror(st);
        }
        f[other] = s;
      }
    return f;
  };
  // members whose outputs leave the segment (other than the tail separator) are pinned to the mirror layout too
  auto pin_map = [&](int k, const DimStrategy* sh, const DimStrategy* st) {
    std::map<int, DimStrategy> pin;
    if (st && k < (int)seps.size()) pin[seps[k]] = *st;
    for (int nid : members[k]) {
      if (k < (int)seps.size() && nid == seps[k]) continue;
      for (int ei : p.edges_of[nid]) {
        const Edge& e = p.edges[ei];
        if (e.prod != nid || seg[e.cons] == k || e.operand < 0) continue;
        DimStrategy s = DimStrategy::Glue();
        ValueRef val{e.prod, e.out_idx};
        if (seg[e.cons] < k && sh && g.type(val).dims == g.nodes[seps[k - 1]].outputs[0].dims) s = mirror(sh);
        else if (seg[e.cons] > k && st && g.type(val).dims == g.nodes[seps[k]].outputs[0].dims) s = mirror(st);
        if (e.out_idx == 0) pin[nid] = s;
      }
    }
    return pin;
  };

  // ---- DP across sub-graphs keyed on the separator layout
  std::map<std::string, SubSolution> memo;
  struct Cell { double cost = kInfCost; int prev = -1; SubSolution sol; };
  std::vector<std::vector<Cell>> dp(S);
  std::set<std::string> distinct;
  // ILP_NUM_THREADS > 1: the sub-problems (one per distinct sub-graph signature x head layout x tail layout) are independent --
  // solve them on a pool first, the DP below then only looks them up (reference: ILP_NUM_THREADS feeds its MIP solver's threads)
  if (opt.num_threads > 1) {
    struct Job { std::string key; int k; const DimStrategy* sh; const DimStrategy* st; SubSolution sol; };
    std::vector<Job> jobs;
    std::set<std::string> queued;
    for (int k = 0; k < S; ++k) {
      const int nh = k > 0 ? (int)sopt[k - 1].size() : 1;
      const int nt = k < S - 1 ? (int)sopt[k].size() : 1;
      const std::string sig = SegmentSignature(p, members[k], k > 0 ? seps[k - 1] : -1, k < S - 1 ? seps[k] : -1);
      for (int it = 0; it < nt; ++it)
        for (int ih = 0; ih < nh; ++ih) {
          const DimStrategy* sh = k > 0 ? &sopt[k - 1][ih] : nullptr;
          const DimStrategy* st = k < S - 1 ? &sopt[k][it] : nullptr;
          std::string key = sig + "|" + (sh ? sh->str() : "-") + "|" + (st ? st->str() : "-");
          if (queued.insert(key).second) jobs.push_back({std::move(key), k, sh, st, SubSolution()});
        }
    }
    std::atomic<size_t> next{0};
    auto worker = [&]() {
      for (size_t i = next++; i < jobs.size(); i = next++) {
        Job& j = jobs[i];
        j.sol = SolveSub(p, members[j.k], pin_map(j.k, j.sh, j.st), foreign_map(j.k, j.sh, j.st));
      }
    };
    const int nthr = (int)std::min<size_t>((size_t)opt.num_threads, std::max<size_t>(1, jobs.size()));
    std::vector<std::thread> pool;
    for (int t = 1; t < nthr; ++t) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
    for (auto& j : jobs) memo.emplace(j.key, std::move(j.sol));
    plan.stats.threads_used = nthr;
  }
  for (int k = 0; k < S; ++k) {
    const int nh = k > 0 ? (int)sopt[k - 1].size() : 1;
    const int nt = k < S - 1 ? (int)sopt[k].size() : 1;
    dp[k].assign(nt, Cell());
    const int head = k > 0 ? seps[k - 1] : -1, tail = k < S - 1 ? seps[k] : -1;
    const std::string sig = SegmentSignature(p, members[k], head, tail);
    distinct.insert(sig);
    for (int it = 0; it < nt; ++it)
      for (int ih = 0; ih < nh; ++ih) {
        if (k > 0 && dp[k - 1][ih].cost >= kInfCost) continue;
        const DimStrategy* sh = k > 0 ? &sopt[k - 1][ih] : nullptr;
        const DimStrategy* st = k < S - 1 ? &sopt[k][it] : nullptr;
        const std::string key = sig + "|" + (sh ? sh->str() : "-") + "|" + (st ? st->str() : "-");
        auto mit = memo.find(key);
        if (mit == memo.end()) {
          SubSolution s = SolveSub(p, members[k], pin_map(k, sh, st), foreign_map(k, sh, st));
          mit = memo.emplace(key, std::move(s)).first;
        }
        const SubSolution& s = mit->second;
        if (s.cost >= kInfCost) continue;
        const double total = (k > 0 ? dp[k - 1][ih].cost : 0.0) + s.cost;
        plan.stats.core_nodes_max = std::max(plan.stats.core_nodes_max, s.core);
        plan.stats.optimal = plan.stats.optimal && s.optimal;
        if (total < dp[k][it].cost) {
          dp[k][it].cost = total;
          dp[k][it].prev = ih;
          dp[k][it].sol = s;
        }
      }
  }
  // ---- back-track
  {
    int it = 0;
    for (int i = 1; i < (int)dp[S - 1].size(); ++i)
      if (dp[S - 1][i].cost < dp[S - 1][it].cost) it = i;
    for (int k = S - 1; k >= 0; --k) {
      const Cell& c = dp[k][it];
      if (c.sol.choice.size() == members[k].size()) {
        for (size_t m = 0; m < members[k].size(); ++m) chosen[members[k][m]] = c.sol.choice[m];
      } else {
        ++plan.stats.infeasible_subgraphs;   // its nodes keep their first candidate; never silent (see SpmdStats)
        plan.stats.optimal = false;
      }
      it = std::max(0, c.prev);
    }
  }

  // ---- record (reference RecordStrategyToInsts) + statistics
  plan.choice.resize(N);
  for (auto& n : g.nodes) {
    plan.choice[n.id] = p.cands[n.id][chosen[n.id]];
    for (int o = 0; o < (int)n.outputs.size(); ++o) n.dist[o].levels.push_back(plan.choice[n.id].outs[o]);
  }
  for (auto& n : g.nodes) {      // communication INSIDE nodes (K / V ring of context-parallel attention, reduce-scatter of "contract_rs")
    const Candidate& c = plan.choice[n.id];
    const bool ring = c.tag == "seq" && (n.op == "attention" || n.op == "attention_bwd");
    const bool rs = n.op == "linear" && c.tag.rfind("contract_rs", 0) == 0;
    if (ring || rs) plan.stats.comm_bytes += c.node_cost;
    if (rs) plan.stats.collectives["reduce_scatter"]++;
  }
  std::set<std::tuple<int, int, std::string>> counted;
  for (auto& e : p.edges) {
    if (e.operand == -2) continue;
    const DimStrategy& from = plan.choice[e.prod].outs[e.out_idx];
    const DimStrategy& to = e.operand >= 0 ? plan.choice[e.cons].ins[e.operand] : plan.choice[e.cons].outs[0];
    Reshard r = ClassifyReshard(from, to);
    if (r == Reshard::kNone) continue;
    // the rewrite re-lays a value out once per TARGET layout, however many consumers want it (transform.cc reshard cache): count
    // it once here too.  (The PBQP objective itself is pairwise -- it charges every consumer edge -- so the solver is biased
    // against re-laying out values with several consumers; the statistics are exact.)
    if (!counted.insert({e.prod, e.out_idx, to.str()}).second) continue;
    plan.stats.collectives[ReshardName(r)]++;
    if (r != Reshard::kInvalid && r != Reshard::kDynamicSlice)
      plan.stats.comm_bytes += ReshardBytes(r, e.bytes, opt.num, opt.cost_factor);
  }
  plan.stats.num_subgraphs = S;
  plan.stats.distinct_subgraphs = (int)distinct.size();
  plan.stats.solve_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return plan;
}

SpmdPlan PlanSpmdByRules(Graph* gp, const SpmdOptions& opt) {
  Graph& g = *gp;
  SpmdPlan plan;
  std::map<ValueRef, DimStrategy> assign;
  for (auto& n : g.nodes)
    if (n.has("shard_dim")) {
      const int d = (int)n.attr_i("shard_dim");
      assign[ValueRef{n.id, 0}] = d < 0 ? DimStrategy::Glue() : DimStrategy::Split(d, opt.num);
    }
  std::string conflict;
  InferGraph(g, opt.num, &assign, &conflict);
  RuleOptions ro;
  ro.allow_glue_compute_intensive = true;
  plan.choice.resize(g.nodes.size());
  for (auto& n : g.nodes) {
    auto cands = EnumerateCandidates(g, n, opt.num, ro);
    int best = -1, best_score = -1;
    for (int i = 0; i < (int)cands.size(); ++i) {
      int score = 0;
      bool ok = true;
      for (int o = 0; o < (int)n.outputs.size() && ok; ++o) {
        auto it = assign.find(ValueRef{n.id, o});
        if (it == assign.end()) { if (!cands[i].outs[o].is_glue()) ok = false; continue; }
        if (cands[i].outs[o] == it->second) score += 2; else ok = false;
      }
      for (int k = 0; k < (int)n.inputs.size() && ok; ++k) {
        auto it = assign.find(n.inputs[k]);
        if (it != assign.end() && cands[i].ins[k] == it->second) score += 1;
      }
      if (ok && score > best_score) { best_score = score; best = i; }
    }
    if (best < 0) best = (int)cands.size() - 1;  // all-glue candidate is always last
    plan.choice[n.id] = cands[best];
    for (int o = 0; o < (int)n.outputs.size(); ++o) n.dist[o].levels.push_back(cands[best].outs[o]);
  }
  return plan;
}

std::string DumpStrategies(const Graph& g, const SpmdPlan& plan) {
  std::ostringstream o;
  for (auto& n : g.nodes) {
    const Candidate& c = plan.choice[n.id];
    o << "%" << n.id << " " << n.op << " [" << c.tag << "] " << n.name << " : (";
    for (size_t i = 0; i < c.ins.size(); ++i) o << (i ? "," : "") << c.ins[i].str();
    o << ") -> (";
    for (size_t i = 0; i < c.outs.size(); ++i) o << (i ? "," : "") << c.outs[i].str();
    o << ")\n";
  }
  return o.str();
}

}  // namespace tepdist
