#include "runtime/task_graph.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <functional>
#include <queue>
#include <set>
#include <sstream>

namespace tepdist {

const char* TaskTypeName(TaskType t) {
  switch (t) {
    case TaskType::kSplit: return "Split";
    case TaskType::kInput: return "Input";
    case TaskType::kCompute: return "Compute";
    case TaskType::kOutput: return "Output";
    case TaskType::kSend: return "Send";
    case TaskType::kRecv: return "Recv";
    case TaskType::kGAInit: return "GAInit";
    case TaskType::kGA: return "GA";
    case TaskType::kAG: return "AG";
    default: return "Merge";
  }
}

int TaskDAG::AddNode(TaskType t, const std::string& name, int stage, int micro, bool backward, double cost, double out_bytes) {
  TaskNode n;
  n.id = (int)nodes.size();
  n.type = t; n.name = name; n.stage = stage; n.micro = micro; n.backward = backward; n.cost = cost; n.out_bytes = out_bytes;
  nodes.push_back(n);
  return n.id;
}
void TaskDAG::AddEdge(int from, int to) {
  nodes[from].children.push_back(to);
  nodes[to].parents.push_back(from);
}
std::vector<int> TaskDAG::TopoOrder() const {
  std::vector<int> indeg(nodes.size()), order;
  for (auto& n : nodes) indeg[n.id] = (int)n.parents.size();
  std::queue<int> q;
  for (auto& n : nodes)
    if (!indeg[n.id]) q.push(n.id);
  while (!q.empty()) {
    int u = q.front();
    q.pop();
    order.push_back(u);
    for (int c : nodes[u].children)
      if (--indeg[c] == 0) q.push(c);
  }
  return order;
}
std::vector<int> TaskDAG::BuildDominanceTree() const {
  // Cooper, Harvey, Kennedy: "A Simple, Fast Dominance Algorithm"
  std::vector<int> order = TopoOrder();  // a DAG's topological order is a reverse post-order
  std::vector<int> rpo(nodes.size(), -1);
  for (int i = 0; i < (int)order.size(); ++i) rpo[order[i]] = i;
  std::vector<int> idom(nodes.size(), -1);
  const int root = source >= 0 ? source : (order.empty() ? -1 : order[0]);
  if (root < 0) return idom;
  idom[root] = root;
  auto intersect = [&](int a, int b) {
    while (a != b) {
      while (rpo[a] > rpo[b]) a = idom[a];
      while (rpo[b] > rpo[a]) b = idom[b];
    }
    return a;
  };
  bool changed = true;
  while (changed) {
    changed = false;
    for (int u : order) {
      if (u == root) continue;
      int nd = -1;
      for (int p : nodes[u].parents) {
        if (idom[p] < 0) continue;
        nd = nd < 0 ? p : intersect(p, nd);
      }
      if (nd >= 0 && idom[u] != nd) { idom[u] = nd; changed = true; }
    }
  }
  return idom;
}
std::string TaskDAG::ToDot() const {
  std::ostringstream o;
  o << "digraph tasks {\n  rankdir=LR;\n";
  for (auto& n : nodes)
    o << "  t" << n.id << " [label=\"" << TaskTypeName(n.type) << "\\n" << n.name << "\\ndev" << n.device << "\"];\n";
  for (auto& n : nodes)
    for (int c : n.children) o << "  t" << n.id << " -> t" << c << ";\n";
  o << "}\n";
  return o.str();
}

TaskDAG BuildPipelineTaskDAG(const PipelineSpec& sp) {
  TaskDAG d;
  const int S = sp.num_stages, M = sp.num_micro;
  auto dev = [&](int s) { return s * sp.spmd; };
  auto at = [](const std::vector<double>& v, int i, double def = 0.0) { return i < (int)v.size() ? v[i] : def; };
  d.source = d.AddNode(TaskType::kSplit, "split", 0, -1, false, 0, 0);
  std::vector<int> gainit(S), ag(S);
  for (int s = 0; s < S; ++s) {
    gainit[s] = d.AddNode(TaskType::kGAInit, "gainit.s" + std::to_string(s), s, -1, false, 1e-6, 0);
    d.nodes[gainit[s]].device = dev(s);
    d.AddEdge(d.source, gainit[s]);
  }
  std::vector<std::vector<int>> f_out(S, std::vector<int>(M)), b_out(S, std::vector<int>(M)), f_cmp(S, std::vector<int>(M));
  auto bundle = [&](int s, int m, bool bwd) {
    const std::string tag = std::string(bwd ? "B" : "F") + ".s" + std::to_string(s) + ".m" + std::to_string(m);
    int in = d.AddNode(TaskType::kInput, "in." + tag, s, m, bwd, 0, 0);
    int cp = d.AddNode(TaskType::kCompute, tag, s, m, bwd, bwd ? at(sp.bwd_seconds, s) : at(sp.fwd_seconds, s),
                       bwd ? 0.0 : at(sp.act_bytes, s));
    int out = d.AddNode(TaskType::kOutput, "out." + tag, s, m, bwd, 0, 0);
    for (int t : {in, cp, out}) d.nodes[t].device = dev(s);
    d.AddEdge(in, cp);
    d.AddEdge(cp, out);
    return std::array<int, 3>{in, cp, out};
  };
  for (int m = 0; m < M; ++m) {
    // forward chain
    int prev_out = -1;
    for (int s = 0; s < S; ++s) {
      auto b = bundle(s, m, false);
      f_cmp[s][m] = b[1];
      f_out[s][m] = b[2];
      if (s == 0) d.AddEdge(d.source, b[0]);
      else {
        const double bytes = at(sp.boundary_bytes, s - 1);
        const double c = sp.p2p_latency + bytes / sp.p2p_bw;
        int snd = d.AddNode(TaskType::kSend, "send.F.s" + std::to_string(s - 1) + ".m" + std::to_string(m), s - 1, m, false, c, 0);
        int rcv = d.AddNode(TaskType::kRecv, "recv.F.s" + std::to_string(s) + ".m" + std::to_string(m), s, m, false, c, bytes);
        d.nodes[snd].device = dev(s - 1); d.nodes[snd].peer_device = dev(s);
        d.nodes[rcv].device = dev(s); d.nodes[rcv].peer_device = dev(s - 1);
        d.AddEdge(prev_out, snd);
        d.AddEdge(snd, rcv);
        d.AddEdge(rcv, b[0]);
      }
      prev_out = b[2];
    }
    // backward chain (mirror order)
    int prev_b_out = -1;
    for (int s = S - 1; s >= 0; --s) {
      auto b = bundle(s, m, true);
      b_out[s][m] = b[2];
      d.AddEdge(f_out[s][m], b[0]);  // stashed activations of the same stage
      if (s == S - 1) {
        // loss lives on the last stage: backward starts right after its forward
      } else {
        const double bytes = at(sp.boundary_bytes, s);
        const double c = sp.p2p_latency + bytes / sp.p2p_bw;
        int snd = d.AddNode(TaskType::kSend, "send.B.s" + std::to_string(s + 1) + ".m" + std::to_string(m), s + 1, m, true, c, 0);
        int rcv = d.AddNode(TaskType::kRecv, "recv.B.s" + std::to_string(s) + ".m" + std::to_string(m), s, m, true, c, bytes);
        d.nodes[snd].device = dev(s + 1); d.nodes[snd].peer_device = dev(s);
        d.nodes[rcv].device = dev(s); d.nodes[rcv].peer_device = dev(s + 1);
        d.AddEdge(prev_b_out, snd);
        d.AddEdge(snd, rcv);
        d.AddEdge(rcv, b[0]);
      }
      prev_b_out = b[2];
      int ga = d.AddNode(TaskType::kGA, "ga.s" + std::to_string(s) + ".m" + std::to_string(m), s, m, true, 1e-6, 0);
      d.nodes[ga].device = dev(s);
      d.AddEdge(b[2], ga);
      d.AddEdge(gainit[s], ga);
    }
  }
  d.sink = d.AddNode(TaskType::kMerge, "merge", 0, -1, false, 0, 0);
  for (int s = 0; s < S; ++s) {
    ag[s] = d.AddNode(TaskType::kAG, "ag.s" + std::to_string(s), s, -1, false, at(sp.ag_seconds, s), 0);
    d.nodes[ag[s]].device = dev(s);
    for (auto& n : d.nodes)
      if (n.type == TaskType::kGA && n.stage == s) d.AddEdge(n.id, ag[s]);
    d.AddEdge(ag[s], d.sink);
  }
  return d;
}

std::vector<std::vector<int>> ComputeReleasePlan(const TaskDAG& dag, const std::vector<int>& order) {
  std::vector<int> refs(dag.nodes.size());
  for (auto& n : dag.nodes) refs[n.id] = (int)n.children.size();
  std::vector<std::vector<int>> rel(order.size());
  for (size_t i = 0; i < order.size(); ++i) {
    const TaskNode& n = dag.nodes[order[i]];
    for (int p : n.parents)
      if (--refs[p] == 0) rel[i].push_back(p);
    if (n.children.empty()) rel[i].push_back(n.id);
  }
  return rel;
}

std::string Schedule::Dump(const TaskDAG& dag) const {
  std::ostringstream o;
  o << "makespan=" << makespan * 1e3 << "ms bubble=" << bubble_ratio << (oom ? " OOM" : "") << "\n";
  for (auto& kv : device_tasks) {
    o << "dev" << kv.first << ":";
    for (int t : kv.second)
      if (dag.nodes[t].type == TaskType::kCompute || dag.nodes[t].type == TaskType::kAG) o << " " << dag.nodes[t].name;
    o << "\n";
  }
  return o.str();
}

Schedule ScheduleTasks(TaskDAG* dagp, const PipelineSpec& sp, const ScheduleOptions& opt) {
  TaskDAG& dag = *dagp;
  Schedule sch;
  const int N = (int)dag.nodes.size();
  const int limit = opt.micro_num_limit > 0 ? opt.micro_num_limit : sp.num_stages;
  sch.start.assign(N, -1);
  sch.finish.assign(N, -1);
  std::vector<int> remaining(N);
  for (auto& n : dag.nodes) remaining[n.id] = (int)n.parents.size();
  std::map<int, double> dev_free;           // device -> time it becomes free
  // GROUP_SCHED_COUNT (reference task_scheduler.cc:125 should_ignore_by_sched_id, :1313): micro-batch m belongs to group
  // m % G; the reference schedules every group on its own and merges the per-device sequences, i.e. each group is an
  // independent 1F1B stream.  Here one event-driven pass schedules all of them, with the admission window kept per group.
  const int G = std::max(1, opt.group_sched_count);
  std::map<std::pair<int, int>, int> active_fwd;   // (device, group) -> forward micro-batches admitted, backward not yet run
  std::map<int, double> live_bytes;
  std::set<int> ready;
  for (auto& n : dag.nodes) {
    if (!remaining[n.id]) ready.insert(n.id);
    dev_free[n.device] = 0;
  }
  std::vector<double> ready_time(N, 0.0);
  int done = 0;
  // priority key (smaller = earlier): GA (EARLY_GA, the default here; reference task_scheduler.cc:1367 ReorderGA hoists GA
  // next to its micro-batch's output only when the flag is set), then backward compute bundles, then forward (by micro
  // id), sends/recvs follow their producers, AG last
  auto prio = [&](const TaskNode& t) {
    int cls;
    switch (t.type) {
      case TaskType::kGA: cls = opt.early_ga ? 0 : 4; break;   // lazy: behind any ready compute of the device
      case TaskType::kGAInit: case TaskType::kSplit: cls = 0; break;
      case TaskType::kSend: cls = 1; break;
      case TaskType::kRecv: cls = 1; break;
      case TaskType::kAG: cls = 9; break;
      case TaskType::kMerge: cls = 10; break;
      default: cls = t.backward ? 2 : 3;
    }
    return std::make_tuple(cls, t.micro < 0 ? 0 : t.micro, t.id);
  };
  while (done < N) {
    // candidate = ready task with the smallest (available time, priority) that passes admission control
    int best = -1;
    double best_t = 0;
    for (int id : ready) {
      const TaskNode& t = dag.nodes[id];
      // 1F1B admission: a forward Input is held back while `limit` forward micro-batches are already in flight
      const auto grp = std::make_pair(t.device, t.micro < 0 ? 0 : t.micro % G);
      if (t.type == TaskType::kInput && !t.backward && active_fwd[grp] >= limit - t.stage && limit - t.stage > 0) {
        bool other_work = false;
        for (int o : ready)
          if (o != id && dag.nodes[o].device == t.device && !(dag.nodes[o].type == TaskType::kInput && !dag.nodes[o].backward)) other_work = true;
        bool pending_bwd = active_fwd[grp] > 0;
        if (other_work || pending_bwd) continue;
      }
      const double avail = std::max(ready_time[id], dev_free[t.device]);
      if (best < 0 || avail < best_t - 1e-12 || (std::fabs(avail - best_t) <= 1e-12 && prio(t) < prio(dag.nodes[best]))) {
        best = id;
        best_t = avail;
      }
    }
    if (best < 0) {  // everything ready is blocked by admission control: release the oldest forward
      for (int id : ready)
        if (best < 0 || prio(dag.nodes[id]) < prio(dag.nodes[best])) best = id;
      best_t = std::max(ready_time[best], dev_free[dag.nodes[best].device]);
    }
    TaskNode& t = dag.nodes[best];
    ready.erase(best);
    sch.start[best] = best_t;
    // sends / recvs run on side streams: they do not occupy the compute timeline of the device
    const bool side = t.type == TaskType::kSend || t.type == TaskType::kRecv;
    sch.finish[best] = best_t + t.cost;
    if (!side) dev_free[t.device] = sch.finish[best];
    sch.device_tasks[t.device].push_back(best);
    const auto tgrp = std::make_pair(t.device, t.micro < 0 ? 0 : t.micro % G);
    if (t.type == TaskType::kInput && !t.backward) active_fwd[tgrp]++;
    if (t.type == TaskType::kOutput && t.backward) active_fwd[tgrp]--;
    if (t.type == TaskType::kCompute) {
      if (!t.backward) live_bytes[t.device] += t.out_bytes;
      else live_bytes[t.device] -= (t.stage < (int)sp.act_bytes.size() ? sp.act_bytes[t.stage] : 0.0);
      sch.peak_bytes[t.device] = std::max(sch.peak_bytes[t.device], live_bytes[t.device]);
    }
    ++done;
    for (int c : t.children) {
      ready_time[c] = std::max(ready_time[c], sch.finish[best]);
      if (--remaining[c] == 0) ready.insert(c);
    }
    sch.makespan = std::max(sch.makespan, sch.finish[best]);
  }
  if (opt.reorder_send) {
    // ReorderSend: within a device list move every Send right behind the task that produced its payload
    for (auto& kv : sch.device_tasks) {
      auto& lst = kv.second;
      for (size_t i = 0; i < lst.size(); ++i) {
        if (dag.nodes[lst[i]].type != TaskType::kSend) continue;
        const int prod = dag.nodes[lst[i]].parents.empty() ? -1 : dag.nodes[lst[i]].parents[0];
        size_t p = i;
        while (p > 0 && lst[p - 1] != prod && dag.nodes[lst[p - 1]].type != TaskType::kSend &&
               dag.nodes[lst[p - 1]].type != TaskType::kRecv) {
          std::swap(lst[p], lst[p - 1]);
          --p;
        }
      }
    }
  }
  if (opt.buffer_save) {
    // recv buffers of the same (stage, direction) class rotate through a ring of `ring` slots; slot = occurrence index mod
    // ring; a slot is "reused" from the (ring+1)-th receive of its class on.  The reference sizes the ring with
    // GROUP_SCHED_COUNT (execution_plan.cc:203: one buffer per group, consumed by the group's next Input); here a received
    // activation stays referenced until its micro-batch's backward, so the ring has to cover everything in flight:
    // groups x in-flight limit.  `recv_ring` overrides that (an undersized ring is legal: the stage worker bypasses an occupied
    // slot with a fresh buffer and counts a miss).
    const int ring = std::max(1, opt.recv_ring > 0 ? opt.recv_ring : G * limit);
    std::map<std::pair<int, bool>, int> counter;
    for (auto& kv : sch.device_tasks)
      for (int id : kv.second)
        if (dag.nodes[id].type == TaskType::kRecv) {
          auto key = std::make_pair(dag.nodes[id].stage, dag.nodes[id].backward);
          const int occ = counter[key]++;
          dag.nodes[id].buffer_id = occ % ring;
          dag.nodes[id].buffer_reused = occ >= ring;
        }
  }
  // GC plan per device
  for (auto& kv : sch.device_tasks) {
    auto rel = ComputeReleasePlan(dag, kv.second);
    for (size_t i = 0; i < kv.second.size(); ++i) dag.nodes[kv.second[i]].mem_to_release = rel[i];
  }
  double busy = 0;
  for (auto& n : dag.nodes)
    if (n.type == TaskType::kCompute || n.type == TaskType::kAG) busy += n.cost;
  const int ndev = (int)sch.device_tasks.size();
  sch.bubble_ratio = sch.makespan > 0 ? std::max(0.0, 1.0 - busy / (sch.makespan * std::max(1, ndev))) : 0;
  for (auto& kv : sch.peak_bytes)
    if (kv.second > sp.mem_limit) sch.oom = true;
  return sch;
}

}  // namespace tepdist
