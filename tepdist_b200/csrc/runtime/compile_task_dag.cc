// CompileTaskDAG: the task graph of a pipeline plan built from the DefContext tree (not from a hand-filled template).
//
// Reference parity: VirtualClient::CompileTaskDAG (xla/pjrt/virtual_client.cc:613-772) walks the DefContext
// tree -- one Input/Compute/Output bundle per (CG_SLICE_<s>_{F,B}, micro-batch), GAInit / AG per AG_SLICE_<s> -- and
// CrossDeviceCalibration (xla/pjrt/execution_plan.cc:492-600) inserts the Send/Recv pairs for values that cross devices.
// Here the per-task costs come from the contexts themselves (their FLOPs over the device rate; the boundary bytes are the
// StageTransfer list of StageDecompose, i.e. exactly the values the stage workers will put on the wire), every compute task
// records the DefContext it executes (TaskNode::def_ctx), and the resulting PipelineSpec is returned so that the scheduler
// prices the same numbers.
#include <algorithm>
#include <stdexcept>

#include "auto_parallel.h"
#include "runtime/task_graph.h"

namespace tepdist {

TaskDAG CompileTaskDAG(const Graph& g, const Decomposition& d, const std::vector<StageTransfer>& xfers, int num_micro, int spmd,
                       const HwProfile& hw, PipelineSpec* spec_out) {
  int S = 0;
  for (auto& c : d.ctx) S = std::max(S, c.stage + 1);
  if (S == 0) S = 1;
  std::vector<int> ctx_f(S, -1), ctx_b(S, -1), ctx_a(S, -1);
  for (int i = 0; i < (int)d.ctx.size(); ++i) {
    const DefContext& c = d.ctx[i];
    if (c.stage < 0) continue;
    if (c.kind == "stage_fwd") ctx_f[c.stage] = i;
    else if (c.kind == "stage_bwd") ctx_b[c.stage] = i;
    else if (c.kind == "stage_ag") ctx_a[c.stage] = i;
  }
  PipelineSpec sp;
  sp.num_stages = S;
  sp.num_micro = std::max(1, num_micro);
  sp.spmd = std::max(1, spmd);
  sp.p2p_bw = hw.link_bw;
  sp.mem_limit = hw.mem_bytes;
  // `g` is the TRANSFORMED graph (micro-batch level and SPMD level applied): its shapes -- hence the contexts' FLOPs and
  // bytes -- are already per device and per micro-batch
  const double rate = hw.flops;
  auto secs = [&](int ci) { return ci >= 0 ? d.ctx[ci].gflops * 1e9 / rate : 0.0; };
  for (int s = 0; s < S; ++s) {
    if (S > 1 && (ctx_f[s] < 0 || ctx_b[s] < 0)) throw std::runtime_error("CompileTaskDAG: stage without CG_SLICE contexts");
    sp.fwd_seconds.push_back(std::max(secs(ctx_f[s]), 1e-7));
    sp.bwd_seconds.push_back(std::max(secs(ctx_b[s]), 1e-7));
    // optimizer: memory-bound (reads / writes every state byte of the stage once)
    double ag_bytes = ctx_a[s] >= 0 ? d.ctx[ctx_a[s]].in_bytes + d.ctx[ctx_a[s]].out_bytes : 0.0;
    sp.ag_seconds.push_back(std::max(ag_bytes / hw.hbm_bw, 1e-6));
    // activations a micro-batch keeps alive between its forward and its backward on this stage: what the backward context
    // reads from the forward context of the same stage
    double act = 0;
    if (ctx_b[s] >= 0 && ctx_f[s] >= 0) {
      const auto& fn = d.ctx[ctx_f[s]].nodes;
      for (auto& v : d.ctx[ctx_b[s]].inputs)
        if (std::binary_search(fn.begin(), fn.end(), v.node)) act += (double)g.type(v).bytes();
    }
    sp.act_bytes.push_back(act);
  }
  sp.boundary_bytes.assign(std::max(0, S - 1), 0.0);
  for (auto& t : xfers)
    if (!t.backward && t.from_stage >= 0 && t.from_stage < S - 1) sp.boundary_bytes[t.from_stage] += t.bytes;
  TaskDAG dag = BuildPipelineTaskDAG(sp);
  for (auto& n : dag.nodes) {
    if (n.stage < 0 || n.stage >= S) continue;
    if (n.type == TaskType::kCompute || n.type == TaskType::kInput || n.type == TaskType::kOutput)
      n.def_ctx = n.backward ? ctx_b[n.stage] : ctx_f[n.stage];
    else if (n.type == TaskType::kAG || n.type == TaskType::kGAInit || n.type == TaskType::kGA)
      n.def_ctx = ctx_a[n.stage];
  }
  if (spec_out) *spec_out = sp;
  return dag;
}

}  // namespace tepdist
