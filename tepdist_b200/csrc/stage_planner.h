// Graph sketch + ILP pipeline-stage planner.
//
// Reference parity: GraphSketch::BuildGraphSketch / BuildFineGrainedSketch coarsen the
// graph around compute-intensive "core" instructions, op-group bookkeeping ties every backward instruction to its
// forward op, GraphSketch::StagePlan cuts the forward sketch into S stages with an ILP whose objective is the
// cross-stage traffic and whose constraints are data dependences plus a per-stage FLOPs budget
// (UNBALANCED_RATIO); backward instructions take the mirror stage of their forward group (same device).
#pragma once
#include <string>
#include <vector>

#include "ir.h"

namespace tepdist {

struct SketchNode {
  int id = 0;
  std::vector<int> members;      // forward graph nodes absorbed into this sketch node
  double fwd_flops = 0, bwd_flops = 0;
  double param_bytes = 0;
  std::string name;
};
struct SketchEdge {
  int src, dst;
  double bytes;  // forward activation bytes crossing (the backward gradient has the same size)
};
struct GraphSketch {
  std::vector<SketchNode> nodes;
  std::vector<SketchEdge> edges;
  std::vector<int> node_of;  // graph node id -> sketch node (-1 for nodes outside the forward pass)
  bool IsChain() const;
  double TotalFlops() const;
  std::string ToDot() const;  // sketch_raw.dot artefact
};

// fine = one sketch node per compute-intensive forward op; coarse = clustered between critical nodes.
GraphSketch BuildSketch(const Graph& g, bool fine_grained);

struct StagePlanOptions {
  int num_stages = 2;
  double unbalanced_ratio = 0.08;   // UNBALANCED_RATIO (8 %)
  double ilp_time_limit_s = 30.0;
  bool force_ilp = false;           // chains are solved exactly by DP unless this is set
};
struct StagePlanResult {
  std::vector<int> sketch_stage;    // per sketch node
  double cut_bytes = 0;             // objective: bytes crossing stage boundaries (per micro-batch, fwd+bwd)
  std::vector<double> stage_flops;
  std::string method;               // "dp-chain" | "ilp"
  bool optimal = true;
  double seconds = 0;
  std::vector<int> backward_stage;  // per sketch node: device of its backward op group (BackwardPlan)
  std::string backward_method;      // "mirror" | "ilp" | "ilp-timeout"
  int backward_moved = 0;
};
StagePlanResult PlanStagesOnSketch(const GraphSketch& sk, const StagePlanOptions& opt);

// Backward stage plan (reference GraphSketch::BackwardPlan, hlo_graph_sketch.cc:2169-2427, with the op-group bookkeeping of
// BuildOpGroupInfo :1339-1585 and the mirror-stage |d| linearisation :311-476).  Every forward sketch node i owns one op GROUP
// of backward instructions (its gradient ops) with `bwd_flops` and a stash of `act_bytes` activations that live on the
// forward stage sf(i).  The group is placed as ONE unit (op-group constraint) on device sb(i):
//   minimise   sum_edges bytes(e) * (sb(src) - sb(dst))         gradient traffic, neighbour hops (flows run dst -> src)
//            + sum_i act_bytes(i) * |sb(i) - sf(i)|              moving a group off its mirror stage ships its stash
//   s.t.       sb(src) <= sb(dst) for every forward edge src -> dst      (gradients flow from later to earlier stages)
//              sum_{sf(i) = k} fwd(i) + sum_{sb(i) = k} bwd(i) <= budget  (per device, forward work fixed by the forward plan)
// With cumulative binaries z[i][k] = [sb(i) <= k] both absolute values are linear (sf is a constant).  Mirroring (sb = sf) is
// always feasible when the forward plan respected the same budget on fwd + bwd, and is what the ILP returns unless moving a
// group pays -- e.g. a backward-heavy group with a small stash next to an underloaded stage.
struct BackwardPlanResult {
  std::vector<int> sketch_stage;    // sb per sketch node
  double objective = 0;             // bytes
  int moved = 0;                    // groups placed off their mirror stage
  std::vector<double> stage_flops;  // fwd + bwd per device after the backward placement
  std::string method;               // "mirror" | "ilp" | "ilp-timeout"
};
BackwardPlanResult PlanBackwardOnSketch(const GraphSketch& sk, const std::vector<int>& fwd_stage, const std::vector<double>& act_bytes,
                                        const StagePlanOptions& opt);

// Full stage planning: sketch -> plan -> write Node::stage for every node of `g` (forward ops by plan, backward ops
// by op_group mirror, variables/slots/apply by their consumers); returns the plan.
StagePlanResult PlanStages(Graph* g, const StagePlanOptions& opt);

}  // namespace tepdist
