// AutoParallel orchestrator + sync-free (micro-batch) analysis + plan evaluator + decompositions.
//
// Reference parity (§2.B B3, B4; §2.C C2):
//   AutoParallel::Run            auto_parallel.cc:395-409  (rule / config / exploration modes)
//   GenerateSplitProposals       auto_parallel.cc:132-181  (BFS over power-of-2 factorisations, first level = stages)
//   SyncFreeSplittingAnalysis    sync_free_splitting_analysis.cc
//   Evaluator                    evaluator.{h,cc}
//   SyncFreeDecomposition        CG / GAINIT / GA / AG
//   StageDecomposition           per-stage forward/backward sub-graphs, neighbour-only transfers
//   DefContext                   the descriptor tree Entry -> {CG, GAINIT, GA, AG} -> *_SLICE stage children
#pragma once
#include <map>
#include <string>
#include <vector>

#include "cost.h"
#include "ir.h"
#include "spmd_planner.h"
#include "stage_planner.h"
#include "transform.h"

namespace tepdist {

// ---------------------------------------------------------------- sync-free (micro-batch) analysis
struct SyncFreeResult {
  bool ok = false;
  int num_micro = 1;
  std::map<int, int> input_split_dim;  // input node id -> dim split across micro-batches
  int num_split_values = 0;            // how many values the chosen proposal splits
  std::vector<ValueRef> sync_points;   // partial values: accumulate across micro-batches here
  std::string reason;
  SpmdPlan plan;                       // level plan (every value: split / partial / glue)
};
SyncFreeResult SyncFreeAnalysis(const Graph& g, int num_micro);

// ---------------------------------------------------------------- DefContext tree
struct DefContext {
  std::string name;                 // ENTRY, CG, GAINIT, GA, AG, CG_SLICE_<stage>_{F,B}, AG_SLICE_<stage>
  std::string kind;                 // "entry" | "cg" | "gainit" | "ga" | "ag" | "stage_fwd" | "stage_bwd" | "stage_ag"
  int stage = -1;
  std::vector<int> nodes;           // graph nodes that belong to this context (topological order)
  std::vector<ValueRef> inputs;     // values consumed from outside (variables, samples, other contexts)
  std::vector<ValueRef> outputs;    // values produced for other contexts / fetches
  std::map<int, int> input_def;     // index in `inputs` -> producing child context index (-1: entry argument)
  double gflops = 0;
  double in_bytes = 0, out_bytes = 0;
  std::vector<int> children;        // indices into the owning vector
  bool per_micro_batch = false;     // runs once per micro-batch (CG) vs once per step (GAINIT, AG)
};
struct Decomposition {
  std::vector<DefContext> ctx;      // ctx[0] = ENTRY
  std::vector<ValueRef> accumulators;  // gradient (and loss) values accumulated across micro-batches
  std::string Dump() const;
};
// B3: cut at the sync points (collectives of the micro-batch level / gradient inputs of apply nodes).
Decomposition SyncFreeDecompose(const Graph& g, int micro_level);
// B4: split CG / AG per pipeline stage; cross-stage values are threaded through every intermediate stage so only
// neighbours exchange data.  Appends stage children to the decomposition and returns the transfer list.
struct StageTransfer {
  ValueRef value;
  int from_stage, to_stage;  // always |to - from| == 1 after threading
  bool backward;             // gradient flowing to an earlier stage
  double bytes;
};
std::vector<StageTransfer> StageDecompose(const Graph& g, int num_stages, Decomposition* d);

// ---------------------------------------------------------------- evaluator
struct EvalResult {
  bool feasible = true;
  double total_duration = 0;   // seconds per step
  double compute_time = 0, comm_time = 0, p2p_time = 0;
  double gpu_efficiency = 0, coll_ratio = 0, bubble_ratio = 0;
  double mem_bytes_per_device = 0;
  std::string str() const;
};
struct EvalInput {
  int num_stages = 1, num_micro = 1, spmd = 1;
  double spmd_comm_bytes = 0;          // per device per micro-batch (from the SPMD plan)
  double exposed_comm_fraction = 1.0;  // reference executes collectives in-stream: fully exposed
  std::vector<double> stage_flops;     // per stage, whole step, unsharded
  double cut_bytes = 0;                // per micro-batch across all boundaries
  double var_bytes = 0;                // variables + slots + grads, whole model
  double act_bytes = 0;                // activations stashed per micro-batch, whole model
  double rows_per_micro = 0;           // GEMM rows (tokens / pixels) one device processes per micro-batch; 0 = unknown
};
EvalResult Evaluate(const EvalInput& in, const HwProfile& hw);

// ---------------------------------------------------------------- orchestrator
struct DeviceSplitProposal {
  int stages = 1, spmd = 1, micro = 1;
  std::string str() const;
};
std::vector<DeviceSplitProposal> GenerateSplitProposals(int num_devices, int64_t batch, bool allow_pipeline);

struct AutoParallelOptions {
  int num_devices = 1;
  std::string mode = "exploration";  // "exploration" | "config" | "rule"
  int num_stages = 0, num_micro_batches = 0;  // config mode
  SpmdOptions spmd;
  double unbalanced_ratio = 0.08;
  bool allow_pipeline = true;
  bool spmd_rule_mode = false;   // SPMD level by annotation propagation (batch split) instead of the cost-based planner
  double exposed_comm_fraction = -1;   // evaluator: < 0 = the hardware profile's calibrated value for the SPMD group size
  HwProfile hw;
};
struct ParallelPlan {
  DeviceSplitProposal proposal;
  Graph graph;                 // transformed: micro-batch level + SPMD level applied, Node::stage set
  EvalResult eval;
  SpmdStats spmd_stats;
  StagePlanResult stage_plan;
  SyncFreeResult sync_free;
  std::vector<std::pair<std::string, double>> candidates;  // every proposal with its estimated duration
  std::string log;
};
ParallelPlan AutoParallelRun(const Graph& g, const AutoParallelOptions& opt);

}  // namespace tepdist
