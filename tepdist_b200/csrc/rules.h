// Per-op sharding rules: the table that decides which plans the planner can find.
//
// Reference parity (SURVEY §2.A A4 + Appendix A): StrategyUtil::ForwardInfer / BackInfer for ~45 opcodes,
// GenDotProposals / GenConvProposals, VerifyInfer (xla/service/parallel/utils.cc).  Here every op exposes its
// full set of self-consistent (operand strategies, output strategies) combinations for one mesh level
// ("candidates"); ForwardInfer / BackInfer are queries over that set, so the two directions can never
// disagree (which the reference guards with VerifyInfer round trips).
#pragma once
#include <string>
#include <vector>

#include "ir.h"

namespace tepdist {

struct Candidate {
  std::vector<DimStrategy> ins;   // required layout of every operand
  std::vector<DimStrategy> outs;  // produced layout of every output
  double node_cost = 0;           // extra cost paid inside the op (e.g. sync-BN statistics all-reduce), bytes
  std::string tag;                // "batch", "contract", "row", "col", "glue", "dim2", ...
};

struct RuleOptions {
  bool allow_glue_compute_intensive = false;  // dots/convs never run replicated (reference: no Glue candidate)
  bool save_variable_mem = false;             // suppress batch-split proposals (reference: split_for_mem_save_)
  bool sequence_parallel = false;             // row-parallel linears may reduce-scatter over a token dim inside the node ("contract_rs<d>")
  bool context_parallel = false;              // attention keeps a sequence split (ring over K / V) instead of resharding to heads
};

// All candidates of node `n` for a mesh level of `num` devices.
std::vector<Candidate> EnumerateCandidates(const Graph& g, const Node& n, int num, const RuleOptions& opt = RuleOptions());

// Given the strategy of operand `operand_idx`, the consistent candidates (peer operands + outputs).
std::vector<Candidate> ForwardInfer(const Graph& g, const Node& n, int num, int operand_idx, const DimStrategy& s);
// Given the strategy of output `out_idx`, the consistent candidates (operand requirements).
std::vector<Candidate> BackInfer(const Graph& g, const Node& n, int num, int out_idx, const DimStrategy& s);

// Shape of one shard of `t` under `s` (reference DistUtil::MakeNewShape).
TensorType ShardType(const TensorType& t, const DimStrategy& s);

// Whole-graph fixpoint propagation from seeds (reference InferGraph, utils.cc:2036-2146): alternating forward and
// backward sweeps; returns false on conflict.  `assign` maps value -> strategy (absent = undecided).
bool InferGraph(const Graph& g, int num, std::map<ValueRef, DimStrategy>* assign, std::string* conflict = nullptr);

// Ops that hit the "replicate everything" fallback since the last call (the planner surfaces them as a warning).
std::vector<std::string> UnknownOps(bool clear = true);

// Round-trip / shape-algebra check of the rule table over every node of `g` (reference VerifyInfer); returns the violations.
std::vector<std::string> VerifyInfer(const Graph& g, int num);

}  // namespace tepdist
