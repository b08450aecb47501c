// Program transformations on planner IR.
//
// Reference parity:
//   B1 SpmdTransform            xla/service/parallel/spmd_transform.cc  (shape substitution + collective insertion)
//   B2 CustomCollectiveExpander (lowering of abstract collectives; here collectives stay first-class IR ops that
//      the runtime executes on arbitrary dims — TMA/strided kernels remove the reshape/transpose sandwiches)
//   B3 SyncFreeDecomposition    CG / GAINIT / GA / AG split at the micro-batch sync points
//   B4 StageDecomposition       per-stage sub-graphs with neighbour-only send/recv
//   B6 HloLivenessOptimizer     duplicate cheap casts of parameters per user
//   B7 DAPPLEAllReduceCombiner  bucket independent gradient collectives
#pragma once
#include <map>
#include <string>
#include <vector>

#include "ir.h"
#include "spmd_planner.h"

namespace tepdist {

struct TransformStats {
  int num_all_reduce = 0, num_all_gather = 0, num_all_to_all = 0, num_reduce_scatter = 0, num_dynamic_slice = 0;
  double comm_bytes = 0;
  std::string CommInfo() const;  // "comm_info.<ordinal>.txt" artefact
};

// Rewrites `g` (full shapes + plan for mesh level `level`) into the per-shard SPMD program.  Every rank of the
// level's device group runs the same returned graph; rank-dependence enters only through `dynamic_slice`
// (partition id), sharded variable initialisation and sharded input feeding.
Graph SpmdTransform(const Graph& g, const SpmdPlan& plan, int level, int num, TransformStats* stats = nullptr);

// B7: groups gradient collectives (all_reduce / reduce_scatter feeding apply_* nodes) into buckets of at most
// `bucket_bytes`; annotates each collective with attrs {bucket, bucket_offset} (the runtime issues one
// fused collective per bucket over the flat gradient buffer).  Returns the number of buckets.
int CombineGradientCollectives(Graph* g, int64_t bucket_bytes, int max_per_bucket = 1 << 30);

// B6: give every user of cast(parameter) its own cast so the casted copy is not live across the step.
// B6: per-user copies of convert(parameter) results of at least `min_bytes` (see transform.cc).  Renumbers nodes.
// B7, storage-order variant -- the one the runtime executes: bucket boundaries over the FLAT gradient buffer (variables in
// storage order = layer order; `offsets` are the element offsets of the regular variables, ascending, all < end).  Sizes are
// graded: the first bucket is `first` elements and every following one doubles up to `cap`, because the variables at the
// front of the buffer belong to the first layers, whose gradients the backward pass produces LAST -- whatever is still in flight
// when backward ends is exposed, so the buckets that become ready last are the small ones.  A boundary is only placed at a
// variable start that is a multiple of `gran` (= group size x alignment: every rank's chunk of a bucket stays aligned).
// Returns the boundaries [0, b1, ..., end].
std::vector<int64_t> PlanFlatBuckets(const std::vector<int64_t>& offsets, int64_t end, int64_t gran, int64_t first, int64_t cap);

int LivenessOptimize(Graph* g, int64_t min_bytes = 1 << 20);

}  // namespace tepdist
