#include "cost.h"

namespace tepdist {

const char* ReshardName(Reshard r) {
  switch (r) {
    case Reshard::kNone: return "none";
    case Reshard::kDynamicSlice: return "dynamic_slice";
    case Reshard::kAllGather: return "all_gather";
    case Reshard::kAllToAll: return "all_to_all";
    case Reshard::kAllReduce: return "all_reduce";
    case Reshard::kReduceScatter: return "reduce_scatter";
    default: return "invalid";
  }
}

Reshard ClassifyReshard(const DimStrategy& from, const DimStrategy& to) {
  if (to.partial) return from.partial ? Reshard::kNone : Reshard::kInvalid;
  if (from.partial) return to.is_glue() ? Reshard::kAllReduce : Reshard::kReduceScatter;
  if (from == to) return Reshard::kNone;
  if (from.is_glue()) return Reshard::kDynamicSlice;  // replicated -> sharded: local slice, no traffic
  if (to.is_glue()) return Reshard::kAllGather;
  return Reshard::kAllToAll;
}

double ReshardBytes(Reshard kind, double B, int n, double cost_factor) {
  switch (kind) {
    case Reshard::kNone: return 0;
    case Reshard::kDynamicSlice: return 10.0;  // reference: "+10" so equal plans prefer no slice
    case Reshard::kAllGather: return B - B / n;
    case Reshard::kAllToAll: return (B / n - B / ((double)n * n)) * cost_factor;
    case Reshard::kAllReduce: return 2.0 * B * (n - 1) / n;
    case Reshard::kReduceScatter: return B * (n - 1) / n;
    default: return kInfCost;
  }
}

double ReshardCost(const DimStrategy& from, const DimStrategy& to, double full_bytes, int n, double cost_factor) {
  return ReshardBytes(ClassifyReshard(from, to), full_bytes, n, cost_factor);
}

double CollectiveSeconds(const HwProfile& hw, double bytes, bool spans_nodes) {
  if (bytes <= 0) return 0;
  return hw.coll_latency + bytes / (spans_nodes ? hw.inter_node_bw : hw.link_bw);
}

}  // namespace tepdist
