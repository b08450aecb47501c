// Communication / compute cost model.
//
// Reference parity: Cost() / DataTransferSize() (cost_spmd_strategy.cc:74-174),
// PerfUtils (performance_utils.cc), Evaluator constants (evaluator.h:45-57).  Re-parameterised for one
// 8xB200 NVSwitch box (single bandwidth tier + launch latency); the reference's V100 numbers stay selectable
// as the "reference" profile.
#pragma once
#include <string>

#include "ir.h"

namespace tepdist {

struct HwProfile {
  std::string name = "b200";
  double flops = 1.4e15;         // sustained dense bf16 FLOP/s per GPU (MEASURED_PEAKS bf16_tflops_sustained)
  double hbm_bw = 6.4e12;        // B/s
  double link_bw = 7.7e11;       // B/s per direction per GPU through NVSwitch (measured peer copy)
  double inter_node_bw = 5.0e10; // B/s (unused on one box)
  double coll_latency = 8e-6;    // s per collective launch/sync
  double mem_bytes = 180e9 * 0.9;
  // Share of the SPMD collectives' wire time that is NOT hidden under compute, by group size: measured on this pool's B200s with
  // bench.py's dry-comm subtraction (GPT-2 345M, batch 4 x 1024 per GPU, fused bucketed reduce-scatter + AdamW + all-gather on a
  // side stream; profiles/README.md): exposed 0.45 / 1.44 / 2.22 ms of 1.38 / 2.07 / 2.42 ms of wire time at n = 2 / 4 / 8.
  double ExposedCommFraction(int n) const {
    if (name == "reference_v100") return 1.0;          // the reference issues every collective in-stream
    if (n <= 2) return 0.33;
    if (n <= 4) return 0.70;
    return 0.90;
  }
  // Tensor-core work on few rows runs far below the sustained rate (tile-count quantisation on 148 SMs, per-kernel prologue /
  // epilogue of ~10 us against ~4 us of MMAs): time = flops / rate * (1 + half_rows / rows).  half_rows = 4096 fits the two
  // measurements we have for GPT-2 345M layers: 4096 rows per step at ~50 % of the sustained rate (17 ms data parallel) and
  // 1024 rows per micro-batch 2.5x slower per row (42 ms, 2-stage pipeline x 8 micro-batches; profiles/README.md).
  double small_batch_half_rows = 4096;
  double ComputeSlowdown(double rows) const { return (rows > 0 && name != "reference_v100") ? 1.0 + small_batch_half_rows / rows : 1.0; }
  static HwProfile B200() { return HwProfile(); }
  static HwProfile ReferenceV100() {
    HwProfile h;
    h.name = "reference_v100";
    h.flops = 15e12; h.hbm_bw = 9e11; h.link_bw = 300e9; h.inter_node_bw = 3.125e9; h.coll_latency = 0;
    h.mem_bytes = 32e9 * 0.9;
    return h;
  }
};

enum class Reshard { kNone, kDynamicSlice, kAllGather, kAllToAll, kAllReduce, kReduceScatter, kInvalid };
const char* ReshardName(Reshard r);

// What it takes to turn a value laid out as `from` into layout `to`.
Reshard ClassifyReshard(const DimStrategy& from, const DimStrategy& to);
// Bytes each device moves for that reshard (B = full tensor bytes, n = group size).
//   AR = 2B(n-1)/n   AG = B - B/n   RS = B(n-1)/n   A2A = (B/n - B/n^2) * cost_factor   slice ~ 0
double ReshardBytes(Reshard kind, double full_bytes, int n, double cost_factor = 1.0);
double ReshardCost(const DimStrategy& from, const DimStrategy& to, double full_bytes, int n, double cost_factor = 1.0);
// Seconds for a collective of `bytes` per device.
double CollectiveSeconds(const HwProfile& hw, double bytes, bool spans_nodes = false);

constexpr double kInfCost = 1e30;

}  // namespace tepdist
