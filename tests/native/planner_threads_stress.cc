// ThreadSanitizer driver for the planner's worker pool (SpmdOptions::num_threads, spmd_planner.cc): a transformer-shaped forward
// graph (LayerNorm -> column-parallel linear -> GELU -> row-parallel linear + residual, L layers) planned with 1 and with 6 threads;
// the two plans must be identical.  Built by bench/sanitize_native.sh.
#include <cstdio>
#include <string>
#include <vector>

#include "ir.h"
#include "spmd_planner.h"

using namespace tepdist;

static Graph Build(int layers, int64_t B, int64_t S, int64_t C) {
  Graph g;
  auto T = [](std::vector<int64_t> d, const char* t = "bf16") { return TensorType{std::move(d), t}; };
  auto node = [&](const std::string& op, std::vector<ValueRef> in, std::vector<TensorType> out, std::map<std::string, Attr> a, const std::string& name) {
    return ValueRef{g.AddNode(op, in, out, a, name, /*group=*/(int)g.nodes.size(), /*backward=*/false), 0};
  };
  ValueRef x = node("input", {}, {T({B, S, C})}, {}, "x");
  for (int l = 0; l < layers; ++l) {
    const std::string p = "h" + std::to_string(l) + "/";
    ValueRef gam = node("parameter", {}, {T({C}, "f32")}, {}, p + "ln/g"), bet = node("parameter", {}, {T({C}, "f32")}, {}, p + "ln/b");
    ValueRef ln = node("layernorm", {x, gam, bet}, {T({B, S, C})}, {{"eps", 1e-5}}, p + "ln");
    ValueRef w1 = node("parameter", {}, {T({4 * C, C})}, {}, p + "fc/w"), b1 = node("parameter", {}, {T({4 * C}, "f32")}, {}, p + "fc/b");
    ValueRef h = node("linear", {ln, w1, b1}, {T({B, S, 4 * C})}, {{"bias", true}, {"residual", false}}, p + "fc");
    ValueRef a = node("gelu", {h}, {T({B, S, 4 * C})}, {}, p + "gelu");
    ValueRef w2 = node("parameter", {}, {T({C, 4 * C})}, {}, p + "proj/w"), b2 = node("parameter", {}, {T({C}, "f32")}, {}, p + "proj/b");
    x = node("linear", {a, w2, b2, x}, {T({B, S, C})}, {{"bias", true}, {"residual", true}}, p + "proj");
  }
  g.outputs.push_back(x);
  return g;
}

int main() {
  std::vector<std::string> tags[2];
  double bytes[2] = {0, 0};
  for (int run = 0; run < 2; ++run) {
    Graph g = Build(12, 32, 1024, 1024);
    SpmdOptions o;
    o.num = 8;
    o.num_threads = run == 0 ? 1 : 6;
    o.var_mem_limit = 1.0;          // weights must be stored sharded: the tensor-parallel plan, most sub-problems
    o.mem_split_min_rank = 2;
    SpmdPlan plan = PlanSpmdLevel(&g, o);
    for (auto& c : plan.choice) tags[run].push_back(c.tag);
    bytes[run] = plan.stats.comm_bytes;
    std::printf("threads %d: %d sub-graphs (%d distinct), comm %.4g bytes, threads_used %d\n", o.num_threads, plan.stats.num_subgraphs,
                plan.stats.distinct_subgraphs, plan.stats.comm_bytes, plan.stats.threads_used);
  }
  const bool same = tags[0] == tags[1] && bytes[0] == bytes[1];
  std::printf("planner_threads_stress: %s\n", same ? "OK (identical plans)" : "FAILED (plans differ)");
  return same ? 0 : 1;
}
