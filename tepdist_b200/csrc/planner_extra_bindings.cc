#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "auto_parallel.h"
#include "ilp.h"
#include "stage_planner.h"
#include "runtime/task_graph.h"

namespace py = pybind11;
using namespace tepdist;

static std::vector<std::pair<int, int>> VR(const std::vector<ValueRef>& v) {
  std::vector<std::pair<int, int>> r;
  for (auto& x : v) r.push_back({x.node, x.idx});
  return r;
}

void BindPlannerExtra(py::module_& m) {
  // ---- ILP
  py::class_<IlpModel>(m, "IlpModel")
      .def(py::init<>())
      .def("add_var", &IlpModel::AddVar, py::arg("lo"), py::arg("hi"), py::arg("cost"), py::arg("integer"))
      .def("add_row", &IlpModel::AddRow)
      .def_readonly("num_vars", &IlpModel::num_vars);
  m.def("solve_ilp", [](const IlpModel& mod, double tl) {
    IlpResult r = SolveIlp(mod, tl);
    py::dict d;
    d["status"] = r.StatusName(); d["x"] = r.x; d["objective"] = r.objective; d["nodes"] = r.nodes; d["seconds"] = r.seconds;
    return d;
  }, py::arg("model"), py::arg("time_limit_s") = 60.0);
  m.def("solve_lp", [](const IlpModel& mod) {
    IlpResult r = SolveLp(mod);
    py::dict d;
    d["status"] = r.StatusName(); d["x"] = r.x; d["objective"] = r.objective;
    return d;
  });

  // ---- sketch / stage planner
  py::class_<SketchNode>(m, "SketchNode")
      .def_readonly("id", &SketchNode::id).def_readonly("members", &SketchNode::members)
      .def_readonly("fwd_flops", &SketchNode::fwd_flops).def_readonly("bwd_flops", &SketchNode::bwd_flops)
      .def_readonly("param_bytes", &SketchNode::param_bytes).def_readonly("name", &SketchNode::name);
  py::class_<SketchEdge>(m, "SketchEdge")
      .def_readonly("src", &SketchEdge::src).def_readonly("dst", &SketchEdge::dst).def_readonly("bytes", &SketchEdge::bytes);
  py::class_<GraphSketch>(m, "GraphSketch")
      .def_readonly("nodes", &GraphSketch::nodes).def_readonly("edges", &GraphSketch::edges)
      .def_readonly("node_of", &GraphSketch::node_of)
      .def("is_chain", &GraphSketch::IsChain).def("total_flops", &GraphSketch::TotalFlops).def("to_dot", &GraphSketch::ToDot);
  m.def("build_sketch", &BuildSketch, py::arg("graph"), py::arg("fine_grained") = false);
  py::class_<StagePlanOptions>(m, "StagePlanOptions")
      .def(py::init<>())
      .def_readwrite("num_stages", &StagePlanOptions::num_stages)
      .def_readwrite("unbalanced_ratio", &StagePlanOptions::unbalanced_ratio)
      .def_readwrite("ilp_time_limit_s", &StagePlanOptions::ilp_time_limit_s)
      .def_readwrite("force_ilp", &StagePlanOptions::force_ilp);
  py::class_<StagePlanResult>(m, "StagePlanResult")
      .def_readonly("sketch_stage", &StagePlanResult::sketch_stage).def_readonly("cut_bytes", &StagePlanResult::cut_bytes)
      .def_readonly("stage_flops", &StagePlanResult::stage_flops).def_readonly("method", &StagePlanResult::method)
      .def_readonly("optimal", &StagePlanResult::optimal).def_readonly("seconds", &StagePlanResult::seconds)
      .def_readonly("backward_stage", &StagePlanResult::backward_stage).def_readonly("backward_method", &StagePlanResult::backward_method)
      .def_readonly("backward_moved", &StagePlanResult::backward_moved);
  py::class_<BackwardPlanResult>(m, "BackwardPlanResult")
      .def_readonly("sketch_stage", &BackwardPlanResult::sketch_stage).def_readonly("objective", &BackwardPlanResult::objective)
      .def_readonly("moved", &BackwardPlanResult::moved).def_readonly("stage_flops", &BackwardPlanResult::stage_flops)
      .def_readonly("method", &BackwardPlanResult::method);
  m.def("plan_backward_on_sketch", &PlanBackwardOnSketch);
  m.def("make_sketch", [](const std::vector<double>& fwd, const std::vector<double>& bwd, const std::vector<std::tuple<int, int, double>>& edges) {
    GraphSketch sk;
    for (size_t i = 0; i < fwd.size(); ++i) {
      SketchNode nd;
      nd.id = (int)i; nd.fwd_flops = fwd[i]; nd.bwd_flops = bwd[i]; nd.name = "n" + std::to_string(i);
      sk.nodes.push_back(nd);
    }
    for (auto& e : edges) sk.edges.push_back({std::get<0>(e), std::get<1>(e), std::get<2>(e)});
    return sk;
  });
  m.def("plan_stages_on_sketch", &PlanStagesOnSketch);
  m.def("plan_stages", [](Graph& g, const StagePlanOptions& o) { return PlanStages(&g, o); });

  // ---- sync-free, decomposition, evaluator, orchestrator
  py::class_<SyncFreeResult>(m, "SyncFreeResult")
      .def_readonly("ok", &SyncFreeResult::ok).def_readonly("num_micro", &SyncFreeResult::num_micro)
      .def_readonly("input_split_dim", &SyncFreeResult::input_split_dim)
      .def_readonly("num_split_values", &SyncFreeResult::num_split_values)
      .def_readonly("reason", &SyncFreeResult::reason).def_readonly("plan", &SyncFreeResult::plan)
      .def("sync_points", [](const SyncFreeResult& r) { return VR(r.sync_points); });
  m.def("sync_free_analysis", &SyncFreeAnalysis);
  py::class_<DefContext>(m, "DefContext")
      .def_readonly("name", &DefContext::name).def_readonly("kind", &DefContext::kind).def_readonly("stage", &DefContext::stage)
      .def_readonly("nodes", &DefContext::nodes).def_readonly("input_def", &DefContext::input_def)
      .def_readonly("gflops", &DefContext::gflops).def_readonly("in_bytes", &DefContext::in_bytes)
      .def_readonly("out_bytes", &DefContext::out_bytes).def_readonly("children", &DefContext::children)
      .def_readonly("per_micro_batch", &DefContext::per_micro_batch)
      .def("inputs", [](const DefContext& c) { return VR(c.inputs); })
      .def("outputs", [](const DefContext& c) { return VR(c.outputs); });
  py::class_<Decomposition>(m, "Decomposition")
      .def_readonly("ctx", &Decomposition::ctx)
      .def("accumulators", [](const Decomposition& d) { return VR(d.accumulators); })
      .def("dump", &Decomposition::Dump);
  m.def("sync_free_decompose", &SyncFreeDecompose);
  py::class_<StageTransfer>(m, "StageTransfer")
      .def_property_readonly("value", [](const StageTransfer& t) { return std::make_pair(t.value.node, t.value.idx); })
      .def_readonly("from_stage", &StageTransfer::from_stage).def_readonly("to_stage", &StageTransfer::to_stage)
      .def_readonly("backward", &StageTransfer::backward).def_readonly("bytes", &StageTransfer::bytes);
  m.def("stage_decompose", [](const Graph& g, int stages, Decomposition& d) { return StageDecompose(g, stages, &d); });
  m.def("compile_task_dag", [](const Graph& g, const Decomposition& d, const std::vector<StageTransfer>& x, int micro, int spmd,
                               const HwProfile& hw) {
    PipelineSpec sp;
    TaskDAG dag = CompileTaskDAG(g, d, x, micro, spmd, hw, &sp);
    return py::make_tuple(dag, sp);
  });

  py::class_<EvalInput>(m, "EvalInput")
      .def(py::init<>())
      .def_readwrite("num_stages", &EvalInput::num_stages).def_readwrite("num_micro", &EvalInput::num_micro)
      .def_readwrite("spmd", &EvalInput::spmd).def_readwrite("spmd_comm_bytes", &EvalInput::spmd_comm_bytes)
      .def_readwrite("exposed_comm_fraction", &EvalInput::exposed_comm_fraction)
      .def_readwrite("stage_flops", &EvalInput::stage_flops).def_readwrite("cut_bytes", &EvalInput::cut_bytes)
      .def_readwrite("var_bytes", &EvalInput::var_bytes).def_readwrite("act_bytes", &EvalInput::act_bytes)
      .def_readwrite("rows_per_micro", &EvalInput::rows_per_micro);
  py::class_<EvalResult>(m, "EvalResult")
      .def_readonly("feasible", &EvalResult::feasible).def_readonly("total_duration", &EvalResult::total_duration)
      .def_readonly("compute_time", &EvalResult::compute_time).def_readonly("comm_time", &EvalResult::comm_time)
      .def_readonly("p2p_time", &EvalResult::p2p_time).def_readonly("gpu_efficiency", &EvalResult::gpu_efficiency)
      .def_readonly("coll_ratio", &EvalResult::coll_ratio).def_readonly("bubble_ratio", &EvalResult::bubble_ratio)
      .def_readonly("mem_bytes_per_device", &EvalResult::mem_bytes_per_device)
      .def("__repr__", &EvalResult::str);
  m.def("evaluate", &Evaluate);
  py::class_<DeviceSplitProposal>(m, "DeviceSplitProposal")
      .def_readonly("stages", &DeviceSplitProposal::stages).def_readonly("spmd", &DeviceSplitProposal::spmd)
      .def_readonly("micro", &DeviceSplitProposal::micro).def("__repr__", &DeviceSplitProposal::str);
  m.def("generate_split_proposals", &GenerateSplitProposals);
  py::class_<AutoParallelOptions>(m, "AutoParallelOptions")
      .def(py::init<>())
      .def_readwrite("num_devices", &AutoParallelOptions::num_devices).def_readwrite("mode", &AutoParallelOptions::mode)
      .def_readwrite("num_stages", &AutoParallelOptions::num_stages)
      .def_readwrite("num_micro_batches", &AutoParallelOptions::num_micro_batches)
      .def_readwrite("spmd", &AutoParallelOptions::spmd)
      .def_readwrite("unbalanced_ratio", &AutoParallelOptions::unbalanced_ratio)
      .def_readwrite("allow_pipeline", &AutoParallelOptions::allow_pipeline)
      .def_readwrite("spmd_rule_mode", &AutoParallelOptions::spmd_rule_mode)
      .def_readwrite("exposed_comm_fraction", &AutoParallelOptions::exposed_comm_fraction)
      .def_readwrite("hw", &AutoParallelOptions::hw);
  py::class_<ParallelPlan>(m, "ParallelPlan")
      .def_readonly("proposal", &ParallelPlan::proposal).def_readonly("graph", &ParallelPlan::graph)
      .def_readonly("eval", &ParallelPlan::eval).def_readonly("spmd_stats", &ParallelPlan::spmd_stats)
      .def_readonly("stage_plan", &ParallelPlan::stage_plan).def_readonly("sync_free", &ParallelPlan::sync_free)
      .def_readonly("candidates", &ParallelPlan::candidates).def_readonly("log", &ParallelPlan::log);
  m.def("auto_parallel", &AutoParallelRun);
}
