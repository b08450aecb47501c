#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "spmd_planner.h"
#include "transform.h"

namespace py = pybind11;
using namespace tepdist;

void BindPlannerExtra(py::module_& m);

void BindPlanner(py::module_& m) {
  py::class_<TransformStats>(m, "TransformStats")
      .def(py::init<>())
      .def_readonly("num_all_reduce", &TransformStats::num_all_reduce)
      .def_readonly("num_all_gather", &TransformStats::num_all_gather)
      .def_readonly("num_all_to_all", &TransformStats::num_all_to_all)
      .def_readonly("num_reduce_scatter", &TransformStats::num_reduce_scatter)
      .def_readonly("num_dynamic_slice", &TransformStats::num_dynamic_slice)
      .def_readonly("comm_bytes", &TransformStats::comm_bytes)
      .def("comm_info", &TransformStats::CommInfo);
  m.def("spmd_transform", [](const Graph& g, const SpmdPlan& plan, int level, int num) {
    TransformStats st;
    Graph out = SpmdTransform(g, plan, level, num, &st);
    return py::make_tuple(out, st);
  });
  m.def("plan_flat_buckets", &PlanFlatBuckets);
  m.def("combine_gradient_collectives", [](Graph& g, int64_t bucket_bytes, int max_per_bucket) {
    return CombineGradientCollectives(&g, bucket_bytes, max_per_bucket);
  }, py::arg("graph"), py::arg("bucket_bytes"), py::arg("max_per_bucket") = 1 << 30);
  m.def("liveness_optimize", [](Graph& g, int64_t min_bytes) { return LivenessOptimize(&g, min_bytes); }, py::arg("graph"),
        py::arg("min_bytes") = 1 << 20);
  BindPlannerExtra(m);
}
