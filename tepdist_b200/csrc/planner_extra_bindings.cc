#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
namespace py = pybind11;
void BindPlannerExtra(py::module_& m) { (void)m; }
