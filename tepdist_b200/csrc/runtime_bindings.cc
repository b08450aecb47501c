#include <pybind11/pybind11.h>
#include <pybind11/stl.h>
namespace py = pybind11;
void BindRuntime(py::module_& m) { (void)m; }
