// Additional binding groups implemented in other translation units.
#pragma once
#include <pybind11/pybind11.h>
void ddpt_bind_extra(pybind11::module_& m);
