// Binding groups added incrementally (elementwise NN kernels, comm, reducer, host loader).
#include <torch/extension.h>
#include "bindings_decl.h"

void ddpt_bind_nn(pybind11::module_& m);
void ddpt_bind_comm(pybind11::module_& m);
void ddpt_bind_loader(pybind11::module_& m);
void ddpt_bind_reducer(pybind11::module_& m);

void ddpt_bind_extra(pybind11::module_& m) {
#ifdef DDPT_HAVE_NN
  ddpt_bind_nn(m);
#endif
#ifdef DDPT_HAVE_COMM
  ddpt_bind_comm(m);
#endif
#ifdef DDPT_HAVE_LOADER
  ddpt_bind_loader(m);
#endif
#ifdef DDPT_HAVE_REDUCER
  ddpt_bind_reducer(m);
#endif
}
