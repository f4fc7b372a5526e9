// Bindings for the NVLink communication layer: fused all-reduce / broadcast kernels and the CUDA-IPC
// symmetric allocation fallback (used when torch's symmetric-memory rendezvous is unavailable).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda_runtime.h>

#include "include/ddpt_comm.h"

namespace {

cudaStream_t cur() { return at::cuda::getCurrentCUDAStream().stream(); }

void cuda_ok(cudaError_t e, const char* what) {
  TORCH_CHECK(e == cudaSuccess, "ddpt_b200 comm: ", what, ": ", cudaGetErrorString(e));
}

ddpt::AllReduceArgs make_args(const std::vector<int64_t>& peer_ptrs, const std::vector<int64_t>& signal_ptrs,
                              int64_t mc_ptr, const std::vector<int64_t>& scratch_ptrs, int64_t mc_scratch_ptr,
                              int64_t rank, int64_t offset, int64_t numel,
                              const c10::optional<at::Tensor>& scaler_state, const c10::optional<at::Tensor>& found_inf,
                              int64_t channel, int64_t algo, int64_t wire, int64_t blocks, int64_t timeout_ms = 0) {
  const int world = (int)peer_ptrs.size();
  TORCH_CHECK(world >= 1 && world <= ddpt::kMaxRanks && signal_ptrs.size() == peer_ptrs.size());
  TORCH_CHECK(scratch_ptrs.empty() || scratch_ptrs.size() == peer_ptrs.size(), "one scratch pointer per rank (or none)");
  TORCH_CHECK(rank >= 0 && rank < world, "bad rank");
  ddpt::AllReduceArgs a{};
  for (int r = 0; r < world; ++r) {
    a.peer_bufs[r] = reinterpret_cast<float*>(peer_ptrs[r]);
    a.peer_signals[r] = reinterpret_cast<uint32_t*>(signal_ptrs[r]);
    a.peer_scratch[r] = scratch_ptrs.empty() ? nullptr : reinterpret_cast<unsigned char*>(scratch_ptrs[r]);
  }
  a.mc_buf = reinterpret_cast<float*>(mc_ptr);
  a.mc_scratch = reinterpret_cast<unsigned char*>(mc_scratch_ptr);
  a.rank = (int)rank; a.world = world; a.offset = offset; a.numel = numel;
  if (scaler_state.has_value()) {
    TORCH_CHECK(scaler_state->is_cuda() && scaler_state->scalar_type() == at::kFloat && scaler_state->numel() >= 5);
    a.scaler_state = scaler_state->data_ptr<float>();
  }
  if (found_inf.has_value()) {
    TORCH_CHECK(found_inf->is_cuda() && found_inf->scalar_type() == at::kFloat);
    a.found_inf = found_inf->data_ptr<float>();
  }
  a.channel = (int)channel; a.algo = (int)algo; a.wire = (int)wire; a.blocks = (int)blocks; a.timeout_ms = (int)timeout_ms;
  return a;
}

void fused_allreduce(std::vector<int64_t> peer_ptrs, std::vector<int64_t> signal_ptrs, int64_t mc_ptr,
                     std::vector<int64_t> scratch_ptrs, int64_t mc_scratch_ptr, int64_t rank,
                     int64_t offset, int64_t numel, c10::optional<at::Tensor> scaler_state,
                     c10::optional<at::Tensor> found_inf, int64_t channel, int64_t algo, int64_t wire, int64_t blocks,
                     int64_t timeout_ms) {
  ddpt::launch_fused_allreduce(make_args(peer_ptrs, signal_ptrs, mc_ptr, scratch_ptrs, mc_scratch_ptr, rank, offset,
                                         numel, scaler_state, found_inf, channel, algo, wire, blocks, timeout_ms), cur());
}

// root's [offset, offset + numel) floats of a symmetric region replace every rank's copy (one launch)
void symm_broadcast(std::vector<int64_t> peer_ptrs, std::vector<int64_t> signal_ptrs, int64_t mc_ptr, int64_t rank,
                    int64_t root, int64_t offset, int64_t numel, int64_t channel, int64_t timeout_ms) {
  ddpt::launch_symm_broadcast(make_args(peer_ptrs, signal_ptrs, mc_ptr, {}, 0, rank, offset, numel, c10::nullopt,
                                        c10::nullopt, channel, 0, 0, 0, timeout_ms), (int)root, cur());
}

int64_t signal_pad_bytes() { return (int64_t)ddpt::kSignalWords * 4; }
// bytes of the scratch region for an arena of `stage_elems` elements: push slots + 16-bit staging
int64_t scratch_bytes(int64_t stage_elems) { return (int64_t)ddpt::kPushBytes + 2 * ((stage_elems + 127) / 128 * 128); }
int64_t oneshot_max_elems() { return (int64_t)ddpt::kOneShotMaxElems; }

// ---- CUDA IPC symmetric allocation (fallback) -----------------------------------------------------------
// Returns (device pointer, 64-byte IPC handle).  Memory comes from cudaMalloc (not the caching allocator)
// so that the handle maps the allocation's base.
std::pair<int64_t, py::bytes> ipc_alloc(int64_t nbytes, int64_t device) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  void* p = nullptr;
  cuda_ok(cudaMalloc(&p, (size_t)nbytes), "cudaMalloc");
  cuda_ok(cudaMemset(p, 0, (size_t)nbytes), "cudaMemset");
  cudaIpcMemHandle_t h;
  cuda_ok(cudaIpcGetMemHandle(&h, p), "cudaIpcGetMemHandle");
  return {reinterpret_cast<int64_t>(p), py::bytes(reinterpret_cast<const char*>(&h), sizeof(h))};
}

int64_t ipc_open(const std::string& handle, int64_t device) {
  TORCH_CHECK(handle.size() == sizeof(cudaIpcMemHandle_t), "bad IPC handle size");
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle.data(), sizeof(h));
  void* p = nullptr;
  cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
  return reinterpret_cast<int64_t>(p);
}

void ipc_close(int64_t ptr) { cudaIpcCloseMemHandle(reinterpret_cast<void*>(ptr)); }
void ipc_free(int64_t ptr) { cudaFree(reinterpret_cast<void*>(ptr)); }

// Enable peer access from the current device to `peer` (same-process multi-device simulation / tests).
bool enable_peer_access(int64_t device, int64_t peer) {
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  int can = 0;
  cuda_ok(cudaDeviceCanAccessPeer(&can, (int)device, (int)peer), "cudaDeviceCanAccessPeer");
  if (!can) return false;
  cudaError_t e = cudaDeviceEnablePeerAccess((int)peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return true; }
  return e == cudaSuccess;
}

at::Tensor tensor_from_ptr(int64_t ptr, int64_t numel, int64_t device, bool is_uint32) {
  auto opts = at::TensorOptions().device(at::kCUDA, (c10::DeviceIndex)device).dtype(is_uint32 ? at::kInt : at::kFloat);
  return at::from_blob(reinterpret_cast<void*>(ptr), {numel}, opts);
}

}  // namespace

void ddpt_bind_comm(pybind11::module_& m) {
  m.def("fused_allreduce", &fused_allreduce, py::arg("peer_ptrs"), py::arg("signal_ptrs"), py::arg("mc_ptr"),
        py::arg("scratch_ptrs"), py::arg("mc_scratch_ptr"), py::arg("rank"), py::arg("offset"), py::arg("numel"),
        py::arg("scaler_state"), py::arg("found_inf"), py::arg("channel"), py::arg("algo"), py::arg("wire") = 0,
        py::arg("blocks") = 0, py::arg("timeout_ms") = 0);
  m.def("symm_broadcast", &symm_broadcast, py::arg("peer_ptrs"), py::arg("signal_ptrs"), py::arg("mc_ptr"),
        py::arg("rank"), py::arg("root"), py::arg("offset"), py::arg("numel"), py::arg("channel"),
        py::arg("timeout_ms") = 0);
  m.def("scratch_bytes", &scratch_bytes);
  m.def("oneshot_max_elems", &oneshot_max_elems);
  m.def("signal_pad_bytes", &signal_pad_bytes);
  m.def("ipc_alloc", &ipc_alloc);
  m.def("ipc_open", &ipc_open);
  m.def("ipc_close", &ipc_close);
  m.def("ipc_free", &ipc_free);
  m.def("enable_peer_access", &enable_peer_access);
  m.def("tensor_from_ptr", &tensor_from_ptr);
}
