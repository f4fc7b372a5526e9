// Native gradient reducer (the framework's counterpart of torch's C++ DDP Reducer).
//
// One post-hook per parameter is attached to the parameter's AccumulateGrad node, so everything below
// runs on the autograd engine thread without touching Python:
//   hook  -> fold the fresh gradient into its slice of the flat arena (only if autograd produced a
//            separate tensor), count down the bucket, and when a bucket is complete launch its
//            all-reduce -- strictly in bucket order so every rank issues the same sequence
//   end of backward (engine callback) -> join the communication stream / wait for the collectives
//
// Back ends:
//   fused  : our sm_100a kernel over NVLink peer/multicast memory on a side stream (event-gated fork,
//            event join); the kernel also applies 1/(world*loss_scale) and raises found_inf
//   pg     : c10d::ProcessGroup::allreduce on the bucket slice (gloo on CPU, NCCL as the baseline)
// The bucket plan is static (reverse definition order, see parallel/buckets.py), so there is no rebuild
// iteration and no bucket-index broadcast.
#include <torch/extension.h>
#include <torch/csrc/autograd/engine.h>
#include <torch/csrc/autograd/functions/accumulate_grad.h>
#include <torch/csrc/autograd/utils/lambda_post_hook.h>
#include <torch/csrc/distributed/c10d/ProcessGroup.hpp>
#include <torch/csrc/utils/pybind.h>

#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/CUDAEvent.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

#include "../include/ddpt_comm.h"

namespace {

struct BucketSpec {
  std::vector<int64_t> params;
  int64_t start = 0, numel = 0;
};

class NativeReducer : public std::enable_shared_from_this<NativeReducer> {
 public:
  NativeReducer(std::vector<at::Tensor> params, std::vector<at::Tensor> grad_views, at::Tensor grad_arena,
                std::vector<std::vector<int64_t>> bucket_params, std::vector<int64_t> bucket_starts,
                std::vector<int64_t> bucket_numels)
      : params_(std::move(params)), grad_views_(std::move(grad_views)), arena_(std::move(grad_arena)) {
    TORCH_CHECK(params_.size() == grad_views_.size(), "one arena view per parameter");
    TORCH_CHECK(bucket_params.size() == bucket_starts.size() && bucket_params.size() == bucket_numels.size());
    bucket_of_.assign(params_.size(), -1);
    for (size_t b = 0; b < bucket_params.size(); ++b) {
      BucketSpec s;
      s.params = bucket_params[b];
      s.start = bucket_starts[b];
      s.numel = bucket_numels[b];
      for (auto p : s.params) {
        TORCH_CHECK(p >= 0 && p < (int64_t)params_.size() && bucket_of_[p] < 0, "bad bucket plan");
        bucket_of_[p] = (int)b;
      }
      buckets_.push_back(std::move(s));
    }
    pending_.assign(buckets_.size(), 0);
    ready_.assign(buckets_.size(), 0);
  }

  // Hooks capture a weak reference: a reducer that was dropped simply stops reacting.
  void install_hooks() {
    std::weak_ptr<NativeReducer> weak = shared_from_this();
    for (size_t i = 0; i < params_.size(); ++i) {
      if (!params_[i].requires_grad()) continue;
      auto acc = torch::autograd::impl::grad_accumulator(params_[i]);
      const int64_t idx = (int64_t)i;
      acc->add_post_hook(std::make_unique<torch::autograd::utils::LambdaPostHook>(
          [weak, idx](const torch::autograd::variable_list& outputs, const torch::autograd::variable_list&) {
            if (auto self = weak.lock()) self->on_grad_ready(idx);
            return outputs;
          }));
      accumulators_.push_back(std::move(acc));
    }
  }

  // ---- back ends -------------------------------------------------------------------------------------
  void use_process_group(c10::intrusive_ptr<c10d::ProcessGroup> pg) {
    pg_ = std::move(pg);
    world_ = pg_->getSize();
    fused_ = false;
  }

  void use_fused(std::vector<int64_t> peer_ptrs, std::vector<int64_t> signal_ptrs, int64_t mc_ptr,
                 std::vector<int64_t> scratch_ptrs, int64_t mc_scratch_ptr, int64_t rank, int64_t algo, int64_t wire,
                 int64_t blocks, at::Tensor unit_state) {
    world_ = (int)peer_ptrs.size();
    TORCH_CHECK(world_ >= 1 && world_ <= ddpt::kMaxRanks && signal_ptrs.size() == peer_ptrs.size());
    TORCH_CHECK(scratch_ptrs.empty() || scratch_ptrs.size() == peer_ptrs.size());
    TORCH_CHECK(arena_.is_cuda(), "fused back end needs a CUDA arena");
    base_ = ddpt::AllReduceArgs{};
    for (int r = 0; r < world_; ++r) {
      base_.peer_bufs[r] = reinterpret_cast<float*>(peer_ptrs[r]);
      base_.peer_signals[r] = reinterpret_cast<uint32_t*>(signal_ptrs[r]);
      base_.peer_scratch[r] = scratch_ptrs.empty() ? nullptr : reinterpret_cast<unsigned char*>(scratch_ptrs[r]);
    }
    base_.mc_buf = reinterpret_cast<float*>(mc_ptr);
    base_.mc_scratch = reinterpret_cast<unsigned char*>(mc_scratch_ptr);
    base_.rank = (int)rank;
    base_.world = world_;
    base_.algo = (int)algo;
    base_.wire = (int)wire;
    base_.blocks = (int)blocks;
    unit_state_ = std::move(unit_state);
    comm_stream_ = c10::cuda::getStreamFromPool(/*isHighPriority=*/true, arena_.device().index());
    fused_ = true;
  }

  // ---- amp integration: the fused kernel unscales and checks for overflow itself -------------------------
  void expect_scaled_grads(c10::optional<at::Tensor> scaler_state) {
    scaler_state_ = std::move(scaler_state);
    scaler_consumed_ = false;
  }
  bool consumed_scaler() const { return scaler_consumed_; }

  // ---- per-iteration ----------------------------------------------------------------------------------------
  void prepare_for_backward() {
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t b = 0; b < buckets_.size(); ++b) {
      int n = 0;
      for (auto p : buckets_[b].params) n += params_[p].requires_grad() ? 1 : 0;
      pending_[b] = n;
      ready_[b] = n == 0 ? 1 : 0;
    }
    next_to_launch_ = 0;
    armed_ = true;
    callback_queued_ = false;
    works_.clear();
  }

  void set_require_sync(bool v) { require_sync_ = v; }
  bool require_sync() const { return require_sync_; }
  int64_t launches() const { return launches_; }
  int64_t steps() const { return steps_; }
  int64_t folded_copies() const { return folded_; }

 private:
  void on_grad_ready(int64_t idx) {
    std::lock_guard<std::mutex> lk(mu_);
    if (!armed_ || !require_sync_) return;
    at::Tensor& grad = params_[idx].mutable_grad();
    const at::Tensor& view = grad_views_[idx];
    if (grad.defined() && grad.data_ptr() != view.data_ptr()) {
      at::NoGradGuard ng;
      const_cast<at::Tensor&>(view).copy_(grad);          // fold the fresh tensor into the arena
      grad = view;
      ++folded_;
    }
    if (!callback_queued_) {
      callback_queued_ = true;
      std::weak_ptr<NativeReducer> weak = shared_from_this();
      torch::autograd::Engine::get_default_engine().queue_callback([weak] {
        if (auto self = weak.lock()) self->finalize();
      });
    }
    const int b = bucket_of_[idx];
    if (--pending_[b] == 0) {
      ready_[b] = 1;
      while (next_to_launch_ < (int)buckets_.size() && ready_[next_to_launch_]) launch(next_to_launch_++);
    }
  }

  void launch(int b) {
    const BucketSpec& spec = buckets_[b];
    if (fused_) {
      const auto dev = arena_.device().index();
      c10::cuda::CUDAGuard guard(dev);
      auto main = c10::cuda::getCurrentCUDAStream(dev);
      at::cuda::CUDAEvent ev;
      ev.record(main);
      ev.block(*comm_stream_);
      ddpt::AllReduceArgs a = base_;
      a.offset = spec.start;
      a.numel = spec.numel;
      a.channel = b % (ddpt::kCommChannels - 1);          // the last channel is the buffer broadcast's
      at::Tensor state = scaler_state_.has_value() ? *scaler_state_ : unit_state_;
      a.scaler_state = state.data_ptr<float>();
      a.found_inf = state.data_ptr<float>() + 2;
      ddpt::launch_fused_allreduce(a, comm_stream_->stream());
      if (scaler_state_.has_value()) scaler_consumed_ = true;
      comm_used_ = true;
    } else {
      at::NoGradGuard ng;
      at::Tensor flat = arena_.narrow(0, spec.start, spec.numel);
      flat.div_((double)world_);
      std::vector<at::Tensor> tensors{flat};
      works_.push_back(pg_->allreduce(tensors));
    }
    ++launches_;
  }

  void finalize() {
    std::lock_guard<std::mutex> lk(mu_);
    if (!armed_) return;
    armed_ = false;
    if (next_to_launch_ != (int)buckets_.size()) {
      TORCH_CHECK(false, "backward finished but gradient bucket ", next_to_launch_,
                  " never became ready: some parameters did not receive gradients (unused parameters are not "
                  "supported; freeze them with requires_grad=False before wrapping)");
    }
    if (fused_) {
      if (comm_used_) {
        const auto dev = arena_.device().index();
        at::cuda::CUDAEvent ev;
        ev.record(*comm_stream_);
        ev.block(c10::cuda::getCurrentCUDAStream(dev));
        comm_used_ = false;
      }
    } else {
      for (auto& w : works_) w->wait();
      works_.clear();
    }
    ++steps_;
  }

  std::vector<at::Tensor> params_, grad_views_;
  at::Tensor arena_;
  std::vector<BucketSpec> buckets_;
  std::vector<int> bucket_of_;
  std::vector<int> pending_;
  std::vector<char> ready_;
  std::vector<std::shared_ptr<torch::autograd::Node>> accumulators_;
  int next_to_launch_ = 0;
  bool armed_ = false, callback_queued_ = false, require_sync_ = true;
  std::mutex mu_;

  bool fused_ = false, comm_used_ = false, scaler_consumed_ = false;
  int world_ = 1;
  c10::intrusive_ptr<c10d::ProcessGroup> pg_;
  std::vector<c10::intrusive_ptr<c10d::Work>> works_;
  ddpt::AllReduceArgs base_{};
  at::Tensor unit_state_;
  c10::optional<at::Tensor> scaler_state_;
  c10::optional<c10::cuda::CUDAStream> comm_stream_;
  int64_t launches_ = 0, steps_ = 0, folded_ = 0;
};

}  // namespace

void ddpt_bind_reducer(pybind11::module_& m) {
  py::class_<NativeReducer, std::shared_ptr<NativeReducer>>(m, "NativeReducer")
      .def(py::init([](std::vector<at::Tensor> params, std::vector<at::Tensor> grad_views, at::Tensor arena,
                       std::vector<std::vector<int64_t>> bucket_params, std::vector<int64_t> starts,
                       std::vector<int64_t> numels) {
        auto r = std::make_shared<NativeReducer>(std::move(params), std::move(grad_views), std::move(arena),
                                                 std::move(bucket_params), std::move(starts), std::move(numels));
        r->install_hooks();
        return r;
      }))
      .def("use_process_group", &NativeReducer::use_process_group)
      .def("use_fused", &NativeReducer::use_fused)
      .def("expect_scaled_grads", &NativeReducer::expect_scaled_grads)
      .def("consumed_scaler", &NativeReducer::consumed_scaler)
      .def("prepare_for_backward", &NativeReducer::prepare_for_backward)
      .def_property("require_sync", &NativeReducer::require_sync, &NativeReducer::set_require_sync)
      .def_property_readonly("launches", &NativeReducer::launches)
      .def_property_readonly("steps", &NativeReducer::steps)
      .def_property_readonly("folded_copies", &NativeReducer::folded_copies);
}
