// Native host-side batch loader (replaces the reference's per-sample PIL -> ToTensor -> collate path,
// reference utils/train.py:25-30 via torch DataLoader).
//
// The dataset is one dense uint8 matrix [N, row_bytes] plus int64 labels.  For every epoch the caller
// hands over the DistributedSampler-parity index list; worker threads gather whole batches straight
// into rotating *pinned* host buffers (allocated by the caller with cudaHostAlloc via torch), several
// batches ahead of the consumer.  The consumer (Python) issues one async H2D copy per batch from the
// slot and releases the slot once that copy's CUDA event has completed, so host work overlaps the GPU
// step completely and no per-sample Python executes.
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <thread>
#include <vector>

namespace py = pybind11;

namespace {

class PinnedGatherLoader {
 public:
  PinnedGatherLoader(uintptr_t images, uintptr_t labels, int64_t num_rows, int64_t row_bytes, int64_t batch,
                     std::vector<uintptr_t> image_slots, std::vector<uintptr_t> label_slots, int threads)
      : images_(reinterpret_cast<const uint8_t*>(images)), labels_(reinterpret_cast<const int64_t*>(labels)),
        num_rows_(num_rows), row_bytes_(row_bytes), batch_(batch) {
    if (image_slots.size() != label_slots.size() || image_slots.size() < 2)
      throw std::invalid_argument("need >= 2 pinned slots");
    for (size_t i = 0; i < image_slots.size(); ++i)
      slots_.push_back(Slot{reinterpret_cast<uint8_t*>(image_slots[i]), reinterpret_cast<int64_t*>(label_slots[i])});
    if (threads < 1) threads = 1;
    if (threads > (int)slots_.size()) threads = (int)slots_.size();
    for (int t = 0; t < threads; ++t) workers_.emplace_back([this, t, threads] { work(t, threads); });
  }

  ~PinnedGatherLoader() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
      ++generation_;
    }
    cv_.notify_all();
    for (auto& w : workers_) w.join();
  }

  // Begin an epoch: `n` int64 sample indices at `indices_ptr` are copied (no Python list, GIL released while the
  // previous epoch drains); batches are [b*batch, min((b+1)*batch, n)).
  void start_epoch(uintptr_t indices_ptr, int64_t n, int64_t num_batches) {
    const int64_t* src = reinterpret_cast<const int64_t*>(indices_ptr);
    for (int64_t i = 0; i < n; ++i)
      if (src[i] < 0 || src[i] >= num_rows_) throw std::out_of_range("sample index out of range");
    py::gil_scoped_release nogil;
    std::unique_lock<std::mutex> lk(mu_);
    // a previous (possibly speculative, never consumed) epoch may still have workers parked on "slot free": tell
    // them to drop it, then wait until every worker has left the batch loop
    abort_ = true;
    cv_.notify_all();
    cv_.wait(lk, [this] { return active_workers_ == 0; });
    abort_ = false;
    indices_.assign(src, src + n);
    num_batches_ = num_batches;
    for (auto& s : slots_) { s.batch = -1; s.ready = false; s.free = true; }
    next_consume_ = 0;
    ++generation_;
    active_workers_ = (int)workers_.size();
    cv_.notify_all();
  }

  // Blocks (GIL released) until batch `next` is ready; returns (slot, rows) or (-1, 0) at the end of the epoch.
  std::pair<int, int64_t> next() {
    py::gil_scoped_release nogil;
    std::unique_lock<std::mutex> lk(mu_);
    if (next_consume_ >= num_batches_) return {-1, 0};
    const int slot = (int)(next_consume_ % (int64_t)slots_.size());
    const int64_t want = next_consume_;
    cv_.wait(lk, [&] { return slots_[slot].ready && slots_[slot].batch == want; });
    ++next_consume_;
    return {slot, slots_[slot].rows};
  }

  // The consumer's H2D copy out of `slot` has completed: the slot may be refilled.
  void release(int slot) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      slots_.at(slot).ready = false;
      slots_.at(slot).free = true;
    }
    cv_.notify_all();
  }

  int64_t batches_gathered() const { return gathered_.load(); }

 private:
  struct Slot {
    uint8_t* images;
    int64_t* labels;
    int64_t batch = -1;
    int64_t rows = 0;
    bool ready = false;
    bool free = true;
  };

  void work(int tid, int nthreads) {
    uint64_t seen_generation = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen_generation; });
        if (stop_) return;
        seen_generation = generation_;
      }
      const int64_t total = (int64_t)indices_.size();
      for (int64_t b = tid; b < num_batches_; b += nthreads) {
        const int slot = (int)(b % (int64_t)slots_.size());
        {
          std::unique_lock<std::mutex> lk(mu_);
          // a slot is reusable once the consumer released the batch that used it `slots` batches ago
          cv_.wait(lk, [&] { return stop_ || abort_ || (slots_[slot].free && next_consume_ + (int64_t)slots_.size() > b); });
          if (stop_) return;
          if (abort_) break;
          slots_[slot].free = false;
        }
        const int64_t begin = b * batch_, end = std::min(total, begin + batch_);
        Slot& s = slots_[slot];
        for (int64_t i = begin; i < end; ++i) {
          const int64_t src = indices_[i];
          std::memcpy(s.images + (i - begin) * row_bytes_, images_ + src * row_bytes_, (size_t)row_bytes_);
          s.labels[i - begin] = labels_[src];
        }
        {
          std::lock_guard<std::mutex> lk(mu_);
          s.batch = b;
          s.rows = end - begin;
          s.ready = true;
        }
        gathered_.fetch_add(1);
        cv_.notify_all();
      }
      {
        std::lock_guard<std::mutex> lk(mu_);
        --active_workers_;
      }
      cv_.notify_all();
    }
  }

  const uint8_t* images_;
  const int64_t* labels_;
  int64_t num_rows_, row_bytes_, batch_;
  std::vector<Slot> slots_;
  std::vector<std::thread> workers_;
  std::vector<int64_t> indices_;
  int64_t num_batches_ = 0, next_consume_ = 0;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t generation_ = 0;
  int active_workers_ = 0;
  bool stop_ = false;
  bool abort_ = false;       // start_epoch: abandon the epoch in progress
  std::atomic<int64_t> gathered_{0};
};

}  // namespace

void ddpt_bind_loader(py::module_& m) {
  py::class_<PinnedGatherLoader>(m, "PinnedGatherLoader")
      .def(py::init<uintptr_t, uintptr_t, int64_t, int64_t, int64_t, std::vector<uintptr_t>, std::vector<uintptr_t>, int>())
      .def("start_epoch", &PinnedGatherLoader::start_epoch)
      .def("next", &PinnedGatherLoader::next)
      .def("release", &PinnedGatherLoader::release)
      .def("batches_gathered", &PinnedGatherLoader::batches_gathered);
}
