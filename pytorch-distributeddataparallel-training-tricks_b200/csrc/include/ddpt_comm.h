// Fused gradient all-reduce over NVLink peer / multicast memory -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ddpt {

constexpr int kMaxRanks = 8;
constexpr int kCommMaxBlocks = 32;          // CTAs per all-reduce launch (upper bound)
constexpr int kCommChannels = 8;            // independent signal regions (one per in-flight bucket)
// Signal pad layout (uint32 words) per rank:  [channel][block][peer] barrier flags, then [channel] overflow word.
constexpr int kSignalWords = kCommChannels * kCommMaxBlocks * kMaxRanks + kCommChannels;

enum AllReduceAlgo { kAlgoAuto = 0, kAlgoOneShot = 1, kAlgoTwoShot = 2, kAlgoMultimem = 3 };

struct AllReduceArgs {
  float* peer_bufs[kMaxRanks];        // every rank's mapping of the symmetric gradient arena (index = rank)
  uint32_t* peer_signals[kMaxRanks];  // every rank's signal pad
  float* mc_buf;                      // NVLS multicast mapping of the arena, or nullptr
  int rank, world;
  long long offset, numel;            // bucket slice, in elements (numel % (4 * world) == 0)
  const float* scaler_state;          // optional: multiply by state[kInvScale]; always multiply by 1/world
  float* found_inf;                   // optional: raised (on every rank) if the reduced bucket has inf/nan
  int channel;
  int algo;
  int blocks;                         // 0 = default
  int timeout_ms;                     // peers missing for longer than this trap the kernel (0 = 30 s)
};

// out-of-place is never needed: the reduction happens in place in every rank's arena.
void launch_fused_allreduce(const AllReduceArgs& args, cudaStream_t stream);

// rank-0 -> all broadcast of a small symmetric region (BatchNorm running statistics)
void launch_symm_broadcast(const AllReduceArgs& args, int root, cudaStream_t stream);

}  // namespace ddpt
