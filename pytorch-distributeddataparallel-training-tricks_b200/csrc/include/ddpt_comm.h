// Fused gradient all-reduce over NVLink peer / multicast memory -- host API.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace ddpt {

constexpr int kMaxRanks = 8;
constexpr int kCommMaxBlocks = 128;         // CTAs per all-reduce launch (upper bound)
constexpr int kCommChannels = 8;            // independent signal regions (one per in-flight bucket)
constexpr long long kOneShotMaxElems = 64 * 1024;   // buckets up to 256 KiB of fp32 take the one-shot (push) path

// Signal pad layout (uint32 words) per rank:
//   flags[channel][block][phase 0|1][peer]   written by the peers: (epoch << 1) | payload bit, monotonic
//   epoch[channel][block]                    this rank's launch counter for (channel, block)
// Epochs live in device memory and are advanced by the kernels themselves, so the protocol needs no host-side
// sequence numbers and is CUDA-graph replay safe; flags are never reset.
constexpr int kFlagWords = kCommChannels * kCommMaxBlocks * 2 * kMaxRanks;
constexpr int kSignalWords = kFlagWords + kCommChannels * kCommMaxBlocks;

// Scratch region layout (bytes) per rank (symmetric, mapped on every peer, optionally multicast):
//   push[channel][parity][source rank][kPushSlotBytes]   one-shot: every rank pushes its bucket to every peer
//   stage[stage_elems] bf16                              two-shot bf16 wire: arena-sized 16-bit staging
constexpr long long kPushSlotBytes = kOneShotMaxElems * 4;
constexpr long long kPushBytes = (long long)kCommChannels * 2 * kMaxRanks * kPushSlotBytes;   // 32 MiB

enum AllReduceAlgo { kAlgoAuto = 0, kAlgoOneShot = 1, kAlgoTwoShot = 2, kAlgoMultimem = 3 };
enum AllReduceWire { kWireF32 = 0, kWireBF16 = 1 };

struct AllReduceArgs {
  float* peer_bufs[kMaxRanks];            // every rank's mapping of the symmetric gradient arena (index = rank)
  uint32_t* peer_signals[kMaxRanks];      // every rank's signal pad
  unsigned char* peer_scratch[kMaxRanks]; // every rank's scratch region (nullptr: one-shot / bf16 wire unavailable)
  float* mc_buf;                          // NVLS multicast mapping of the arena, or nullptr
  unsigned char* mc_scratch;              // NVLS multicast mapping of the scratch region, or nullptr
  int rank, world;
  long long offset, numel;                // bucket slice, in elements
  const float* scaler_state;              // optional: multiply by state[kInvScale]; always multiply by 1/world
  float* found_inf;                       // optional: raised (on every rank) if the reduced bucket has inf/nan
  int channel;
  int algo;
  int wire;                               // kWireF32: fp32 on the wire; kWireBF16: bf16 on the wire, fp32 accumulate
  int blocks;                             // 0 = default
  int timeout_ms;                         // peers missing for longer than this trap the kernel (0 = 30 s)
};

// out-of-place is never needed: the reduction happens in place in every rank's arena.
void launch_fused_allreduce(const AllReduceArgs& args, cudaStream_t stream);

// root -> all broadcast of a symmetric region (BatchNorm running statistics; parameters, N2)
void launch_symm_broadcast(const AllReduceArgs& args, int root, cudaStream_t stream);

}  // namespace ddpt
