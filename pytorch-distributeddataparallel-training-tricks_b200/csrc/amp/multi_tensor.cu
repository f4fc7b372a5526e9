// amp support kernels for sm_100a: multi-tensor scale / axpby with overflow detection, flat-arena
// unscale, and the device-side dynamic loss-scale update.
//
// Reference behaviour being replaced: Apex ``amp_C.multi_tensor_scale`` / ``multi_tensor_axpby``
// invoked from ``scale_loss.__exit__`` (reference utils/process.py:26-27) and
// the Python-side ``LossScaler.update_scale`` that reads the overflow flag back to the host.
//
// Design: one launch covers the whole tensor list.  The list travels in the kernel parameter
// space (no staging copy, CUDA-graph friendly).  Work is split in fixed 16 KiB chunks of the
// virtual concatenation of all tensors; a CTA finds its tensor by walking the (<=48 entry) chunk
// prefix held in registers/constant bank.  Loads are 128-bit when the tensor base is 16 B
// aligned, 4x unrolled so each thread keeps 64 B in flight (memory-level parallelism for HBM3e).
#include "../include/ddpt_amp.h"
#include "../include/ddpt_common.cuh"

namespace ddpt {
namespace {

constexpr int kChunkElems = 4096;   // 16 KiB of fp32 per (CTA, iteration)
constexpr int kThreads = 256;

struct ChunkMap {
  int count;
  int first_chunk[kMaxTensorsPerLaunch + 1];
};

template <typename Op>
__device__ __forceinline__ void apply_chunk(const float* __restrict__ a, const float* __restrict__ b,
                                            float* __restrict__ o, long long n, long long start, Op op,
                                            bool& bad) {
  const long long end = min(n, start + (long long)kChunkElems);
  const bool vec = (((uintptr_t)a | (uintptr_t)o | (uintptr_t)(b ? b : a)) & 15u) == 0;
  if (vec) {
    const long long v0 = start / 4, v1 = end / 4;   // start is a multiple of 4096
    for (long long i = v0 + threadIdx.x; i < v1; i += kThreads) {
      float4 x = ld_nc_f4(reinterpret_cast<const float4*>(a) + i);
      float4 y = b ? ld_nc_f4(reinterpret_cast<const float4*>(b) + i) : make_float4(0, 0, 0, 0);
      float4 r;
      r.x = op(x.x, y.x, bad); r.y = op(x.y, y.y, bad); r.z = op(x.z, y.z, bad); r.w = op(x.w, y.w, bad);
      reinterpret_cast<float4*>(o)[i] = r;
    }
    for (long long i = v1 * 4 + threadIdx.x; i < end; i += kThreads) o[i] = op(a[i], b ? b[i] : 0.f, bad);
  } else {
    for (long long i = start + threadIdx.x; i < end; i += kThreads) o[i] = op(a[i], b ? b[i] : 0.f, bad);
  }
}

template <typename Op>
__global__ void __launch_bounds__(kThreads)
multi_tensor_kernel(TensorList list, ChunkMap map, Op op, float* __restrict__ found_inf) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  const int total = map.first_chunk[map.count];
  bool bad = false;
  op.prepare();
  for (int chunk = blockIdx.x; chunk < total; chunk += gridDim.x) {
    int t = 0;
    while (t + 1 < map.count && map.first_chunk[t + 1] <= chunk) ++t;
    const long long start = (long long)(chunk - map.first_chunk[t]) * kChunkElems;
    apply_chunk(static_cast<const float*>(list.in0[t]), static_cast<const float*>(list.in1[t]),
                static_cast<float*>(list.out[t]), list.numel[t], start, op, bad);
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0 && found_inf) *found_inf = 1.0f;
}

struct ScaleOp {
  const float* scale_ptr;
  float s;
  __device__ void prepare() { s = __ldg(scale_ptr); }
  __device__ float operator()(float x, float, bool& bad) const {
    float r = x * s;
    bad |= !is_finite_f32(r);
    return r;
  }
};

struct AxpbyOp {
  float a, b;
  int check;  // -1: both, 0: x only, 1: y only
  __device__ void prepare() {}
  __device__ float operator()(float x, float y, bool& bad) const {
    if (check <= 0) bad |= !is_finite_f32(x);
    if (check != 0) bad |= !is_finite_f32(y);
    return a * x + b * y;
  }
};

ChunkMap build_map(const TensorList& list) {
  ChunkMap m;
  m.count = list.count;
  int acc = 0;
  for (int i = 0; i < list.count; ++i) {
    m.first_chunk[i] = acc;
    acc += (int)((list.numel[i] + kChunkElems - 1) / kChunkElems);
  }
  m.first_chunk[list.count] = acc;
  return m;
}

int grid_for(int chunks) {
  int cap = sm_count() * 8;
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}

__global__ void __launch_bounds__(kThreads)
unscale_arena_kernel(float* __restrict__ g, long long n, float* __restrict__ state) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  const float inv = state[kInvScale];
  bool bad = false;
  const long long nv = n / 4;
  float4* g4 = reinterpret_cast<float4*>(g);
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nv; i += (long long)gridDim.x * kThreads) {
    float4 v = g4[i];
    v.x *= inv; v.y *= inv; v.z *= inv; v.w *= inv;
    bad |= !(is_finite_f32(v.x) && is_finite_f32(v.y) && is_finite_f32(v.z) && is_finite_f32(v.w));
    g4[i] = v;
  }
  for (long long i = nv * 4 + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
    float v = g[i] * inv;
    bad |= !is_finite_f32(v);
    g[i] = v;
  }
  if (__syncthreads_or(bad) && threadIdx.x == 0) state[kFoundInf] = 1.0f;
}

__global__ void update_scale_kernel(float* __restrict__ s, int growth_interval, float min_scale, float max_scale) {
  pdl_wait();      // previous kernel of the stream has completed and flushed
  pdl_trigger();   // the next kernel may be scheduled behind this one
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float scale = s[kScale];
  if (s[kFoundInf] != 0.0f) {
    scale = fmaxf(scale * 0.5f, min_scale);
    s[kCleanSteps] = 0.0f;
    s[kSkipped] += 1.0f;
  } else {
    float clean = s[kCleanSteps] + 1.0f;
    if (clean >= (float)growth_interval) {
      scale = fminf(scale * 2.0f, max_scale);
      clean = 0.0f;
    }
    s[kCleanSteps] = clean;
  }
  s[kScale] = scale;
  s[kInvScale] = 1.0f / scale;   // power of two: exact
}

}  // namespace

void launch_multi_tensor_scale(const TensorList& list, const float* scale_ptr, float* found_inf,
                               cudaStream_t stream) {
  if (list.count == 0) return;
  ChunkMap map = build_map(list);
  ScaleOp op{scale_ptr, 1.0f};
  launch_pdl(multi_tensor_kernel<ScaleOp>, dim3(grid_for(map.first_chunk[map.count])), dim3(kThreads), 0, stream, list, map, op, found_inf);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_multi_tensor_axpby(const TensorList& list, float a, float b, int arg_to_check, float* found_inf,
                               cudaStream_t stream) {
  if (list.count == 0) return;
  ChunkMap map = build_map(list);
  AxpbyOp op{a, b, arg_to_check};
  launch_pdl(multi_tensor_kernel<AxpbyOp>, dim3(grid_for(map.first_chunk[map.count])), dim3(kThreads), 0, stream, list, map, op, found_inf);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_unscale_arena(float* grads, long long n, float* scaler_state, cudaStream_t stream) {
  if (n <= 0) return;
  long long want = (n / 4 + kThreads - 1) / kThreads;
  int grid = (int)(want < (long long)sm_count() * 8 ? (want > 0 ? want : 1) : sm_count() * 8);
  launch_pdl(unscale_arena_kernel, dim3(grid), dim3(kThreads), 0, stream, grads, n, scaler_state);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

void launch_update_scale(float* scaler_state, int growth_interval, float min_scale, float max_scale,
                         cudaStream_t stream) {
  launch_pdl(update_scale_kernel, dim3(1), dim3(32), 0, stream, scaler_state, growth_interval, min_scale, max_scale);
  DDPT_CUDA_CHECK(cudaGetLastError());
}

}  // namespace ddpt
