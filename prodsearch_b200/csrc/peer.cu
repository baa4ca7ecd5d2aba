// Row-sharded tables over NVLink peer memory: one process per GPU, every rank maps
// every other rank's shard / staging buffers into its own address space (CUDA IPC over NVLink P2P) and the
// exchange steps are plain loads inside the kernels -- no all-to-all, no host synchronisation, so the
// whole multi-GPU training step replays as one CUDA graph per rank.
//
//   peer_gather_rows   out[i] = shard[id % G][id / G]      (forward "fetch": P2P 128-bit row loads)
//   peer_fold_rows     owner-side gradient fold: dense[id / G] += scale * vals[i] for the slots of ONE
//                      peer's compact (sorted unique rows, values) list that this rank owns; launched once
//                      per peer in rank order => deterministic, no float atomics
//   peer_allreduce     one-shot sum of the replicated dense gradients: every rank reads all G buffers and
//                      adds them in rank order (bit-identical result on every rank)
//   peer_barrier       cross-GPU barrier on flag words in peer memory (release/acquire at system scope),
//                      with a clock64 time-out that raises a device error flag instead of hanging
#include <string.h>

#include "psb_common.cuh"

namespace psb {

constexpr int kMaxPeers = 16;

struct PeerPtrs {
  const void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// flags.p[r] = rank r's flag block (uint32[kMaxPeers], in r's peer memory).  Thread t signals peer t by writing
// the new epoch into flags[t][rank] and waits until flags[rank][t] reaches it.
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerPtrs flags, int rank, int G, uint32_t* __restrict__ epoch,
                                                          int32_t* __restrict__ err, long long timeout_cycles) {
  __shared__ uint32_t e_sh;
  const int t = threadIdx.x;
  if (t == 0) {
    e_sh = *epoch + 1u;
    *epoch = e_sh;
  }
  __syncthreads();
  const uint32_t e = e_sh;
  if (t < G) {
    __threadfence_system();
    uint32_t* theirs = static_cast<uint32_t*>(const_cast<void*>(flags.p[t])) + rank;
    st_release_sys(theirs, e);
    const uint32_t* mine = static_cast<const uint32_t*>(flags.p[rank]) + t;
    const long long t0 = clock64();
    while (static_cast<int32_t>(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > timeout_cycles) {
        if (err != nullptr) *err = 1 + t;
        break;
      }
      __nanosleep(64);
    }
    __threadfence_system();
  }
}

template <int R>
__global__ void __launch_bounds__(256)
peer_gather_rows_kernel(PeerPtrs shards, int G, int64_t rows_total, int d4, const int64_t* __restrict__ idx, int64_t n,
                        float4* __restrict__ out, int64_t* __restrict__ remap, int64_t pad_id, int64_t pad_pos,
                        int32_t* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = warp * R; base < n; base += nwarps * R) {
    const float4* src[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int64_t row = base + i;
      src[i] = nullptr;
      if (row < n) {
        const int64_t v = idx[row];
        if (remap != nullptr && lane == 0) remap[row] = v == pad_id ? pad_pos : row;
        if (v < 0 || v >= rows_total) {
          if (err != nullptr && lane == 0) *err = 1;
        } else {
          src[i] = static_cast<const float4*>(shards.p[v % G]) + (v / G) * d4;
        }
      }
    }
    for (int c = lane; c < d4; c += 32) {
      float4 v[R];
#pragma unroll
      for (int i = 0; i < R; ++i) v[i] = src[i] != nullptr ? ldg_row4(src[i] + c) : zero4();
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (base + i < n) stg4(out + (base + i) * d4 + c, v[i]);
    }
  }
}

// One peer's compact list: rows[0..*n_rows) ascending global ids (each at most once), vals[i,:].  A warp per slot.
__global__ void __launch_bounds__(256)
peer_fold_rows_kernel(const int32_t* __restrict__ rows, const float4* __restrict__ vals, const float* __restrict__ bias_vals,
                      const int32_t* __restrict__ n_rows, int64_t cap, int rank, int G, int d4, float scale,
                      float4* __restrict__ dense, float* __restrict__ dense_bias, int64_t shard_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  int64_t n = *n_rows;
  n = n < cap ? n : cap;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
    const int32_t g = rows[i];
    if (g < 0 || g % G != rank) continue;
    const int64_t local = g / G;
    if (local >= shard_rows) continue;
    if (dense != nullptr) {
      for (int c = lane; c < d4; c += 32) {
        const float4 v = ldg_row4(vals + i * d4 + c);
        float4 a = dense[local * d4 + c];
        fma4(a, scale, v);
        dense[local * d4 + c] = a;
      }
    }
    if (dense_bias != nullptr && bias_vals != nullptr && lane == 0)
      dense_bias[local] = fmaf(scale, bias_vals[i], dense_bias[local]);
  }
}

__global__ void __launch_bounds__(256)
peer_allreduce_kernel(PeerPtrs bufs, int G, int64_t n, float scale, float* __restrict__ out) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 acc = zero4();
    for (int r = 0; r < G; ++r) {
      const float4 v = ldg_row4(static_cast<const float4*>(bufs.p[r]) + i);
      acc.x += v.x;
      acc.y += v.y;
      acc.z += v.z;
      acc.w += v.w;
    }
    acc.x *= scale;
    acc.y *= scale;
    acc.z *= scale;
    acc.w *= scale;
    reinterpret_cast<float4*>(out)[i] = acc;
  }
  for (int64_t i = (n4 << 2) + static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float acc = 0.f;
    for (int r = 0; r < G; ++r) acc += static_cast<const float*>(bufs.p[r])[i];
    out[i] = acc * scale;
  }
}

}  // namespace psb

using namespace psb;

extern "C" int psb_peer_alloc(int64_t bytes, void** out) {
  if (bytes <= 0 || out == nullptr) return PSB_E_ARG;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  e = cudaMemset(p, 0, static_cast<size_t>(bytes));
  if (e != cudaSuccess) return static_cast<int>(e);
  *out = p;
  return PSB_OK;
}

extern "C" int psb_peer_free(void* p) {
  if (p == nullptr) return PSB_OK;
  cudaError_t e = cudaFree(p);
  return e == cudaSuccess ? PSB_OK : static_cast<int>(e);
}

extern "C" int psb_peer_export(void* p, void* handle) {
  static_assert(sizeof(cudaIpcMemHandle_t) == PSB_PEER_HANDLE_BYTES, "handle size");
  if (p == nullptr || handle == nullptr) return PSB_E_ARG;
  cudaError_t e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle), p);
  return e == cudaSuccess ? PSB_OK : static_cast<int>(e);
}

extern "C" int psb_peer_open(const void* handle, void** out) {
  if (handle == nullptr || out == nullptr) return PSB_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return static_cast<int>(e);
  *out = p;
  return PSB_OK;
}

extern "C" int psb_peer_close(void* p) {
  if (p == nullptr) return PSB_OK;
  cudaError_t e = cudaIpcCloseMemHandle(p);
  return e == cudaSuccess ? PSB_OK : static_cast<int>(e);
}

static int fill_ptrs(PeerPtrs* P, const void* const* host_ptrs, int32_t G) {
  if (host_ptrs == nullptr || G <= 0 || G > kMaxPeers) return PSB_E_ARG;
  for (int i = 0; i < kMaxPeers; ++i) P->p[i] = i < G ? host_ptrs[i] : nullptr;
  for (int i = 0; i < G; ++i)
    if (P->p[i] == nullptr) return PSB_E_ARG;
  return PSB_OK;
}

extern "C" int psb_peer_barrier(const void* const* flag_blocks, int32_t rank, int32_t G, uint32_t* epoch_dev,
                                int32_t* err_dev, int64_t timeout_cycles, psb_stream_t stream) {
  PeerPtrs F;
  int st = fill_ptrs(&F, flag_blocks, G);
  if (st != PSB_OK) return st;
  if (rank < 0 || rank >= G || epoch_dev == nullptr) return PSB_E_ARG;
  if (timeout_cycles <= 0) timeout_cycles = 4000000000ll;  // ~2 s at 1.9 GHz
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("peer_barrier_kernel", s);
  peer_barrier_kernel<<<1, 32, 0, s>>>(F, rank, G, epoch_dev, err_dev, timeout_cycles);
  return launch_status();
}

extern "C" int psb_peer_gather_rows(const void* const* shards, int32_t G, int64_t rows_total, int64_t d,
                                    const int64_t* idx, int64_t n, float* out, int64_t* remap_out, int64_t pad_id,
                                    int64_t pad_pos, int32_t* err_flag, psb_stream_t stream) {
  PeerPtrs S;
  int st = fill_ptrs(&S, shards, G);
  if (st != PSB_OK) return st;
  if (idx == nullptr || out == nullptr || n < 0 || rows_total <= 0) return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512) return PSB_E_DIM;
  for (int i = 0; i < G; ++i)
    if (misaligned16(S.p[i])) return PSB_E_ALIGN;
  if (misaligned16(out)) return PSB_E_ALIGN;
  if (n == 0) return PSB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int R = 8;
  PSB_PROF("peer_gather_rows_kernel", s);
  peer_gather_rows_kernel<R><<<grid_for(n, 8 * R), 256, 0, s>>>(S, G, rows_total, static_cast<int>(d / 4), idx, n,
                                                                reinterpret_cast<float4*>(out), remap_out, pad_id, pad_pos,
                                                                err_flag);
  return launch_status();
}

extern "C" int psb_peer_fold_rows(const int32_t* rows, const float* vals, const float* bias_vals, const int32_t* n_rows,
                                  int64_t cap, int32_t rank, int32_t G, int64_t d, float scale, float* dense,
                                  float* dense_bias, int64_t shard_rows, psb_stream_t stream) {
  if (rows == nullptr || n_rows == nullptr || cap < 0 || G <= 0 || rank < 0 || rank >= G || shard_rows <= 0)
    return PSB_E_ARG;
  if (dense != nullptr && vals == nullptr) return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512) return PSB_E_DIM;
  if (misaligned16(vals) || misaligned16(dense)) return PSB_E_ALIGN;
  if (cap == 0) return PSB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("peer_fold_rows_kernel", s);
  peer_fold_rows_kernel<<<grid_for(cap, 8, 8), 256, 0, s>>>(rows, reinterpret_cast<const float4*>(vals), bias_vals, n_rows,
                                                           cap, rank, G, static_cast<int>(d / 4), scale,
                                                           reinterpret_cast<float4*>(dense), dense_bias, shard_rows);
  return launch_status();
}

extern "C" int psb_peer_allreduce(const void* const* bufs, int32_t G, int64_t n, float scale, float* out,
                                  psb_stream_t stream) {
  PeerPtrs B;
  int st = fill_ptrs(&B, bufs, G);
  if (st != PSB_OK) return st;
  if (out == nullptr || n < 0) return PSB_E_ARG;
  for (int i = 0; i < G; ++i)
    if (misaligned16(B.p[i])) return PSB_E_ALIGN;
  if (misaligned16(out)) return PSB_E_ALIGN;
  if (n == 0) return PSB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("peer_allreduce_kernel", s);
  peer_allreduce_kernel<<<grid_for(n, 256 * 4, 4), 256, 0, s>>>(B, G, n, scale, out);
  return launch_status();
}
