// Shared device/host helpers for libpsb_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/psb.h"

namespace psb {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this
constexpr unsigned kFull = 0xffffffffu;

extern int64_t g_launches;  // host-side launch counter (psb_launch_count)

inline int check_table_args(const void* table, int64_t rows, int64_t d) {
  if (table == nullptr || rows <= 0) return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512) return PSB_E_DIM;
  if ((reinterpret_cast<uintptr_t>(table) & 15) != 0) return PSB_E_ALIGN;
  return PSB_OK;
}

// cudaFuncSetAttribute applies to the CURRENT device only, so "already configured" is remembered per device (a process
// may drive several GPUs: the single-process peer simulation, multi-device tests).
struct DeviceAttr {
  size_t bytes[64] = {};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool need(size_t want = 1) const { return want > bytes[dev()]; }
  void done(size_t want = 1) { bytes[dev()] = want; }
};

inline bool misaligned16(const void* p) { return p != nullptr && (reinterpret_cast<uintptr_t>(p) & 15) != 0; }

// Per-kernel timing (runtime.cu): when psb_profile_enable(1) is in effect, PSB_PROF records a CUDA event on the
// launching stream right before a kernel and launch_status() records one right after it; psb_profile_dump sums
// the elapsed times per kernel name.  Disabled (the default): one predictable branch per launch.
extern bool g_prof_on;
void prof_begin(const char* name, cudaStream_t s);
void prof_end();
#define PSB_PROF(name, stream) \
  do {                         \
    if (::psb::g_prof_on) ::psb::prof_begin(name, stream); \
  } while (0)

inline int launch_status() {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (g_prof_on) prof_end();
  return e == cudaSuccess ? PSB_OK : static_cast<int>(e);
}

// Library-owned side stream (+ fork event) of the current device, created on first use (runtime.cu): lets one ABI
// call run an independent tail of its work next to the caller's stream; capturable (fork = event record / wait).
int side_stream(cudaStream_t* stream, cudaEvent_t* fork_event, int slot = 0, cudaEvent_t* join_event = nullptr);

// Run `side_work(side stream)` next to `main_work()` on s: fork by event, join by event (both capturable).
template <typename FSide, typename FMain>
inline int fork_join(cudaStream_t s, int slot, FSide&& side_work, FMain&& main_work) {
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
  int st = side_stream(&s2, &fork_ev, slot, &join_ev);
  if (st != PSB_OK) return st;
  cudaError_t e = cudaEventRecord(fork_ev, s);
  if (e == cudaSuccess) e = cudaStreamWaitEvent(s2, fork_ev, 0);
  if (e != cudaSuccess) return static_cast<int>(e);
  if ((st = side_work(s2)) != PSB_OK) return st;
  if ((e = cudaEventRecord(join_ev, s2)) != cudaSuccess) return static_cast<int>(e);
  if ((st = main_work()) != PSB_OK) return st;
  if ((e = cudaStreamWaitEvent(s, join_ev, 0)) != cudaSuccess) return static_cast<int>(e);
  return PSB_OK;
}

// kernel<<<grid, block, smem, s>>>(args...) with the programmatic-stream-serialization attribute: the launch may begin
// before the previous kernel of the stream has finished (see pdl_trigger / pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int grid_for(int64_t work_items, int items_per_block, int max_waves = 16) {
  int64_t need = (work_items + items_per_block - 1) / items_per_block;
  int64_t cap = static_cast<int64_t>(kNumSMs) * max_waves;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

// 128-bit streaming load of table rows: read-only path, do not allocate in L1 (each
// gathered row is used once by the loading warp).
__device__ __forceinline__ float4 ldg_row4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

__device__ __forceinline__ void stg4(float4* p, const float4& v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)));
}

__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x);
  acc.y = fmaf(s, v.y, acc.y);
  acc.z = fmaf(s, v.z, acc.z);
  acc.w = fmaf(s, v.w, acc.w);
}

// Programmatic dependent launch (stream order K1 -> K2, K2 launched with launch_pdl): K1 lets the runtime start K2's CTAs
// while it is still running (pdl_trigger, first instruction), K2 does the part of its work that does not read K1's
// outputs and only then waits for K1's grid to have completed and flushed (pdl_wait).  Both are no-ops in a kernel
// launched the ordinary way.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }

}  // namespace psb
