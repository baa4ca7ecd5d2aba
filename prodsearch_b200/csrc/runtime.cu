// Library-wide host state: launch counter and the per-kernel event profiler behind
// psb_profile_enable / psb_profile_dump (bench.py's roofline leg reads per-kernel durations from it;
// CUDA events on the stream each kernel is launched on, no external profiler).
#include <stdio.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "psb_common.cuh"

namespace psb {

int64_t g_launches = 0;
bool g_prof_on = false;

namespace {
struct Span {
  const char* name;
  cudaEvent_t a, b;
  bool closed;
};
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;
cudaStream_t g_open_stream = nullptr;
bool g_open = false;

cudaEvent_t take_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

void recycle() {
  for (Span& s : g_spans) {
    g_pool.push_back(s.a);
    g_pool.push_back(s.b);
  }
  g_spans.clear();
  g_open = false;
}
}  // namespace

int side_stream(cudaStream_t* stream, cudaEvent_t* fork_event, int slot, cudaEvent_t* join_event) {
  constexpr int kMaxDev = 64, kSlots = 3;
  static cudaStream_t streams[kMaxDev][kSlots] = {};
  static cudaEvent_t forks[kMaxDev][kSlots] = {};
  static cudaEvent_t joins[kMaxDev][kSlots] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (dev < 0 || dev >= kMaxDev || slot < 0 || slot >= kSlots) return PSB_E_UNSUPPORTED;
  if (streams[dev][slot] == nullptr) {
    e = cudaStreamCreateWithFlags(&streams[dev][slot], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&forks[dev][slot], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&joins[dev][slot], cudaEventDisableTiming);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  *stream = streams[dev][slot];
  *fork_event = forks[dev][slot];
  if (join_event != nullptr) *join_event = joins[dev][slot];
  return PSB_OK;
}

void prof_begin(const char* name, cudaStream_t s) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return;  // not in graphs
  if (g_spans.size() >= (1u << 20)) return;
  Span sp{name, take_event(), take_event(), false};
  cudaEventRecord(sp.a, s);
  g_spans.push_back(sp);
  g_open_stream = s;
  g_open = true;
}

void prof_end() {
  if (!g_open) return;
  Span& sp = g_spans.back();
  cudaEventRecord(sp.b, g_open_stream);
  sp.closed = true;
  g_open = false;
}

}  // namespace psb

using namespace psb;

extern "C" int64_t psb_launch_count(void) { return g_launches; }

extern "C" int psb_profile_enable(int32_t on) {
  recycle();
  g_prof_on = on != 0;
  return PSB_OK;
}

extern "C" int64_t psb_profile_dump(char* buf, int64_t cap) {
  if (buf == nullptr || cap <= 0) return PSB_E_ARG;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return -static_cast<int64_t>(e) - 1000;
  struct Acc {
    int64_t n = 0;
    double ms = 0.0, mn = 1e30, mx = 0.0;
  };
  std::map<std::string, Acc> acc;
  std::vector<std::string> order;
  for (const Span& s : g_spans) {
    if (!s.closed) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s.a, s.b) != cudaSuccess) continue;
    if (acc.find(s.name) == acc.end()) order.push_back(s.name);
    Acc& a = acc[s.name];
    a.n += 1;
    a.ms += ms;
    a.mn = ms < a.mn ? ms : a.mn;
    a.mx = ms > a.mx ? ms : a.mx;
  }
  int64_t off = 0;
  for (const std::string& name : order) {
    const Acc& a = acc[name];
    char line[256];
    const int len = snprintf(line, sizeof(line), "%s %lld %.6f %.6f %.6f\n", name.c_str(),
                             static_cast<long long>(a.n), a.ms, a.mn, a.mx);
    if (off + len + 1 > cap) break;
    memcpy(buf + off, line, len);
    off += len;
  }
  buf[off] = 0;
  recycle();
  return off;
}
