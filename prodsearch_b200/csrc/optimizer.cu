// Fused clipped Adam over ALL parameter tensors of the model in three launches
// (psb_adam_step; reference models/optimizers.py:205-243 + torch.optim.Adam):
//   sqnorm partials -> final (global L2 norm, step counter += 1) -> update.
// HBM-bound: per element 4 loads (p, g, m, v) + 3 stores = 28 B; the dense embedding tables dominate.
// No host synchronisation: the step counter, the norm and the clip factor live in device memory, so
// the whole optimizer step replays inside a CUDA graph.  Deterministic (fixed-order partial sums).
#include "psb_common.cuh"

namespace psb {

constexpr int kAdamChunk = 4096;  // floats per CTA (256 threads x 4 float4)

struct AdamTensors {
  psb_adam_tensor_t t[PSB_ADAM_MAX_TENSORS];
  int n;
};

__device__ __forceinline__ bool locate(const AdamTensors& T, int b, int* ti, int64_t* start) {
  for (int q = 0; q < T.n; ++q) {
    const int64_t chunks = (T.t[q].n + kAdamChunk - 1) / kAdamChunk;
    if (b < chunks) {
      *ti = q;
      *start = static_cast<int64_t>(b) * kAdamChunk;
      return true;
    }
    b -= static_cast<int>(chunks);
  }
  return false;
}

__global__ void __launch_bounds__(256) sqnorm_partial_kernel(const AdamTensors T, float* __restrict__ partial) {
  __shared__ float wsum[8];
  int ti;
  int64_t start;
  float acc = 0.f;
  if (locate(T, blockIdx.x, &ti, &start)) {
    const float* g = T.t[ti].g;
    const int64_t n = T.t[ti].n;
    const int64_t end = min(n, start + kAdamChunk);
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      for (int64_t i = start + threadIdx.x * 4; i < end; i += 1024) {
        if (i + 4 <= end) {
          const float4 v = *reinterpret_cast<const float4*>(g + i);
          acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        } else {
          for (int64_t j = i; j < end; ++j) acc = fmaf(g[j], g[j], acc);
        }
      }
    } else {
      for (int64_t i = start + threadIdx.x; i < end; i += 256) acc = fmaf(g[i], g[i], acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    partial[blockIdx.x] = s;
  }
}

// state[0] = step (as float-exact int64 in state_i), out: sqnorm
__global__ void __launch_bounds__(256) sqnorm_final_kernel(const float* __restrict__ partial, int n,
                                                           float* __restrict__ sqnorm, int64_t* __restrict__ step) {
  __shared__ double wsum[8];
  double acc = 0.0;
  const int per = (n + 255) / 256;
  const int lo = threadIdx.x * per, hi = min(n, lo + per);
  for (int i = lo; i < hi; ++i) acc += static_cast<double>(partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wsum[w];
    *sqnorm = static_cast<float>(s);
    if (step != nullptr) *step += 1;
  }
}

__global__ void bump_step_kernel(int64_t* __restrict__ step) { *step += 1; }

struct AdamHyper {  // the reference's hyper-parameters are Python doubles: 1 - beta is formed in double, then rounded
  double lr, beta1, beta2;
  float b1, b2, omb1, omb2, eps, max_norm, weight_decay;
  int noam;
  float warmup;
};

struct AdamCoef {
  float clip, wd, omb1, b2, omb2, step_size, inv_bc2_sqrt, eps;
};

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamCoef& c) {
  const float wd = c.wd, step_size = c.step_size, inv_bc2_sqrt = c.inv_bc2_sqrt, eps = c.eps;
  g *= c.clip;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(c.omb1, g - m, m);                 // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(c.omb2, g * g, v * c.b2);          // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * inv_bc2_sqrt + eps;
  p = fmaf(-step_size, m / denom, p);         // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamTensors T, const AdamHyper h,
                                                   const float* __restrict__ sqnorm,
                                                   const int64_t* __restrict__ step_dev) {
  int ti;
  int64_t start;
  if (!locate(T, blockIdx.x, &ti, &start)) return;      // uniform per CTA
  // The step's coefficients are the same for every element: ONE thread per CTA forms them (double-precision pow /
  // sqrt: a few hundred FP64 instructions) and the CTA reads them from shared memory -- evaluated by all 256
  // threads they kept the SM's FP64 pipe busy for longer than the CTA's memory traffic takes.
  __shared__ AdamCoef coef;
  if (threadIdx.x == 0) {
    const float step = static_cast<float>(*step_dev);
    float clip = 1.f;
    if (h.max_norm > 0.f) {
      const float total = sqrtf(*sqnorm);
      clip = fminf(h.max_norm / (total + 1e-6f), 1.f);   // torch.nn.utils.clip_grad_norm_
    }
    double lr = h.lr;
    if (h.noam) {  // optimizers.py:214-219
      const double sd = static_cast<double>(step);
      lr = h.lr * fmin(1.0 / sqrt(sd), sd * pow(static_cast<double>(h.warmup), -1.5));
    }
    const double bc1 = 1.0 - pow(h.beta1, static_cast<double>(step));
    const double bc2 = 1.0 - pow(h.beta2, static_cast<double>(step));
    AdamCoef c0;
    c0.clip = clip;
    c0.wd = h.weight_decay;
    c0.omb1 = h.omb1;
    c0.b2 = h.b2;
    c0.omb2 = h.omb2;
    c0.step_size = static_cast<float>(lr / bc1);
    c0.inv_bc2_sqrt = static_cast<float>(1.0 / sqrt(bc2));
    c0.eps = h.eps;
    coef = c0;
  }
  __syncthreads();
  const AdamCoef c = coef;
  const psb_adam_tensor_t t = T.t[ti];
  const int64_t end = min(t.n, start + kAdamChunk);
  const bool al = ((reinterpret_cast<uintptr_t>(t.p) | reinterpret_cast<uintptr_t>(t.g) |
                    reinterpret_cast<uintptr_t>(t.m) | reinterpret_cast<uintptr_t>(t.v)) & 15) == 0;
  if (al) {
    for (int64_t i = start + threadIdx.x * 4; i < end; i += 1024) {
      if (i + 4 <= end) {
        float4 p = *reinterpret_cast<float4*>(t.p + i);
        const float4 g = *reinterpret_cast<const float4*>(t.g + i);
        float4 m = *reinterpret_cast<float4*>(t.m + i);
        float4 v = *reinterpret_cast<float4*>(t.v + i);
        adam_one(p.x, g.x, m.x, v.x, c);
        adam_one(p.y, g.y, m.y, v.y, c);
        adam_one(p.z, g.z, m.z, v.z, c);
        adam_one(p.w, g.w, m.w, v.w, c);
        *reinterpret_cast<float4*>(t.p + i) = p;
        *reinterpret_cast<float4*>(t.m + i) = m;
        *reinterpret_cast<float4*>(t.v + i) = v;
      } else {
        for (int64_t j = i; j < end; ++j)
          adam_one(t.p[j], t.g[j], t.m[j], t.v[j], c);
      }
    }
  } else {
    for (int64_t i = start + threadIdx.x; i < end; i += 256)
      adam_one(t.p[i], t.g[i], t.m[i], t.v[i], c);
  }
}

}  // namespace psb

using namespace psb;

static int64_t adam_chunks(const psb_adam_tensor_t* t, int32_t n) {
  int64_t c = 0;
  for (int i = 0; i < n; ++i) c += (t[i].n + kAdamChunk - 1) / kAdamChunk;
  return c;
}

extern "C" int64_t psb_adam_workspace_bytes(const psb_adam_tensor_t* tensors, int32_t n_tensors) {
  if (tensors == nullptr || n_tensors <= 0 || n_tensors > PSB_ADAM_MAX_TENSORS) return PSB_E_ARG;
  return (adam_chunks(tensors, n_tensors) + 4) * static_cast<int64_t>(sizeof(float));
}

extern "C" int psb_adam_step(const psb_adam_tensor_t* tensors, int32_t n_tensors, double lr, double beta1, double beta2,
                             double eps, double weight_decay, double max_grad_norm, int32_t noam, double warmup_steps,
                             int32_t norm_given, int64_t* step_dev, float* sqnorm_dev, void* workspace, int64_t workspace_bytes,
                             psb_stream_t stream) {
  if (tensors == nullptr || n_tensors <= 0 || n_tensors > PSB_ADAM_MAX_TENSORS || step_dev == nullptr ||
      sqnorm_dev == nullptr || workspace == nullptr)
    return PSB_E_ARG;
  AdamTensors T;
  T.n = n_tensors;
  for (int i = 0; i < n_tensors; ++i) {
    if (tensors[i].p == nullptr || tensors[i].g == nullptr || tensors[i].m == nullptr || tensors[i].v == nullptr ||
        tensors[i].n <= 0)
      return PSB_E_ARG;
    T.t[i] = tensors[i];
  }
  const int64_t chunks = adam_chunks(tensors, n_tensors);
  if (chunks > (1ll << 30)) return PSB_E_DIM;
  if (workspace_bytes < (chunks + 4) * static_cast<int64_t>(sizeof(float))) return PSB_E_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  int st;
  if (norm_given == 2) {
    // norm and step counter were both written by the caller
  } else if (norm_given) {
    PSB_PROF("bump_step_kernel", s);
    bump_step_kernel<<<1, 1, 0, s>>>(step_dev);
    if ((st = launch_status()) != PSB_OK) return st;
  } else {
    PSB_PROF("sqnorm_partial_kernel", s);
    sqnorm_partial_kernel<<<static_cast<int>(chunks), 256, 0, s>>>(T, partial);
    if ((st = launch_status()) != PSB_OK) return st;
    PSB_PROF("sqnorm_final_kernel", s);
    sqnorm_final_kernel<<<1, 256, 0, s>>>(partial, static_cast<int>(chunks), sqnorm_dev, step_dev);
    if ((st = launch_status()) != PSB_OK) return st;
  }
  AdamHyper h;
  h.lr = lr;
  h.beta1 = beta1;
  h.beta2 = beta2;
  h.b1 = static_cast<float>(beta1);
  h.b2 = static_cast<float>(beta2);
  h.omb1 = static_cast<float>(1.0 - beta1);
  h.omb2 = static_cast<float>(1.0 - beta2);
  h.eps = static_cast<float>(eps);
  h.max_norm = static_cast<float>(max_grad_norm);
  h.weight_decay = static_cast<float>(weight_decay);
  h.noam = noam;
  h.warmup = static_cast<float>(warmup_steps);
  PSB_PROF("adam_kernel", s);
  adam_kernel<<<static_cast<int>(chunks), 256, 0, s>>>(T, h, sqnorm_dev, step_dev);
  return launch_status();
}


extern "C" int psb_grad_sqnorm(const psb_adam_tensor_t* tensors, int32_t n_tensors, float* sqnorm_out, void* workspace,
                               int64_t workspace_bytes, psb_stream_t stream) {
  if (tensors == nullptr || n_tensors <= 0 || n_tensors > PSB_ADAM_MAX_TENSORS || sqnorm_out == nullptr ||
      workspace == nullptr)
    return PSB_E_ARG;
  AdamTensors T;
  T.n = n_tensors;
  for (int i = 0; i < n_tensors; ++i) {
    if (tensors[i].g == nullptr || tensors[i].n <= 0) return PSB_E_ARG;
    T.t[i] = tensors[i];
  }
  const int64_t chunks = adam_chunks(tensors, n_tensors);
  if (chunks > (1ll << 30)) return PSB_E_DIM;
  if (workspace_bytes < (chunks + 4) * static_cast<int64_t>(sizeof(float))) return PSB_E_WORKSPACE;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* partial = static_cast<float*>(workspace);
  int st;
  PSB_PROF("sqnorm_partial_kernel", s);
  sqnorm_partial_kernel<<<static_cast<int>(chunks), 256, 0, s>>>(T, partial);
  if ((st = launch_status()) != PSB_OK) return st;
  PSB_PROF("sqnorm_final_kernel", s);
  sqnorm_final_kernel<<<1, 256, 0, s>>>(partial, static_cast<int>(chunks), sqnorm_out, nullptr);
  return launch_status();
}
