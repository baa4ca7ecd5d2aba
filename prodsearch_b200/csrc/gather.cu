// G1 row gather and G4 fused gather + masked mean (+ dropout multiplier, + fs projection).
//
// HBM-bound byte movers: one warp owns one output row, every lane
// moves 128-bit pieces of a 512 B table row (d = 128 -> exactly one LDG.128 per lane
// per row), and several independent rows are kept in flight per warp so a full SM
// carries > 40 KB of outstanding loads.  Table rows are read through the read-only
// path without L1 allocation (each row is used once by the warp that loads it).
#include "psb_common.cuh"

namespace psb {


// ---------------------------------------------------------------- G1 -----
template <int R>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float4* __restrict__ table, int64_t table_rows, int d4,
                   const int64_t* __restrict__ idx, int64_t n, float4* __restrict__ out,
                   int32_t* __restrict__ err) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t base = warp * R; base < n; base += nwarps * R) {
    int64_t r[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int64_t row = base + i;
      int64_t v = row < n ? idx[row] : -1;
      if (row < n && (v < 0 || v >= table_rows)) {
        if (err != nullptr && lane == 0) *err = 1;
        v = -1;
      }
      r[i] = v;
    }
    for (int c = lane; c < d4; c += 32) {
      float4 v[R];
#pragma unroll
      for (int i = 0; i < R; ++i) v[i] = r[i] >= 0 ? ldg_row4(table + r[i] * d4 + c) : zero4();
#pragma unroll
      for (int i = 0; i < R; ++i)
        if (base + i < n) stg4(out + (base + i) * d4 + c, v[i]);
    }
  }
}

// ---------------------------------------------------------------- G4 -----
// C = float4 chunks per lane (d4 <= 32*C).  One warp per pooled row.
// SMEMW (FS with d4 <= 32): the projection W [d, d] is staged once per CTA in shared memory (rows padded by 4
// floats so the 128-bit row reads of the 32 lanes are conflict-free) and every lane produces the outputs
// j = lane + 32 u from broadcast reads of the pooled row -- instead of 4 * d dependent L2 row loads and d warp
// reductions per pooled row (31 us of the batch-384 step were this matvec).
template <int C, bool FS, bool SMEMW = false>
__global__ void __launch_bounds__(256)
meanpool_kernel(const float4* __restrict__ table, int64_t table_rows, int d4,
                const int64_t* __restrict__ idx, int64_t n, int w, int64_t pad_idx,
                const uint8_t* __restrict__ mask, const float* __restrict__ tok_scale,
                const float4* __restrict__ keep_scale, const float4* __restrict__ fs_weight,
                const float* __restrict__ fs_bias, float4* __restrict__ mean_out,
                float* __restrict__ out, float* __restrict__ inv_count) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int d = d4 * 4;
  constexpr int U = 4;  // rows in flight per warp
  extern __shared__ float4 smem_w4[];
  const int ldw4 = d4 + 1;                       // padded row length of W in float4
  float4* s_mean4 = smem_w4 + d * ldw4 + (threadIdx.x >> 5) * d4;
  if (SMEMW) {
    for (int e = threadIdx.x; e < d * d4; e += blockDim.x) {
      const int j = e / d4, c4 = e - j * d4;
      smem_w4[j * ldw4 + c4] = __ldg(fs_weight + e);
    }
    __syncthreads();
  }
  for (int64_t i = warp; i < n; i += nwarps) {
    float4 acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = zero4();
    int cnt = 0;
    for (int j0 = 0; j0 < w; j0 += 32) {
      const int j = j0 + lane;
      int64_t my = -1;
      float sc = 0.f;
      bool valid = false;
      if (j < w) {
        my = idx[i * w + j];
        valid = mask != nullptr ? mask[i * w + j] != 0 : (pad_idx < 0 || my != pad_idx);
        if (valid) sc = tok_scale != nullptr ? tok_scale[i * w + j] : 1.f;
        if (my < 0 || my >= table_rows) sc = 0.f;
      }
      cnt += __popc(__ballot_sync(kFull, valid));
      const int nj = min(32, w - j0);
      for (int jj = 0; jj < nj; jj += U) {
        int64_t r[U];
        float s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int src = min(jj + u, 31);
          r[u] = __shfl_sync(kFull, my, src);
          s[u] = (jj + u < nj) ? __shfl_sync(kFull, sc, src) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int col = lane + 32 * c;
          float4 v[U];
#pragma unroll
          for (int u = 0; u < U; ++u)
            v[u] = (s[u] != 0.f && col < d4) ? ldg_row4(table + r[u] * d4 + col) : zero4();
          // (row * scale) rounded, then added, in token order: the reference's
          // (x * mask).sum(1) (text_encoder.py:8) without FMA contraction.
#pragma unroll
          for (int u = 0; u < U; ++u) {
            acc[c].x = __fadd_rn(acc[c].x, __fmul_rn(v[u].x, s[u]));
            acc[c].y = __fadd_rn(acc[c].y, __fmul_rn(v[u].y, s[u]));
            acc[c].z = __fadd_rn(acc[c].z, __fmul_rn(v[u].z, s[u]));
            acc[c].w = __fadd_rn(acc[c].w, __fmul_rn(v[u].w, s[u]));
          }
        }
      }
    }
    const float denom = static_cast<float>(cnt > 0 ? cnt : 1);
    if (inv_count != nullptr && lane == 0) inv_count[i] = 1.f / denom;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int col = lane + 32 * c;
      acc[c].x = acc[c].x / denom;
      acc[c].y = acc[c].y / denom;
      acc[c].z = acc[c].z / denom;
      acc[c].w = acc[c].w / denom;
      if (col < d4) {
        if (keep_scale != nullptr) {
          const float4 ks = keep_scale[i * d4 + col];
          acc[c].x *= ks.x; acc[c].y *= ks.y; acc[c].z *= ks.z; acc[c].w *= ks.w;
        }
        if (mean_out != nullptr) stg4(mean_out + i * d4 + col, acc[c]);
        if (!FS) stg4(reinterpret_cast<float4*>(out) + i * d4 + col, acc[c]);
      } else {
        acc[c] = zero4();
      }
    }
    if (FS && SMEMW) {
      // out[i, j] = tanh(<W[j, :], mean> + b[j]) with W and the pooled row in shared memory
      __syncwarp();
      if (lane < d4) s_mean4[lane] = acc[0];
      __syncwarp();
      float o[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c4 = 0; c4 < d4; ++c4) {
        const float4 m = s_mean4[c4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = lane + 32 * u;
          if (j < d) {
            const float4 wv = smem_w4[j * ldw4 + c4];
            o[u] = fmaf(m.w, wv.w, fmaf(m.z, wv.z, fmaf(m.y, wv.y, fmaf(m.x, wv.x, o[u]))));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = lane + 32 * u;
        if (j < d) out[i * d + j] = tanhf(o[u] + fs_bias[j]);
      }
    } else if (FS) {
      // out[i, j] = tanh(<W[j, :], mean> + b[j]); W (d x d fp32) stays L1/L2 resident.
      for (int j0 = 0; j0 < d; j0 += 32) {
        float mine = 0.f;
        for (int jj = 0; jj < 32; jj += 4) {
          float p[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + jj + u;
            float t = 0.f;
            if (j < d) {
#pragma unroll
              for (int c = 0; c < C; ++c) {
                const int col = lane + 32 * c;
                if (col < d4) t += dot4(acc[c], __ldg(fs_weight + static_cast<int64_t>(j) * d4 + col));
              }
            }
            p[u] = t;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float t = warp_sum(p[u]);
            if (lane == jj + u) mine = t;
          }
        }
        const int j = j0 + lane;
        if (j < d) out[i * d + j] = tanhf(mine + fs_bias[j]);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
token_weights_kernel(const int64_t* __restrict__ idx, int64_t n, int w, int64_t pad_idx,
                     const uint8_t* __restrict__ mask, float* __restrict__ tw) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t warp = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int64_t i = warp; i < n; i += nwarps) {
    int cnt = 0;
    for (int j0 = 0; j0 < w; j0 += 32) {
      const int j = j0 + lane;
      bool valid = false;
      if (j < w) valid = mask != nullptr ? mask[i * w + j] != 0 : (pad_idx < 0 || idx[i * w + j] != pad_idx);
      cnt += __popc(__ballot_sync(kFull, valid));
    }
    const float denom = static_cast<float>(cnt > 0 ? cnt : 1);
    for (int j = lane; j < w; j += 32) {
      const bool valid = mask != nullptr ? mask[i * w + j] != 0 : (pad_idx < 0 || idx[i * w + j] != pad_idx);
      tw[i * w + j] = valid ? 1.f / denom : 0.f;
    }
  }
}

// ---------------------------------------------------------------- fs backward
// Block roles: blockIdx.x < rows_blocks -> grad_mean rows (one warp per row);
// the remaining d blocks -> grad_weight[j,:] / grad_bias[j] (8 warps split i into 8
// contiguous ranges, partials combined in warp order: deterministic for a given n).
template <int C>
__global__ void __launch_bounds__(256)
fs_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ outv,
              const float4* __restrict__ mean, const float4* __restrict__ keep_scale,
              const float4* __restrict__ W, int64_t n, int d4, int rows_blocks,
              float4* __restrict__ grad_W, float* __restrict__ grad_b,
              float4* __restrict__ grad_mean) {
  const int lane = threadIdx.x & 31;
  const int wid = threadIdx.x >> 5;
  const int d = d4 * 4;
  __shared__ float4 part[8][128];
  __shared__ float part_b[8];
  if (static_cast<int>(blockIdx.x) < rows_blocks) {
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * 8 + wid; i < n; i += static_cast<int64_t>(rows_blocks) * 8) {
      float4 acc[C];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = zero4();
      for (int j0 = 0; j0 < d; j0 += 32) {
        const int j = j0 + lane;
        float dz = 0.f;
        if (j < d) {
          const float o = outv[i * d + j];
          dz = grad_out[i * d + j] * (1.f - o * o);
        }
        const int nj = min(32, d - j0);
        // 8 rows of W in flight per batch (the loads do not depend on the running sums)
        for (int jb = 0; jb < nj; jb += 8) {
          float4 w[8][C];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const bool ok = jb + u < nj;
#pragma unroll
            for (int c = 0; c < C; ++c) {
              const int col = lane + 32 * c;
              w[u][c] = (ok && col < d4) ? __ldg(W + static_cast<int64_t>(j0 + jb + u) * d4 + col) : zero4();
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float sdz = __shfl_sync(kFull, dz, (jb + u) & 31);
            if (jb + u < nj) {
#pragma unroll
              for (int c = 0; c < C; ++c) fma4(acc[c], sdz, w[u][c]);
            }
          }
        }
      }
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const int col = lane + 32 * c;
        if (col < d4) {
          if (keep_scale != nullptr) {
            const float4 ks = keep_scale[i * d4 + col];
            acc[c].x *= ks.x; acc[c].y *= ks.y; acc[c].z *= ks.z; acc[c].w *= ks.w;
          }
          grad_mean[i * d4 + col] = acc[c];
        }
      }
    }
    return;
  }
  const int j = blockIdx.x - rows_blocks;  // output row of grad_W
  const int64_t per = (n + 7) / 8;
  const int64_t lo = per * wid;
  const int64_t hi = (lo + per < n) ? lo + per : n;
  float4 acc[C];
#pragma unroll
  for (int c = 0; c < C; ++c) acc[c] = zero4();
  float accb = 0.f;
  // 32 rows per round: lane l computes dz of row i0 + l (independent loads), then the rows are consumed in
  // ascending order -- the summation order of a scalar loop -- with 8 rows of `mean` in flight
  for (int64_t i0 = lo; i0 < hi; i0 += 32) {
    const int64_t il = i0 + lane;
    float dzl = 0.f;
    if (il < hi) {
      const float o = outv[il * d + j];
      dzl = grad_out[il * d + j] * (1.f - o * o);
    }
    const int cnt = static_cast<int>(hi - i0 < 32 ? hi - i0 : 32);
    for (int u0 = 0; u0 < cnt; u0 += 8) {
      float4 m[8][C];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const bool ok = u0 + u < cnt;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const int col = lane + 32 * c;
          m[u][c] = (ok && col < d4) ? mean[(i0 + u0 + u) * d4 + col] : zero4();
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float dz = __shfl_sync(kFull, dzl, (u0 + u) & 31);
        if (u0 + u < cnt) {
          accb += dz;
#pragma unroll
          for (int c = 0; c < C; ++c) fma4(acc[c], dz, m[u][c]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const int col = lane + 32 * c;
    if (col < d4) part[wid][col] = acc[c];
  }
  if (lane == 0) part_b[wid] = accb;
  __syncthreads();
  if (wid == 0) {
    for (int col = lane; col < d4; col += 32) {
      float4 t = part[0][col];
      for (int q = 1; q < 8; ++q) {
        const float4 p = part[q][col];
        t.x += p.x; t.y += p.y; t.z += p.z; t.w += p.w;
      }
      grad_W[static_cast<int64_t>(j) * d4 + col] = t;
    }
    if (lane == 0) {
      float t = part_b[0];
      for (int q = 1; q < 8; ++q) t += part_b[q];
      grad_b[j] = t;
    }
  }
}

}  // namespace psb

using namespace psb;

extern "C" int psb_abi_version(void) { return PSB_ABI_VERSION; }


extern "C" const char* psb_status_string(int status) {
  switch (status) {
    case PSB_OK: return "ok";
    case PSB_E_ARG: return "invalid argument (null pointer or negative size)";
    case PSB_E_DIM: return "unsupported dimension (d must be a multiple of 4 and <= 512; k out of range)";
    case PSB_E_WORKSPACE: return "workspace too small";
    case PSB_E_ALIGN: return "pointer not 16-byte aligned";
    case PSB_E_UNSUPPORTED: return "unsupported mode";
    default: return status > 0 ? cudaGetErrorString(static_cast<cudaError_t>(status)) : "unknown status";
  }
}

extern "C" int psb_gather_rows(const float* table, int64_t table_rows, int64_t d, const int64_t* idx,
                               int64_t n, float* out, int32_t* err_flag, psb_stream_t stream) {
  int st = check_table_args(table, table_rows, d);
  if (st != PSB_OK) return st;
  if (n < 0 || (n > 0 && (idx == nullptr || out == nullptr))) return PSB_E_ARG;
  if (misaligned16(out)) return PSB_E_ALIGN;
  if (n == 0) return PSB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  constexpr int R = 8;
  const int grid = grid_for(n, 8 * R);
  PSB_PROF("gather_rows_kernel", s);
  gather_rows_kernel<R><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(table), table_rows,
                                             static_cast<int>(d / 4), idx, n,
                                             reinterpret_cast<float4*>(out), err_flag);
  return launch_status();
}

template <int C>
static int launch_meanpool(const float* table, int64_t table_rows, int64_t d, const int64_t* idx, int64_t n,
                           int64_t w, int64_t pad_idx, const uint8_t* mask, const float* tok_scale,
                           const float* keep_scale, const float* fs_weight, const float* fs_bias,
                           float* mean_out, float* out, float* inv_count, cudaStream_t s) {
  const int grid = grid_for(n, 8);
  const float4* t4 = reinterpret_cast<const float4*>(table);
  const float4* k4 = reinterpret_cast<const float4*>(keep_scale);
  const float4* w4 = reinterpret_cast<const float4*>(fs_weight);
  float4* m4 = reinterpret_cast<float4*>(mean_out);
  PSB_PROF("meanpool_kernel", s);
  if (fs_weight != nullptr && C == 1) {
    const size_t smem = (static_cast<size_t>(d) * (d / 4 + 1) + 8 * (d / 4)) * sizeof(float4);
    static DeviceAttr configured;
    if (configured.need(smem)) {
      cudaError_t e = cudaFuncSetAttribute(meanpool_kernel<1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return static_cast<int>(e);
      configured.done(smem);
    }
    const int grid_w = grid_for(n, 8, 2);            // W is re-staged per CTA: few, persistent CTAs
    meanpool_kernel<1, true, true><<<grid_w, 256, smem, s>>>(t4, table_rows, static_cast<int>(d / 4), idx, n,
                                                              static_cast<int>(w), pad_idx, mask, tok_scale, k4, w4,
                                                              fs_bias, m4, out, inv_count);
  } else if (fs_weight != nullptr)
    meanpool_kernel<C, true><<<grid, 256, 0, s>>>(t4, table_rows, static_cast<int>(d / 4), idx, n,
                                                  static_cast<int>(w), pad_idx, mask, tok_scale, k4, w4,
                                                  fs_bias, m4, out, inv_count);
  else
    meanpool_kernel<C, false><<<grid, 256, 0, s>>>(t4, table_rows, static_cast<int>(d / 4), idx, n,
                                                   static_cast<int>(w), pad_idx, mask, tok_scale, k4, w4,
                                                   fs_bias, m4, out, inv_count);
  return launch_status();
}

extern "C" int psb_gather_meanpool_fwd(const float* table, int64_t table_rows, int64_t d, const int64_t* idx,
                                       int64_t n, int64_t w, int64_t pad_idx, const uint8_t* mask,
                                       const float* tok_scale, const float* keep_scale,
                                       const float* fs_weight, const float* fs_bias, float* mean_out,
                                       float* out, float* inv_count, psb_stream_t stream) {
  int st = check_table_args(table, table_rows, d);
  if (st != PSB_OK) return st;
  if (n < 0 || w <= 0 || w > (1 << 20)) return PSB_E_ARG;
  if (n > 0 && (idx == nullptr || out == nullptr)) return PSB_E_ARG;
  if (fs_weight != nullptr && (fs_bias == nullptr || mean_out == nullptr)) return PSB_E_ARG;
  if (misaligned16(out) || misaligned16(mean_out) || misaligned16(keep_scale) || misaligned16(fs_weight))
    return PSB_E_ALIGN;
  if (n == 0) return PSB_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t d4 = d / 4;
  if (d4 <= 32)
    return launch_meanpool<1>(table, table_rows, d, idx, n, w, pad_idx, mask, tok_scale, keep_scale, fs_weight,
                              fs_bias, mean_out, out, inv_count, s);
  if (d4 <= 64)
    return launch_meanpool<2>(table, table_rows, d, idx, n, w, pad_idx, mask, tok_scale, keep_scale, fs_weight,
                              fs_bias, mean_out, out, inv_count, s);
  return launch_meanpool<4>(table, table_rows, d, idx, n, w, pad_idx, mask, tok_scale, keep_scale, fs_weight,
                            fs_bias, mean_out, out, inv_count, s);
}

extern "C" int psb_meanpool_token_weights(const int64_t* idx, int64_t n, int64_t w, int64_t pad_idx,
                                          const uint8_t* mask, float* tok_weight, psb_stream_t stream) {
  if (n < 0 || w <= 0 || tok_weight == nullptr || (idx == nullptr && mask == nullptr)) return PSB_E_ARG;
  if (n == 0) return PSB_OK;
  PSB_PROF("token_weights_kernel", static_cast<cudaStream_t>(stream));
  token_weights_kernel<<<grid_for(n, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      idx, n, static_cast<int>(w), pad_idx, mask, tok_weight);
  return launch_status();
}

extern "C" int psb_fs_bwd(const float* grad_out, const float* out, const float* mean, const float* keep_scale,
                          const float* fs_weight, int64_t n, int64_t d, float* grad_weight, float* grad_bias,
                          float* grad_mean, psb_stream_t stream) {
  if (grad_out == nullptr || out == nullptr || mean == nullptr || fs_weight == nullptr ||
      grad_weight == nullptr || grad_bias == nullptr || grad_mean == nullptr || n <= 0)
    return PSB_E_ARG;
  if (d <= 0 || (d & 3) != 0 || d > 512) return PSB_E_DIM;
  if (misaligned16(mean) || misaligned16(keep_scale) || misaligned16(fs_weight) || misaligned16(grad_weight) ||
      misaligned16(grad_mean))
    return PSB_E_ALIGN;
  const int rows_blocks = grid_for(n, 8, 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PSB_PROF("fs_bwd_kernel", s);
#define PSB_FS_LAUNCH(C)                                                                                            \
  fs_bwd_kernel<C><<<rows_blocks + static_cast<int>(d), 256, 0, s>>>(                                               \
      grad_out, out, reinterpret_cast<const float4*>(mean), reinterpret_cast<const float4*>(keep_scale),            \
      reinterpret_cast<const float4*>(fs_weight), n, static_cast<int>(d / 4), rows_blocks,                          \
      reinterpret_cast<float4*>(grad_weight), grad_bias, reinterpret_cast<float4*>(grad_mean))
  if (d <= 128) PSB_FS_LAUNCH(1);
  else if (d <= 256) PSB_FS_LAUNCH(2);
  else PSB_FS_LAUNCH(4);
#undef PSB_FS_LAUNCH
  return launch_status();
}
