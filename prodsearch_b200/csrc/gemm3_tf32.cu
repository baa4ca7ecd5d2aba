// out[M, J] = A[M, K] . Bt[J, K]^T (+ bias) on tcgen05 with fp32-grade accuracy: the "3xTF32" split.
//
// Why: the encoder's projections (models/neural.py:98-231 MultiHeadedAttention linear_keys / linear_values /
// linear_query, models/transformer.py:37-88) are GEMM-shaped but run as fp32 FFMA tiles (encoder_fwd.cu rows_gemm_kernel,
// 18.6 us per launch at batch 384) because plain TF32 misses the 1e-5 parity bar by two orders of magnitude
// (tools/tf32x3_numerics.py: 8.6e-4).  kind::tf32 reads fp32 words and uses 19 of their bits, so
//     a = a_hi + a_lo,   a_hi = a with the low 13 mantissa bits cleared,   a_lo = a - a_hi   (exact in fp32)
//     a.b = a_hi.b_hi + a_lo.b_hi + a_hi.b_lo + O(2^-22 |a||b|)
// is three MMAs per k-step with fp32 accumulation in TMEM.  The two correction products go to an accumulator of their
// own (64 extra TMEM columns): whatever rounding the accumulate step uses, its error there is relative to a sum 2^-11
// times smaller (same study: <= 3e-6 of max|out| even if every accumulate truncates; 8e-7 with round-to-nearest).
//
// One CTA = 128 rows x N output columns (N = 64, K chunks of 64, 2 stages; or N = 128, K chunks of 32, 3 stages):
//   warp 0      TMA producer: A chunk (boxes of 128 rows x 32 fp32) and Bt chunk (boxes of N rows x 32 fp32),
//               128-byte swizzle, straight from the row-major operands (K contiguous = K-major, no transposes)
//   warps 2..5  split: every fp32 word of the landed tiles is rewritten in place as its hi part (so the tensor core sees
//               the same value whether it truncates or rounds) and its lo part goes to the twin tile at the SAME offset --
//               an elementwise map, so the swizzled layout carries over; fence.proxy.async, then one arrive per warp
//   warp 1      MMA issuer (converged warp, elected-lane predicate: see catalog_tc.cu tc_mma_f16_if for why):
//               per k-step of 8: main += hi.hi, corr += lo.hi, corr += hi.lo   (M 128, N, K 8)
//   warps 2..5  epilogue: thread = output row (TMEM lane) holding its N sums (main + corr) in registers, handed to an
//               epilogue functor: bias (projections), bias + gelu + dropout (FFN up), or bias + dropout + residual +
//               LayerNorm over the whole d = 128 row (out-projection, FFN down) -- a row never leaves its thread
// Rows >= M (M may live on the device: the number of active tokens) are computed on whatever the buffer holds and not
// stored; a GEMM row depends on its own A row only.  CTAs whose first row is >= M exit at once.
//
// PSB_ENC_TC=1: the encoder forward's q and K|V projections run here (rows_gemm_kernel otherwise).
// PSB_ENC_TC=2: additionally the forward tail (tail_fwd_kernel: 52 us at batch 384) becomes tail_ctx_kernel + three of
// these GEMMs -- out-projection + LN (18 CTAs), FFN up (144 CTAs), FFN down + LN (18 CTAs) -- writing the same saved
// tensors with the same dropout streams, so the FFMA backward kernels consume them unchanged.
//
// Status: written at the end of round 1 WITHOUT a GPU run (compiled, SASS read).  Off by default; the standalone entry
// psb_debug_gemm3_tf32 + profiles/check_gemm3.py are its first GPU call in round 2, then the encoder / model GPU tests
// with PSB_ENC_TC=1 and 2 (profiles/run_round2_first.sh).
#include <stdlib.h>

#include "encoder_common.cuh"
#include "tc_ptx.cuh"

namespace psb {
namespace enc {

constexpr int kG3M = 128;                                   // rows per CTA (UMMA M = TMEM lanes)
constexpr int kG3Threads = 64 + 128;                        // TMA warp, MMA warp, 4 split / epilogue warps

// Tile shapes: N output columns per CTA (UMMA N), KC floats of K per stage.  <64, 64>: projections and the FFN's first
// layer (many column tiles); <128, 32>: products whose epilogue needs the whole d = 128 row (LayerNorm).
template <int N, int KC>
struct G3Cfg {
  static_assert((N == 64 || N == 128) && (KC == 32 || KC == 64), "tile shapes");
  static constexpr int kKB = KC / 32;                                   // 128-byte swizzle rows per stage
  static constexpr uint32_t kABlock = kG3M * 128;                       // 128 rows x 128 B = 16 KB
  static constexpr uint32_t kBBlock = N * 128;                          // N rows x 128 B
  static constexpr uint32_t kAStage = kKB * kABlock;                    // hi (or lo) A tile of one stage
  static constexpr uint32_t kBStage = kKB * kBBlock;
  static constexpr uint32_t kStageBytes = 2 * (kAStage + kBStage);      // [A hi][A lo][B hi][B lo]
  static constexpr int kStages = kStageBytes <= 64 * 1024 ? 3 : 2;      // 64 KB x 3 or 96 KB x 2
  static constexpr size_t kSmem = static_cast<size_t>(kStages) * kStageBytes + 1024 /* alignment slack */ + 128 /* barriers */;
  static constexpr int kTmemCols = 2 * N;                               // main accumulator: columns 0..N-1, corrections: N..2N-1
  // cute::UMMA::InstrDescriptor (see catalog_tc.cu kIdesc): F32 accumulate, TF32 x TF32, K-major A and B, N, M = 128
  static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(N >> 3) << 17) |
                                     (static_cast<uint32_t>(kG3M >> 4) << 24);
};

struct G3Params {
  const int32_t* m_dev;   // rows on the device (active tokens) or NULL
  int m_host;
  int K;
  int split;              // Bt rows [0, split) come from map_b0, the rest from map_b1 (K | V: two weight tensors)
};

__device__ __forceinline__ uint32_t g3_elect_one() {
  uint32_t leader;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
  return leader;
}
__device__ __forceinline__ void g3_mma_tf32_if(uint32_t leader, uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %5, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void g3_commit_if(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)),
      "r"(leader)
      : "memory");
}

// hi = the 19 bits kind::tf32 uses (sign, exponent, 10 mantissa bits); lo = the rest, exact
__device__ __forceinline__ void g3_split(float4 v, float4& hi, float4& lo) {
  hi.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  hi.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  hi.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  hi.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  lo.x = v.x - hi.x;
  lo.y = v.y - hi.y;
  lo.z = v.z - hi.z;
  lo.w = v.w - hi.w;
}

// ------------------------------------------------------------------ epilogues: thread = one output row, N columns
// out = acc + bias
template <int N>
struct EpiBias {
  const float* bias;      // [J] or NULL
  float* out;
  int ldo;
  __device__ __forceinline__ void operator()(int row, int n0, float (&acc)[N]) const {
    float* o = out + static_cast<size_t>(row) * ldo + n0;
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      float4 r = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
      if (bias != nullptr) {
        const float4 b = *reinterpret_cast<const float4*>(bias + n0 + i);
        r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
      }
      *reinterpret_cast<float4*>(o + i) = r;
    }
  }
};

// FFN first layer (encoder_fwd.cu tail_fwd_kernel step 5): pre1 = acc + b1 (saved), h1 = dropout_3(gelu(pre1)) (saved)
template <int N>
struct EpiFfnUp {
  const float* b1;
  float *pre1, *h1;
  int F;
  const uint64_t* seed_dev;
  uint32_t thr;
  float keep;
  __device__ __forceinline__ void operator()(int row, int n0, float (&acc)[N]) const {
    const Drop drop = make_drop(seed_dev, thr, keep);
    const size_t base = static_cast<size_t>(row) * F + n0;
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(b1 + n0 + i);
      const float4 v = make_float4(acc[i] + b.x, acc[i + 1] + b.y, acc[i + 2] + b.z, acc[i + 3] + b.w);
      *reinterpret_cast<float4*>(pre1 + base + i) = v;
      float4 h = make_float4(gelu_tanh(v.x), gelu_tanh(v.y), gelu_tanh(v.z), gelu_tanh(v.w));
      if (drop.on()) {
        const float4 m = drop.mul4(3u, base + i);
        h.x *= m.x; h.y *= m.y; h.z *= m.z; h.w *= m.w;
      }
      *reinterpret_cast<float4*>(h1 + base + i) = h;
    }
  }
};

// Projection + dropout + residual + LayerNorm over the whole d = N = 128 row (tail_fwd_kernel steps 3 + 4 and 6 + 7):
// pre = dropout_sid(acc + bias) + res[row / res_div] (saved), out = LN(pre).  The row sums are formed in the order the
// warp-per-row kernels use (lane l holds columns 4l..4l+3, then the xor butterfly 16, 8, 4, 2, 1).
template <int N>
struct EpiResLn {
  static_assert(N == 128, "the LayerNorm epilogue owns the whole d = 128 row");
  const float *bias, *res, *ln_g, *ln_b;
  float *pre, *out;
  int res_div;
  uint32_t sid;
  float eps;
  const uint64_t* seed_dev;
  uint32_t thr;
  float keep;
  __device__ __forceinline__ static float butterfly(float (&s)[32]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int l = 0; l < 32; ++l)
        if ((l & o) == 0) {                 // both partners get s[l] + s[l ^ o] (addition commutes): keep one copy
          const float t = s[l] + s[l ^ o];
          s[l] = t;
          s[l ^ o] = t;
        }
    }
    return s[0];
  }
  __device__ __forceinline__ void operator()(int row, int /*n0*/, float (&acc)[N]) const {
    const Drop drop = make_drop(seed_dev, thr, keep);
    const size_t base = static_cast<size_t>(row) * N;
    const float* r = res + static_cast<size_t>(row / res_div) * N;
    float part[32];
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + i);
      float4 v = make_float4(acc[i] + b.x, acc[i + 1] + b.y, acc[i + 2] + b.z, acc[i + 3] + b.w);
      if (drop.on()) {
        const float4 m = drop.mul4(sid, base + i);
        v.x *= m.x; v.y *= m.y; v.z *= m.z; v.w *= m.w;
      }
      const float4 x = *reinterpret_cast<const float4*>(r + i);
      v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
      *reinterpret_cast<float4*>(pre + base + i) = v;
      acc[i] = v.x; acc[i + 1] = v.y; acc[i + 2] = v.z; acc[i + 3] = v.w;
      part[i >> 2] = (v.x + v.y) + (v.z + v.w);
    }
    const float mean = butterfly(part) / static_cast<float>(N);
#pragma unroll
    for (int i = 0; i < N; i += 4) {
      const float a0 = acc[i] - mean, a1 = acc[i + 1] - mean, a2 = acc[i + 2] - mean, a3 = acc[i + 3] - mean;
      part[i >> 2] = (a0 * a0 + a1 * a1) + (a2 * a2 + a3 * a3);
    }
    const RowStats st{mean, 1.f / sqrtf(butterfly(part) / static_cast<float>(N) + eps)};
#pragma unroll
    for (int i = 0; i < N; i += 4)
      *reinterpret_cast<float4*>(out + base + i) =
          ln_apply(make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]), st, *reinterpret_cast<const float4*>(ln_g + i),
                   *reinterpret_cast<const float4*>(ln_b + i));
  }
};

// ------------------------------------------------------------------ the kernel
template <int N, int KC, class Epi>
__global__ void __launch_bounds__(kG3Threads, 1)
gemm3_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                  const __grid_constant__ CUtensorMap map_b1, const G3Params P, const Epi epi) {
  using C = G3Cfg<N, KC>;
  const int M = P.m_dev != nullptr ? *P.m_dev : P.m_host;
  const int r0 = blockIdx.x * kG3M;
  if (r0 >= M) return;                                      // uniform: before any barrier / TMEM allocation
  extern __shared__ unsigned char smem_dyn[];
  // 128-byte-swizzled operand tiles need a 1024-byte aligned base
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(C::kStages) * C::kStageBytes);
  uint64_t* full = bars;                                    // [stage] TMA bytes landed
  uint64_t* split_done = bars + C::kStages;                 // [stage] hi / lo tiles written, visible to the async proxy
  uint64_t* empty = bars + 2 * C::kStages;                  // [stage] the stage's MMAs have read it
  uint64_t* acc_full = bars + 3 * C::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * C::kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * N;
  const int chunks = P.K / KC;

  if (threadIdx.x == 0) {
    for (int st = 0; st < C::kStages; ++st) {
      mbar_init(full + st, 1);
      mbar_init(split_done + st, 4);
      mbar_init(empty + st, 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(C::kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      const bool first = n0 < P.split;
      const CUtensorMap* mb = first ? &map_b0 : &map_b1;
      const int brow = first ? n0 : n0 - P.split;
      for (int c = 0; c < chunks; ++c) {
        const int st = c % C::kStages;
        const uint32_t ph = (c / C::kStages) & 1;
        mbar_wait(empty + st, ph ^ 1);
        mbar_expect_tx(full + st, C::kAStage + C::kBStage);
        unsigned char* base = smem + static_cast<size_t>(st) * C::kStageBytes;
        for (int kb = 0; kb < C::kKB; ++kb) tma_load_2d(base + kb * C::kABlock, &map_a, c * KC + kb * 32, r0, full + st);
        for (int kb = 0; kb < C::kKB; ++kb)
          tma_load_2d(base + 2 * C::kAStage + kb * C::kBBlock, mb, c * KC + kb * 32, brow, full + st);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: all 32 lanes run the loop, the instructions are predicated on the elected lane =====
    const uint32_t leader = g3_elect_one();
    for (int c = 0; c < chunks; ++c) {
      const int st = c % C::kStages;
      const uint32_t ph = (c / C::kStages) & 1;
      mbar_wait(split_done + st, ph);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + static_cast<size_t>(st) * C::kStageBytes);
      const uint64_t a_hi = umma_desc(base);
      const uint64_t a_lo = umma_desc(base + C::kAStage);
      const uint64_t b_hi = umma_desc(base + 2 * C::kAStage);
      const uint64_t b_lo = umma_desc(base + 2 * C::kAStage + C::kBStage);
#pragma unroll
      for (int kb = 0; kb < C::kKB; ++kb) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {                    // 4 x (K = 8 tf32 = 32 bytes) inside one swizzle row
          const uint64_t oa = static_cast<uint64_t>(kb) * (C::kABlock >> 4) + k4 * 2;
          const uint64_t ob = static_cast<uint64_t>(kb) * (C::kBBlock >> 4) + k4 * 2;
          const uint32_t acc = (c | kb | k4) != 0 ? 1u : 0u;
          g3_mma_tf32_if(leader, tmem_base, a_hi + oa, b_hi + ob, C::kIdesc, acc);
          g3_mma_tf32_if(leader, tmem_base + N, a_lo + oa, b_hi + ob, C::kIdesc, acc);
          g3_mma_tf32_if(leader, tmem_base + N, a_hi + oa, b_lo + ob, C::kIdesc, 1u);
        }
      }
      g3_commit_if(leader, empty + st);                     // the stage may be refilled once these MMAs have read it
    }
    g3_commit_if(leader, acc_full);
  } else {
    // ===== split (per stage), then epilogue =====
    const int t = threadIdx.x - 64;                         // 0..127
    for (int c = 0; c < chunks; ++c) {
      const int st = c % C::kStages;
      const uint32_t ph = (c / C::kStages) & 1;
      mbar_wait(full + st, ph);
      unsigned char* base = smem + static_cast<size_t>(st) * C::kStageBytes;
      float4* a_hi = reinterpret_cast<float4*>(base);
      float4* a_lo = reinterpret_cast<float4*>(base + C::kAStage);
      float4* b_hi = reinterpret_cast<float4*>(base + 2 * C::kAStage);
      float4* b_lo = reinterpret_cast<float4*>(base + 2 * C::kAStage + C::kBStage);
#pragma unroll 4
      for (int i = t; i < static_cast<int>(C::kAStage / 16); i += 128) {
        float4 hi, lo;
        g3_split(a_hi[i], hi, lo);
        a_hi[i] = hi;
        a_lo[i] = lo;
      }
#pragma unroll 4
      for (int i = t; i < static_cast<int>(C::kBStage / 16); i += 128) {
        float4 hi, lo;
        g3_split(b_hi[i], hi, lo);
        b_hi[i] = hi;
        b_lo[i] = lo;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
      __syncwarp();
      if (lane == 0) mbar_arrive(split_done + st);
    }
    const int quarter = warp & 3;                           // a warp may only touch TMEM lanes 32 * (warp % 4) .. + 31
    const int row = r0 + quarter * 32 + lane;
    const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float acc[N];
#pragma unroll
    for (int cc = 0; cc < N / 32; ++cc) {
      uint32_t vm[32], vc[32];
      __syncwarp();
      tc_ld32x2(t_addr + static_cast<uint32_t>(cc * 32), vm, vc, static_cast<uint32_t>(N));
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[cc * 32 + i] = __uint_as_float(vm[i]) + __uint_as_float(vc[i]);
    }
    if (row < M) epi(row, n0, acc);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols));
  }
}

// ------------------------------------------------------------------ host side
// rows x cols fp32, row stride ld floats -> boxes of box_rows rows x 32 floats, 128-byte swizzle, OOB reads as zero
static int g3_make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) return PSB_E_UNSUPPORTED;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? PSB_OK : PSB_E_ARG;
}

// PSB_ENC_TC: 0 (default) = FFMA kernels; 1 = forward q and K|V projections on tcgen05; 2 = 1 + the forward tail as
// ctx kernel + three 3xTF32 GEMMs with fused epilogues (launch_tail_fwd_tc).  Read once per process.
static int enc_tc_level() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PSB_ENC_TC");
    const int x = e != nullptr ? atoi(e) : 0;
    v = (x >= 0 && x <= 2) ? x : 0;
  }
  return v;
}
bool rows_gemm_tc_enabled() { return enc_tc_level() >= 1; }
// Unset PSB_ENC_TC: the K|V projection goes to tcgen05 by itself once it is big enough to be throughput- rather than
// latency-bound: from 16 384 token rows on (RTM: 384 x 51 = 19.6k rows for the positives, 1920 x 51 = 98k for the
// negatives: 0.31 ms on FFMA, 0.06 ms here); at TEM's 8k-row plan (~3.5k active rows) the FFMA kernel is as fast
// (profiles/r02a_bench_enc_tc.json).  PSB_ENC_TC=0 switches the automatic choice off.
bool rows_gemm_tc_auto(int64_t m_max) {
  static int allowed = -1;
  if (allowed < 0) {
    const char* e = getenv("PSB_ENC_TC");
    allowed = (e != nullptr && atoi(e) == 0 && e[0] == '0') ? 0 : 1;
  }
  return allowed == 1 && m_max >= 16384;
}
bool tail_tc_enabled() { return enc_tc_level() >= 2; }

// One launch: out-tile epilogue `epi` over A [m rows, K] (row stride lda) and Bt rows [0, split) from Bt0, the rest from Bt1
template <int N, int KC, class Epi>
static int g3_launch(const char* name, const float* A, int lda, const int32_t* m_dev, int m_host, int m_max, int K,
                     const float* Bt0, const float* Bt1, int split, int J, const Epi& epi, cudaStream_t s) {
  using C = G3Cfg<N, KC>;
  if (m_max <= 0) return PSB_OK;
  static DeviceAttr attr_done;                            // one flag per instantiation
  if (attr_done.need()) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_tf32_kernel<N, KC, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(C::kSmem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done.done();
  }
  alignas(64) CUtensorMap map_a, map_b0, map_b1;
  const int rows0 = Bt1 != nullptr ? split : J;
  int st;
  if ((st = g3_make_map(&map_a, A, m_max, K, lda, kG3M)) != PSB_OK) return st;
  if ((st = g3_make_map(&map_b0, Bt0, rows0, K, K, N)) != PSB_OK) return st;
  if (Bt1 != nullptr) {
    if ((st = g3_make_map(&map_b1, Bt1, J - split, K, K, N)) != PSB_OK) return st;
  } else {
    map_b1 = map_b0;
  }
  G3Params P;
  P.m_dev = m_dev;
  P.m_host = m_host;
  P.K = K;
  P.split = rows0;
  const dim3 grid(static_cast<unsigned>((m_max + kG3M - 1) / kG3M), static_cast<unsigned>(J / N));
  PSB_PROF(name, s);
  gemm3_tf32_kernel<N, KC, Epi><<<grid, kG3Threads, C::kSmem, s>>>(map_a, map_b0, map_b1, P, epi);
  return launch_status();
}

bool rows_gemm_tc_supported(const float* A, int lda, int K, const float* Bt0, const float* Bt1, int split, int J,
                            const float* bias, const float* out, int ldo) {
  if (K <= 0 || K % 64 != 0 || J <= 0 || J % 64 != 0 || (lda & 3) != 0 || (ldo & 3) != 0) return false;
  if (Bt1 != nullptr && (split <= 0 || split >= J || split % 64 != 0)) return false;
  return A != nullptr && Bt0 != nullptr && out != nullptr && !misaligned16(A) && !misaligned16(Bt0) &&
         !misaligned16(Bt1) && !misaligned16(bias) && !misaligned16(out);
}

int launch_rows_gemm_tc(const float* A, int lda, const int32_t* m_dev, int m_host, int m_max, int K, const float* Bt0,
                        const float* Bt1, int split, int J, const float* bias, float* out, int ldo, cudaStream_t s) {
  if (!rows_gemm_tc_supported(A, lda, K, Bt0, Bt1, split, J, bias, out, ldo)) return PSB_E_UNSUPPORTED;
  EpiBias<64> epi;
  epi.bias = bias;
  epi.out = out;
  epi.ldo = ldo;
  return g3_launch<64, 64>("gemm3_tf32_kernel", A, lda, m_dev, m_host, m_max, K, Bt0, Bt1, split, J, epi, s);
}

// ------------------------------------------------------------------ forward tail on tcgen05 (PSB_ENC_TC=2)
// Steps 1 + 2 of tail_fwd_kernel on their own: ctx[row] = sum_al dropout_1(P)[h, al] * V[al], one warp per copy row,
// lane = 4 columns (d = 128), the same sequential fmaf chain over the active tokens -> the same bits.
__global__ void __launch_bounds__(256) tail_ctx_kernel(Dims D, const int32_t* __restrict__ nact, const int32_t* __restrict__ off,
                                                       const int32_t* __restrict__ tok, const float* __restrict__ Pw,
                                                       const float* __restrict__ kv, float* __restrict__ ctx,
                                                       const uint64_t* __restrict__ seed_dev) {
  const int lane = threadIdx.x & 31;
  const int grow = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (grow >= D.S * D.C) return;
  const int s = grow / D.C, d = D.d, H = D.H, T = D.T, dh = D.dh;
  const int j = lane * 4;
  const Drop drop = make_drop(seed_dev, D.thr, D.keep);
  const int na = nact[s], base = off[s];
  const int h0 = j / dh, h1 = (j + 1) / dh, h2 = (j + 2) / dh, h3 = (j + 3) / dh;
  float4 acc = zero4();
  for (int al = 0; al < na; ++al) {
    const float* pw = Pw + static_cast<size_t>(base + al) * H;
    float w0 = pw[h0], w1 = pw[h1], w2 = pw[h2], w3 = pw[h3];
    if (drop.on()) {
      const uint64_t t = static_cast<uint64_t>(tok[base + al]);
      const float m0 = drop.mul1(1u, (static_cast<uint64_t>(grow) * H + h0) * T + t);
      if (h0 == h3) {                       // dh % 4 == 0: the lane's four columns belong to one head -> one draw
        w0 *= m0; w1 *= m0; w2 *= m0; w3 *= m0;
      } else {
        w0 *= m0;
        w1 *= drop.mul1(1u, (static_cast<uint64_t>(grow) * H + h1) * T + t);
        w2 *= drop.mul1(1u, (static_cast<uint64_t>(grow) * H + h2) * T + t);
        w3 *= drop.mul1(1u, (static_cast<uint64_t>(grow) * H + h3) * T + t);
      }
    }
    const float4 v = *reinterpret_cast<const float4*>(kv + static_cast<size_t>(base + al) * 2 * d + d + j);
    acc.x = fmaf(w0, v.x, acc.x);
    acc.y = fmaf(w1, v.y, acc.y);
    acc.z = fmaf(w2, v.z, acc.z);
    acc.w = fmaf(w3, v.w, acc.w);
  }
  *reinterpret_cast<float4*>(ctx + static_cast<size_t>(grow) * d + j) = acc;
}

bool tail_tc_supported(const TailTcArgs& a) {
  const Dims& D = a.D;
  if (D.d != 128 || D.F % 64 != 0 || D.F <= 0) return false;
  if (a.P == nullptr || a.nact == nullptr || a.off == nullptr || a.tok == nullptr) return false;
  const float* ptrs[] = {a.kv, a.xo, a.wo, a.bo, a.w1, a.b1, a.w2, a.b2, a.ln_ff_g, a.ln_ff_b, a.ln_out_g,
                         a.ln_out_b, a.ctx, a.y, a.n, a.z, a.pre1, a.h1, a.out};
  for (const float* q : ptrs)
    if (q == nullptr || misaligned16(q)) return false;
  return true;
}

int launch_tail_fwd_tc(const TailTcArgs& a, cudaStream_t s) {
  if (!tail_tc_supported(a)) return PSB_E_UNSUPPORTED;
  const Dims& D = a.D;
  const int SC = D.S * D.C, d = D.d, F = D.F;
  int st;
  PSB_PROF("tail_ctx_kernel", s);
  tail_ctx_kernel<<<(SC + 7) / 8, 256, 0, s>>>(D, a.nact, a.off, a.tok, a.P, a.kv, a.ctx, a.seed_dev);
  if ((st = launch_status()) != PSB_OK) return st;
  // y = dropout_2(ctx . Wo^T + bo) + x[o];  n = LN_ff(y)
  EpiResLn<128> e1;
  e1.bias = a.bo; e1.res = a.xo; e1.ln_g = a.ln_ff_g; e1.ln_b = a.ln_ff_b; e1.pre = a.y; e1.out = a.n;
  e1.res_div = D.C; e1.sid = 2u; e1.eps = D.eps; e1.seed_dev = a.seed_dev; e1.thr = D.thr; e1.keep = D.keep;
  if ((st = g3_launch<128, 32>("gemm3_out_proj_ln_kernel", a.ctx, d, nullptr, SC, SC, d, a.wo, nullptr, 0, d, e1, s)) != PSB_OK)
    return st;
  // pre1 = n . W1^T + b1;  h1 = dropout_3(gelu(pre1))
  EpiFfnUp<64> e2;
  e2.b1 = a.b1; e2.pre1 = a.pre1; e2.h1 = a.h1; e2.F = F; e2.seed_dev = a.seed_dev; e2.thr = D.thr; e2.keep = D.keep;
  if ((st = g3_launch<64, 64>("gemm3_ffn_up_kernel", a.n, d, nullptr, SC, SC, d, a.w1, nullptr, 0, F, e2, s)) != PSB_OK) return st;
  // z = dropout_4(h1 . W2^T + b2) + y;  out = LN_out(z)
  EpiResLn<128> e3;
  e3.bias = a.b2; e3.res = a.y; e3.ln_g = a.ln_out_g; e3.ln_b = a.ln_out_b; e3.pre = a.z; e3.out = a.out;
  e3.res_div = 1; e3.sid = 4u; e3.eps = D.eps; e3.seed_dev = a.seed_dev; e3.thr = D.thr; e3.keep = D.keep;
  return g3_launch<128, 32>("gemm3_ffn_down_ln_kernel", a.h1, F, nullptr, SC, SC, F, a.w2, nullptr, 0, d, e3, s);
}

}  // namespace enc
}  // namespace psb

extern "C" int psb_debug_gemm3_tf32(const float* a, int64_t lda, int64_t m, int64_t k, const float* bt, int64_t j,
                                    const float* bias, float* out, int64_t ldo, psb_stream_t stream) {
  if (m < 0 || m > (1 << 24) || lda < k || ldo < j) return PSB_E_ARG;
  return psb::enc::launch_rows_gemm_tc(a, static_cast<int>(lda), nullptr, static_cast<int>(m), static_cast<int>(m),
                                       static_cast<int>(k), bt, nullptr, 0, static_cast<int>(j), bias, out,
                                       static_cast<int>(ldo), static_cast<cudaStream_t>(stream));
}
