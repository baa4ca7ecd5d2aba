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
// One CTA = 128 rows x 64 output columns, K streamed in chunks of 64 through a 2-stage ring:
//   warp 0      TMA producer: A chunk (2 boxes of 128 rows x 32 fp32) and Bt chunk (2 boxes of 64 rows x 32 fp32),
//               128-byte swizzle, straight from the row-major operands (K contiguous = K-major, no transposes)
//   warps 2..5  split: every fp32 word of the landed tiles is rewritten in place as its hi part (so the tensor core sees
//               the same value whether it truncates or rounds) and its lo part goes to the twin tile at the SAME offset --
//               an elementwise map, so the swizzled layout carries over; fence.proxy.async, then one arrive per warp
//   warp 1      MMA issuer (converged warp, elected-lane predicate: see catalog_tc.cu tc_mma_f16_if for why):
//               per k-step of 8: main += hi.hi, corr += lo.hi, corr += hi.lo   (M 128, N 64, K 8)
//   warps 2..5  epilogue: thread = output row (TMEM lane); main + corr (+ bias) -> float4 stores
// Rows >= M (M may live on the device: the number of active tokens) are computed on whatever the buffer holds and not
// stored; a GEMM row depends on its own A row only.  CTAs whose first row is >= M exit at once.
//
// Status: written at the end of round 1 WITHOUT a GPU run (compiled, SASS read).  Off unless PSB_ENC_TC=1; the
// standalone entry psb_debug_gemm3_tf32 + profiles/check_gemm3.py are its first GPU call in round 2.
#include <stdlib.h>

#include "encoder_common.cuh"
#include "tc_ptx.cuh"

namespace psb {
namespace enc {

constexpr int kG3M = 128;                                   // rows per CTA (UMMA M = TMEM lanes)
constexpr int kG3N = 64;                                    // output columns per CTA (UMMA N)
constexpr int kG3KC = 64;                                   // K per stage: two 128-byte swizzle rows of 32 fp32
constexpr int kG3Stages = 2;
constexpr int kG3Threads = 64 + 128;                        // TMA warp, MMA warp, 4 split / epilogue warps
constexpr uint32_t kG3ABlock = kG3M * 128;                  // 128 rows x 128 B = 16 KB
constexpr uint32_t kG3BBlock = kG3N * 128;                  // 64 rows x 128 B = 8 KB
constexpr uint32_t kG3AStage = 2 * kG3ABlock;               // hi (or lo) A tile of one stage: 32 KB
constexpr uint32_t kG3BStage = 2 * kG3BBlock;               // 16 KB
constexpr uint32_t kG3StageBytes = 2 * (kG3AStage + kG3BStage);   // [A hi][A lo][B hi][B lo] = 96 KB
constexpr size_t kG3Smem = static_cast<size_t>(kG3Stages) * kG3StageBytes + 1024 /* alignment slack */ + 128 /* barriers */;
constexpr int kG3TmemCols = 128;                            // main accumulator: columns 0..63, corrections: 64..127
// cute::UMMA::InstrDescriptor (see catalog_tc.cu kIdesc): F32 accumulate, TF32 x TF32, K-major A and B, N = 64, M = 128
constexpr uint32_t kG3Idesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(kG3N >> 3) << 17) |
                              (static_cast<uint32_t>(kG3M >> 4) << 24);

struct G3Params {
  const int32_t* m_dev;   // rows on the device (active tokens) or NULL
  int m_host;
  int K, J;
  int split;              // Bt rows [0, split) come from map_b0, [split, J) from map_b1 (K | V: two weight tensors)
  const float* bias;      // [J] or NULL
  float* out;
  int ldo;
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

__global__ void __launch_bounds__(kG3Threads, 1)
gemm3_tf32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                  const __grid_constant__ CUtensorMap map_b1, const G3Params P) {
  const int M = P.m_dev != nullptr ? *P.m_dev : P.m_host;
  const int r0 = blockIdx.x * kG3M;
  if (r0 >= M) return;                                      // uniform: before any barrier / TMEM allocation
  extern __shared__ unsigned char smem_dyn[];
  // 128-byte-swizzled operand tiles need a 1024-byte aligned base
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(kG3Stages) * kG3StageBytes);
  uint64_t* full = bars;                                    // [stage] TMA bytes landed
  uint64_t* split_done = bars + kG3Stages;                  // [stage] hi / lo tiles written, visible to the async proxy
  uint64_t* empty = bars + 2 * kG3Stages;                   // [stage] the stage's MMAs have read it
  uint64_t* acc_full = bars + 3 * kG3Stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kG3Stages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * kG3N;
  const int chunks = P.K / kG3KC;

  if (threadIdx.x == 0) {
    for (int st = 0; st < kG3Stages; ++st) {
      mbar_init(full + st, 1);
      mbar_init(split_done + st, 4);
      mbar_init(empty + st, 1);
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(kG3TmemCols));
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
        const int st = c % kG3Stages;
        const uint32_t ph = (c / kG3Stages) & 1;
        mbar_wait(empty + st, ph ^ 1);
        mbar_expect_tx(full + st, kG3AStage + kG3BStage);
        unsigned char* base = smem + static_cast<size_t>(st) * kG3StageBytes;
        for (int kb = 0; kb < 2; ++kb) tma_load_2d(base + kb * kG3ABlock, &map_a, c * kG3KC + kb * 32, r0, full + st);
        for (int kb = 0; kb < 2; ++kb)
          tma_load_2d(base + 2 * kG3AStage + kb * kG3BBlock, mb, c * kG3KC + kb * 32, brow, full + st);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: all 32 lanes run the loop, the instructions are predicated on the elected lane =====
    const uint32_t leader = g3_elect_one();
    for (int c = 0; c < chunks; ++c) {
      const int st = c % kG3Stages;
      const uint32_t ph = (c / kG3Stages) & 1;
      mbar_wait(split_done + st, ph);
      tc_fence_after();
      const uint32_t base = smem_u32(smem + static_cast<size_t>(st) * kG3StageBytes);
      const uint64_t a_hi = umma_desc(base);
      const uint64_t a_lo = umma_desc(base + kG3AStage);
      const uint64_t b_hi = umma_desc(base + 2 * kG3AStage);
      const uint64_t b_lo = umma_desc(base + 2 * kG3AStage + kG3BStage);
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {                    // 4 x (K = 8 tf32 = 32 bytes) inside one swizzle row
          const uint64_t oa = static_cast<uint64_t>(kb) * (kG3ABlock >> 4) + k4 * 2;
          const uint64_t ob = static_cast<uint64_t>(kb) * (kG3BBlock >> 4) + k4 * 2;
          const uint32_t acc = (c | kb | k4) != 0 ? 1u : 0u;
          g3_mma_tf32_if(leader, tmem_base, a_hi + oa, b_hi + ob, kG3Idesc, acc);
          g3_mma_tf32_if(leader, tmem_base + kG3N, a_lo + oa, b_hi + ob, kG3Idesc, acc);
          g3_mma_tf32_if(leader, tmem_base + kG3N, a_hi + oa, b_lo + ob, kG3Idesc, 1u);
        }
      }
      g3_commit_if(leader, empty + st);                     // the stage may be refilled once these MMAs have read it
    }
    g3_commit_if(leader, acc_full);
  } else {
    // ===== split (per stage), then epilogue =====
    const int t = threadIdx.x - 64;                         // 0..127
    for (int c = 0; c < chunks; ++c) {
      const int st = c % kG3Stages;
      const uint32_t ph = (c / kG3Stages) & 1;
      mbar_wait(full + st, ph);
      unsigned char* base = smem + static_cast<size_t>(st) * kG3StageBytes;
      float4* a_hi = reinterpret_cast<float4*>(base);
      float4* a_lo = reinterpret_cast<float4*>(base + kG3AStage);
      float4* b_hi = reinterpret_cast<float4*>(base + 2 * kG3AStage);
      float4* b_lo = reinterpret_cast<float4*>(base + 2 * kG3AStage + kG3BStage);
#pragma unroll 4
      for (int i = t; i < static_cast<int>(kG3AStage / 16); i += 128) {
        float4 hi, lo;
        g3_split(a_hi[i], hi, lo);
        a_hi[i] = hi;
        a_lo[i] = lo;
      }
#pragma unroll 4
      for (int i = t; i < static_cast<int>(kG3BStage / 16); i += 128) {
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
#pragma unroll 1
    for (int cc = 0; cc < kG3N / 32; ++cc) {
      uint32_t vm[32], vc[32];
      __syncwarp();
      tc_ld32x2(t_addr + static_cast<uint32_t>(cc * 32), vm, vc, static_cast<uint32_t>(kG3N));
      if (row < M) {
        float* o = P.out + static_cast<size_t>(row) * P.ldo + n0 + cc * 32;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float4 r;
          r.x = __uint_as_float(vm[i]) + __uint_as_float(vc[i]);
          r.y = __uint_as_float(vm[i + 1]) + __uint_as_float(vc[i + 1]);
          r.z = __uint_as_float(vm[i + 2]) + __uint_as_float(vc[i + 2]);
          r.w = __uint_as_float(vm[i + 3]) + __uint_as_float(vc[i + 3]);
          if (P.bias != nullptr) {
            const float4 b = *reinterpret_cast<const float4*>(P.bias + n0 + cc * 32 + i);
            r.x += b.x; r.y += b.y; r.z += b.z; r.w += b.w;
          }
          *reinterpret_cast<float4*>(o + i) = r;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kG3TmemCols));
  }
}

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

bool rows_gemm_tc_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PSB_ENC_TC");
    v = (e != nullptr && atoi(e) != 0) ? 1 : 0;
  }
  return v == 1;
}

bool rows_gemm_tc_supported(const float* A, int lda, int K, const float* Bt0, const float* Bt1, int split, int J,
                            const float* bias, const float* out, int ldo) {
  if (K <= 0 || K % kG3KC != 0 || J <= 0 || J % kG3N != 0 || (lda & 3) != 0 || (ldo & 3) != 0) return false;
  if (Bt1 != nullptr && (split <= 0 || split >= J || split % kG3N != 0)) return false;
  return A != nullptr && Bt0 != nullptr && out != nullptr && !misaligned16(A) && !misaligned16(Bt0) &&
         !misaligned16(Bt1) && !misaligned16(bias) && !misaligned16(out);
}

int launch_rows_gemm_tc(const float* A, int lda, const int32_t* m_dev, int m_host, int m_max, int K, const float* Bt0,
                        const float* Bt1, int split, int J, const float* bias, float* out, int ldo, cudaStream_t s) {
  if (!rows_gemm_tc_supported(A, lda, K, Bt0, Bt1, split, J, bias, out, ldo)) return PSB_E_UNSUPPORTED;
  if (m_max <= 0) return PSB_OK;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(gemm3_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(kG3Smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    attr_done = true;
  }
  alignas(64) CUtensorMap map_a, map_b0, map_b1;
  const int rows0 = Bt1 != nullptr ? split : J;
  int st;
  if ((st = g3_make_map(&map_a, A, m_max, K, lda, kG3M)) != PSB_OK) return st;
  if ((st = g3_make_map(&map_b0, Bt0, rows0, K, K, kG3N)) != PSB_OK) return st;
  if (Bt1 != nullptr) {
    if ((st = g3_make_map(&map_b1, Bt1, J - split, K, K, kG3N)) != PSB_OK) return st;
  } else {
    map_b1 = map_b0;
  }
  G3Params P;
  P.m_dev = m_dev;
  P.m_host = m_host;
  P.K = K;
  P.J = J;
  P.split = rows0;
  P.bias = bias;
  P.out = out;
  P.ldo = ldo;
  const dim3 grid(static_cast<unsigned>((m_max + kG3M - 1) / kG3M), static_cast<unsigned>(J / kG3N));
  PSB_PROF("gemm3_tf32_kernel", s);
  gemm3_tf32_kernel<<<grid, kG3Threads, kG3Smem, s>>>(map_a, map_b0, map_b1, P);
  return launch_status();
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
