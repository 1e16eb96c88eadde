// BERT cross-encoder forward (ms-marco-MiniLM-L-12 architecture) for sm_100a: the arithmetic behind
// rerank_chunks (reference _search.py:364-397 -> rerankers FlashRankRanker -> onnxruntime, all
// third-party).  Variable-length packed batches (no padding): tokens [T, H], cu_seqlens [P + 1].
//
//   embed_ln_kernel      word + position + token-type embeddings, LayerNorm          (fp32 math, fp16 out)
//   linear_tcgen05_kernel  Y = act(X W^T + b): tcgen05.mma (M=128 tokens, N<=256 outputs per pass, K
//                        sliced by 64), X copied into 128B-swizzled smem by loader warps, W as a
//                        pre-swizzled fp16 image fetched with cp.async.bulk, fp32 accumulate in TMEM,
//                        bias / GELU(erf) fused in the TMEM epilogue
//   attention_kernel     softmax(Q K^T / sqrt(dh)) V per (sequence, head), fp32 math
//   add_ln_kernel        LayerNorm(x + residual)
//   cls_head_kernel      pooler (dense + tanh on [CLS]) -> classifier -> logit, sigmoid score
#include <cuda_fp16.h>

#include "common.cuh"
#include "tcgen05_ptx.cuh"

namespace rl {
namespace {

using namespace tc;

constexpr int kTileM = 128;
constexpr int kSliceK = 64;
constexpr int kMaxN = 256;
constexpr int kNumEpiWarps = 4;
constexpr int kMmaWarp = 4;
constexpr int kWWarp = 5;
constexpr int kFirstLoaderWarp = 6;
constexpr int kNumLoaderWarps = 8;
constexpr int kThreads = (kFirstLoaderWarp + kNumLoaderWarps) * 32;
constexpr int kMaxStages = 8;
constexpr int kABytes = kTileM * 128;
constexpr uint32_t kSmemBudget = 220 * 1024;

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  const __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- weight image: W[N, K] fp32 row-major -> per (pass, k-slice) swizzled fp16 UMMA B tiles -------------
__host__ __device__ inline int pass_rows(int N, int pass) {
  const int rem = N - pass * kMaxN;
  return rem < kMaxN ? rem : kMaxN;
}
__host__ __device__ inline size_t pass_offset_halves(int N, int K, int pass) {
  const int n_ks = (K + kSliceK - 1) / kSliceK;
  return (size_t)pass * kMaxN * n_ks * kSliceK;  // full passes precede; only the last pass is short
}

__global__ void pack_linear_kernel(const float* __restrict__ W, int N, int K, __half* __restrict__ img) {
  const int n_ks = (K + kSliceK - 1) / kSliceK;
  const int64_t total = (int64_t)((N + 15) / 16 * 16) * n_ks * kSliceK;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(idx / (n_ks * kSliceK));
    const int kk = (int)(idx % (n_ks * kSliceK));
    const int pass = n / kMaxN, r = n % kMaxN;
    const int nb = (pass_rows(N, pass) + 15) / 16 * 16;
    const int ks = kk / kSliceK, e = kk % kSliceK;
    const float v = (n < N && kk < K) ? W[(size_t)n * K + kk] : 0.f;
    const size_t off = pass_offset_halves(N, K, pass) + ((size_t)ks * nb + r) * kSliceK +
                       (size_t)((((e >> 3) ^ (r & 7)) << 3) + (e & 7));
    img[off] = __float2half_rn(v);
  }
}

// ---- tcgen05 linear layer ---------------------------------------------------------------------------------
struct LinArgs {
  const __half* X;     // [T, K]
  const __half* img;   // packed weights
  const float* bias;   // [N]
  __half* Y;           // [T, N]
  int T, N, K, act;    // act: 0 none, 1 GELU(erf)
  int n_pass, n_ks, stages;
};

__host__ __device__ inline uint32_t lin_stage_bytes() { return kABytes + kMaxN * 128u; }

__global__ void __launch_bounds__(kThreads, 1) linear_tcgen05_kernel(const LinArgs t) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const uint32_t sbytes = lin_stage_bytes();
  uint64_t* full = reinterpret_cast<uint64_t*>(base + (size_t)t.stages * sbytes);
  uint64_t* empty = full + kMaxStages;
  uint64_t* tmem_full = empty + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (t.T + kTileM - 1) / kTileM;
  const int64_t n_items = (int64_t)m_tiles * t.n_pass;  // item = m_tile * n_pass + pass
  const int64_t first = blockIdx.x, stride = gridDim.x;
  const int64_t my_items = first < n_items ? (n_items - first + stride - 1) / stride : 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < t.stages; ++i) {
      mbar_init(&full[i], kNumLoaderWarps + 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], kNumEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_ptr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp >= kFirstLoaderWarp) {
    // activations: fp16 rows -> swizzled K-major smem tile (UMMA A), 4 x 16-byte chunks per thread and slice
    const int lt = threadIdx.x - kFirstLoaderWarp * 32;
    const int j = lt & 7, r0 = lt >> 3;  // chunk j of rows r0 + 32 i
    int stage = 0;
    uint32_t phase = 0;
    for (int64_t it = 0; it < my_items; ++it) {
      const int64_t item = first + it * stride;
      const int m_tile = (int)(item / t.n_pass);
      for (int ks = 0; ks < t.n_ks; ++ks) {
        uint4 v[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int row = m_tile * kTileM + r0 + 32 * i;
          const int col = ks * kSliceK + j * 8;
          v[i] = make_uint4(0u, 0u, 0u, 0u);
          if (row < t.T && col < t.K) v[i] = __ldg(reinterpret_cast<const uint4*>(t.X + (size_t)row * t.K + col));
        }
        mbar_wait(&empty[stage], phase ^ 1u);
        unsigned char* A = base + (size_t)stage * sbytes;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = r0 + 32 * i;
          *reinterpret_cast<uint4*>(A + (uint32_t)r * 128u + (((uint32_t)j ^ ((uint32_t)r & 7u)) << 4)) = v[i];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&full[stage]);
        if (++stage == t.stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == kWWarp) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t it = 0; it < my_items; ++it) {
        const int64_t item = first + it * stride;
        const int pass = (int)(item % t.n_pass);
        const int nb = (pass_rows(t.N, pass) + 15) / 16 * 16;
        const uint32_t wbytes = (uint32_t)nb * 128u;
        const __half* src = t.img + pass_offset_halves(t.N, t.K, pass);
        for (int ks = 0; ks < t.n_ks; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full[stage], wbytes);
          bulk_g2s(base + (size_t)stage * sbytes + kABytes, src + (size_t)ks * nb * kSliceK, wbytes, &full[stage]);
          if (++stage == t.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int64_t it = 0; it < my_items; ++it) {
        const int64_t item = first + it * stride;
        const int pass = (int)(item % t.n_pass);
        const int nb = (pass_rows(t.N, pass) + 15) / 16 * 16;
        const uint32_t idesc = make_idesc_f16(kTileM, nb);
        const int buf = (int)(it & 1);
        mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * kMaxN);
        for (int ks = 0; ks < t.n_ks; ++ks) {
          mbar_wait(&full[stage], phase);
          fence_proxy_async();
          tc_fence_after();
          const uint32_t a_addr = smem_u32(base + (size_t)stage * sbytes);
          const uint64_t a_desc = make_kmajor_sw128_desc(a_addr);
          const uint64_t b_desc = make_kmajor_sw128_desc(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kSliceK / 16; ++k)
            umma_f16(d_tmem, a_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), idesc, (ks | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == t.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tmem_full[buf]);
      }
    }
  } else {
    // epilogue: TMEM -> + bias -> activation -> fp16 -> global (each thread owns one token row)
    const int q = warp;
    for (int64_t it = 0; it < my_items; ++it) {
      const int64_t item = first + it * stride;
      const int m_tile = (int)(item / t.n_pass), pass = (int)(item % t.n_pass);
      const int nb = pass_rows(t.N, pass);
      const int n0 = pass * kMaxN;
      const int buf = (int)(it & 1);
      const int row = m_tile * kTileM + q * 32 + lane;
      mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * kMaxN);
      auto store_chunk = [&](int c0, const uint32_t (&v)[32]) {
        if (row < t.T) {
          uint32_t packed[16];
#pragma unroll
          for (int jj = 0; jj < 16; ++jj) {
            float x0 = __uint_as_float(v[2 * jj]) + __ldg(t.bias + n0 + c0 + 2 * jj);
            float x1 = __uint_as_float(v[2 * jj + 1]) + __ldg(t.bias + n0 + c0 + 2 * jj + 1);
            if (t.act == 1) {
              x0 = 0.5f * x0 * (1.f + erff(x0 * 0.70710678118654752f));
              x1 = 0.5f * x1 * (1.f + erff(x1 * 0.70710678118654752f));
            }
            packed[jj] = pack_half2(x0, x1);
          }
          uint4* dst = reinterpret_cast<uint4*>(t.Y + (size_t)row * t.N + n0 + c0);
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
            dst[jj] = make_uint4(packed[4 * jj], packed[4 * jj + 1], packed[4 * jj + 2], packed[4 * jj + 3]);
        }
      };
      uint32_t va[32], vb[32];
      tmem_ld32_async(taddr0, va);
      tmem_ld_wait(va);
      for (int c0 = 0; c0 < nb; c0 += 64) {
        const bool has_b = c0 + 32 < nb;
        if (has_b) tmem_ld32_async(taddr0 + (uint32_t)(c0 + 32), vb);
        store_chunk(c0, va);
        if (has_b) {
          tmem_ld_wait(vb);
          if (c0 + 64 < nb) tmem_ld32_async(taddr0 + (uint32_t)(c0 + 64), va);
          store_chunk(c0 + 32, vb);
          if (c0 + 64 < nb) tmem_ld_wait(va);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[buf]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---- embeddings + LayerNorm: one warp per token ----------------------------------------------------------
__global__ void __launch_bounds__(256) embed_ln_kernel(const int32_t* __restrict__ ids, const int32_t* __restrict__ type_ids,
                                                       const int32_t* __restrict__ pos_ids, const __half* __restrict__ word,
                                                       const __half* __restrict__ pos, const __half* __restrict__ type,
                                                       const float* __restrict__ g, const float* __restrict__ bta, float eps,
                                                       int T, int H, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= T) return;
  const __half* w = word + (size_t)ids[tok] * H;
  const __half* p = pos + (size_t)pos_ids[tok] * H;
  const __half* ty = type + (size_t)type_ids[tok] * H;
  float x[16];  // H <= 512
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    x[i] = c < H ? __half2float(w[c]) + __half2float(p[c]) + __half2float(ty[c]) : 0.f;
    s += x[i];
  }
  const float mean = warp_sum_f(s) / (float)H;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    if (c < H) var += (x[i] - mean) * (x[i] - mean);
  }
  const float rstd = rsqrtf(warp_sum_f(var) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    if (c < H) out[(size_t)tok * H + c] = __float2half_rn((x[i] - mean) * rstd * g[c] + bta[c]);
  }
}

// out = LayerNorm(x + res), one warp per token.
__global__ void __launch_bounds__(256) add_ln_kernel(const __half* __restrict__ xin, const __half* __restrict__ res,
                                                     const float* __restrict__ g, const float* __restrict__ bta, float eps,
                                                     int T, int H, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= T) return;
  float x[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    x[i] = c < H ? __half2float(xin[(size_t)tok * H + c]) + __half2float(res[(size_t)tok * H + c]) : 0.f;
    s += x[i];
  }
  const float mean = warp_sum_f(s) / (float)H;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    if (c < H) var += (x[i] - mean) * (x[i] - mean);
  }
  const float rstd = rsqrtf(warp_sum_f(var) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = lane + 32 * i;
    if (c < H) out[(size_t)tok * H + c] = __float2half_rn((x[i] - mean) * rstd * g[c] + bta[c]);
  }
}

// ---- attention: one block per (sequence, head), flash-style on mma.sync tensor cores ---------------------
// qkv [T, 3H] (Q | K | V), ctx [T, H].  head_dim must be 32 (MiniLM-L12-H384: 12 heads x 32).
// K and V of the head sit in shared memory (80-byte row pitch: conflict-free ldmatrix); each warp
// owns 16-query blocks: S = Q K^T with m16n8k16 (fp16 in, fp32 acc), online softmax in the exp2
// domain, O += P V with P re-packed from the S accumulators as the A operand.
constexpr int kAttPitch = 40;  // halves

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}

__global__ void __launch_bounds__(128) attention_kernel(const __half* __restrict__ qkv, const int32_t* __restrict__ cu,
                                                        int H, int n_heads, float scale_log2e, __half* __restrict__ ctx) {
  extern __shared__ __align__(16) unsigned char att_smem[];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int t0 = cu[seq], L = cu[seq + 1] - t0;
  const int Lp = (L + 63) / 64 * 64;
  __half* Ks = reinterpret_cast<__half*>(att_smem);
  __half* Vs = Ks + (size_t)Lp * kAttPitch;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t ld = (size_t)3 * H;
  for (int idx = threadIdx.x; idx < Lp * 4; idx += blockDim.x) {
    const int j = idx >> 2, c = idx & 3;
    uint4 kv = make_uint4(0u, 0u, 0u, 0u), vv = kv;
    if (j < L) {
      const __half* base = qkv + (size_t)(t0 + j) * ld + head * 32 + c * 8;
      kv = __ldg(reinterpret_cast<const uint4*>(base + H));
      vv = __ldg(reinterpret_cast<const uint4*>(base + 2 * H));
    }
    *reinterpret_cast<uint4*>(Ks + (size_t)j * kAttPitch + c * 8) = kv;
    *reinterpret_cast<uint4*>(Vs + (size_t)j * kAttPitch + c * 8) = vv;
  }
  __syncthreads();
  const int r = lane >> 2, cp = (lane & 3) * 2;
  for (int qb = warp; qb * 16 < L; qb += 4) {
    const int q0 = qb * 16 + r, q1 = q0 + 8;
    uint32_t a[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __half* p0 = qkv + (size_t)(t0 + q0) * ld + head * 32 + ks * 16 + cp;
      const __half* p1 = qkv + (size_t)(t0 + q1) * ld + head * 32 + ks * 16 + cp;
      a[ks][0] = q0 < L ? __ldg(reinterpret_cast<const uint32_t*>(p0)) : 0u;
      a[ks][1] = q1 < L ? __ldg(reinterpret_cast<const uint32_t*>(p1)) : 0u;
      a[ks][2] = q0 < L ? __ldg(reinterpret_cast<const uint32_t*>(p0 + 8)) : 0u;
      a[ks][3] = q1 < L ? __ldg(reinterpret_cast<const uint32_t*>(p1 + 8)) : 0u;
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float O[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int e = 0; e < 4; ++e) O[i][e] = 0.f;
    for (int kb = 0; kb < Lp; kb += 64) {
      float S[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) S[j][e] = 0.f;
        uint32_t b[4];
        ldsm_x4(b, Ks + (size_t)(kb + j * 8 + (lane & 7)) * kAttPitch + (lane >> 3) * 8);
        mma16816(S[j], a[0], b[0], b[1]);
        mma16816(S[j], a[1], b[2], b[3]);
      }
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb + j * 8 + cp + (e & 1);
          const float v = key < L ? S[j][e] * scale_log2e : -INFINITY;
          S[j][e] = v;
          if (e < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
        }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);   // finite: every key block holds a valid key
      const float c0 = exp2f(m0 - mn0), c1 = exp2f(m1 - mn1);
      l0 *= c0; l1 *= c1;
#pragma unroll
      for (int i = 0; i < 4; ++i) { O[i][0] *= c0; O[i][1] *= c0; O[i][2] *= c1; O[i][3] *= c1; }
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pexp = exp2f(S[j][e] - (e < 2 ? mn0 : mn1));
          S[j][e] = pexp;
          if (e < 2) l0 += pexp; else l1 += pexp;
        }
      m0 = mn0; m1 = mn1;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_half2(S[2 * kk][0], S[2 * kk][1]);
        pa[1] = pack_half2(S[2 * kk][2], S[2 * kk][3]);
        pa[2] = pack_half2(S[2 * kk + 1][0], S[2 * kk + 1][1]);
        pa[3] = pack_half2(S[2 * kk + 1][2], S[2 * kk + 1][3]);
#pragma unroll
        for (int dn2 = 0; dn2 < 2; ++dn2) {
          uint32_t vb[4];
          ldsm_x4_trans(vb, Vs + (size_t)(kb + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kAttPitch +
                                (dn2 * 2 + (lane >> 4)) * 8);
          mma16816(O[dn2 * 2], pa, vb[0], vb[1]);
          mma16816(O[dn2 * 2 + 1], pa, vb[2], vb[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      if (q0 < L)
        *reinterpret_cast<uint32_t*>(ctx + (size_t)(t0 + q0) * H + head * 32 + dn * 8 + cp) = pack_half2(O[dn][0] * inv0, O[dn][1] * inv0);
      if (q1 < L)
        *reinterpret_cast<uint32_t*>(ctx + (size_t)(t0 + q1) * H + head * 32 + dn * 8 + cp) = pack_half2(O[dn][2] * inv1, O[dn][3] * inv1);
    }
  }
}

// ---- pooler + classifier: one warp per sequence -----------------------------------------------------------
__global__ void __launch_bounds__(128) cls_head_kernel(const __half* __restrict__ hidden, const int32_t* __restrict__ cu,
                                                       const float* __restrict__ Wp, const float* __restrict__ bp,
                                                       const float* __restrict__ Wc, const float* __restrict__ bc, int P, int H,
                                                       float* __restrict__ logit, float* __restrict__ score) {
  extern __shared__ float cls_smem[];  // [warps][H]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int seq = blockIdx.x * (blockDim.x >> 5) + warp;
  if (seq >= P) return;
  float* h = cls_smem + (size_t)warp * H;
  const __half* src = hidden + (size_t)cu[seq] * H;  // [CLS] token
  for (int c = lane; c < H; c += 32) h[c] = __half2float(src[c]);
  __syncwarp();
  float out = 0.f;
  for (int o = lane; o < H; o += 32) {
    const float* w = Wp + (size_t)o * H;
    float a = bp[o];
    for (int c = 0; c < H; ++c) a = fmaf(w[c], h[c], a);
    out += tanhf(a) * Wc[o];
  }
  out = warp_sum_f(out) + bc[0];
  if (lane == 0) {
    logit[seq] = out;
    score[seq] = 1.f / (1.f + __expf(-out));  // FlashRank: sigmoid of the single logit
  }
}

}  // namespace
}  // namespace rl

using namespace rl;

extern "C" size_t rl_xenc_linear_image_bytes(int N, int K) {
  const int n_ks = (K + kSliceK - 1) / kSliceK;
  const int n_pad = (N + 15) / 16 * 16;
  return (size_t)((n_pad + kMaxN - 1) / kMaxN) * kMaxN * n_ks * kSliceK * sizeof(__half);
}

extern "C" int rl_xenc_pack_linear(const float* W, int N, int K, void* image, void* stream) {
  RL_REQUIRE(W && image && N > 0 && K > 0, RL_EINVAL, "rl_xenc_pack_linear: bad arguments");
  RL_REQUIRE(N % 16 == 0 && K % 8 == 0, RL_EUNSUPPORTED, "rl_xenc_pack_linear: N %% 16 and K %% 8 must be 0");
  RL_CUDA_CHECK(cudaMemsetAsync(image, 0, rl_xenc_linear_image_bytes(N, K), (cudaStream_t)stream));
  pack_linear_kernel<<<1024, 256, 0, (cudaStream_t)stream>>>(W, N, K, reinterpret_cast<__half*>(image));
  RL_CUDA_CHECK(cudaGetLastError());
  return RL_OK;
}

static int launch_linear(const __half* X, const void* img, const float* bias, __half* Y, int T, int N, int K, int act,
                         int sm_count, cudaStream_t stream) {
  LinArgs t;
  t.X = X; t.img = reinterpret_cast<const __half*>(img); t.bias = bias; t.Y = Y; t.T = T; t.N = N; t.K = K; t.act = act;
  t.n_pass = (N + kMaxN - 1) / kMaxN;
  t.n_ks = (K + kSliceK - 1) / kSliceK;
  const uint32_t tail = (2 * kMaxStages + 4) * 8 + 64;
  int stages = (int)((kSmemBudget - 1024 - tail) / lin_stage_bytes());
  if (stages > kMaxStages) stages = kMaxStages;
  t.stages = stages;
  const size_t smem = (size_t)stages * lin_stage_bytes() + tail + 1024;
  RL_CUDA_CHECK(cudaFuncSetAttribute(linear_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t items = (int64_t)((T + kTileM - 1) / kTileM) * t.n_pass;
  const int grid = (int)(items < sm_count ? items : sm_count);
  linear_tcgen05_kernel<<<grid, kThreads, smem, stream>>>(t);
  RL_CUDA_CHECK(cudaGetLastError());
  return RL_OK;
}

extern "C" int rl_xenc_linear(const void* X, const void* image, const float* bias, void* Y, int T, int N, int K, int act,
                              void* stream) {
  RL_REQUIRE(X && image && bias && Y && T >= 0, RL_EINVAL, "rl_xenc_linear: bad arguments");
  RL_REQUIRE(N % 32 == 0 && K % 8 == 0, RL_EUNSUPPORTED, "rl_xenc_linear: N %% 32 and K %% 8 must be 0");
  if (T == 0) return RL_OK;
  int dev = 0, sms = 148;
  RL_CUDA_CHECK(cudaGetDevice(&dev));
  RL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  return launch_linear(reinterpret_cast<const __half*>(X), image, bias, reinterpret_cast<__half*>(Y), T, N, K, act, sms,
                       (cudaStream_t)stream);
}

extern "C" size_t rl_xenc_workspace_bytes(const rl_xenc_weights* w, int T) {
  if (w == nullptr || T < 0) return 0;
  const size_t H = (size_t)w->hidden, F = (size_t)w->ffn;
  // hidden, qkv (3H), ctx, tmp (H), ffn (F) -- fp16 rows
  return ((size_t)T * (H + 3 * H + H + H + F) * sizeof(__half) + 4096);
}

extern "C" int rl_xenc_score(const rl_xenc_weights* w, const int32_t* input_ids, const int32_t* type_ids,
                             const int32_t* pos_ids, const int32_t* cu_seqlens, int P, int T, int max_len,
                             float* out_logit, float* out_score, void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RL_REQUIRE(w && w->layers && input_ids && type_ids && pos_ids && cu_seqlens && out_logit && out_score, RL_EINVAL,
             "rl_xenc_score: null pointer");
  if (P == 0 || T == 0) return RL_OK;
  const int H = w->hidden, F = w->ffn, nh = w->n_heads;
  RL_REQUIRE(H % 32 == 0 && H <= 512 && nh > 0 && H / nh == 32, RL_EUNSUPPORTED,
             "rl_xenc_score: hidden=%d heads=%d unsupported (head_dim must be 32, hidden <= 512)", H, nh);
  RL_REQUIRE(F % 32 == 0 && max_len > 0 && max_len <= w->max_pos, RL_EUNSUPPORTED, "rl_xenc_score: bad ffn / max_len");
  RL_REQUIRE(workspace && workspace_bytes >= rl_xenc_workspace_bytes(w, T), RL_ENOSPACE, "rl_xenc_score: workspace too small");
  int dev = 0, sms = 148;
  RL_CUDA_CHECK(cudaGetDevice(&dev));
  RL_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  __half* hidden = reinterpret_cast<__half*>(workspace);
  __half* qkv = hidden + (size_t)T * H;
  __half* ctx = qkv + (size_t)T * 3 * H;
  __half* tmp = ctx + (size_t)T * H;
  __half* ffn = tmp + (size_t)T * H;
  const int tok_blocks = (T + 7) / 8;
  embed_ln_kernel<<<tok_blocks, 256, 0, stream>>>(input_ids, type_ids, pos_ids, reinterpret_cast<const __half*>(w->word_emb),
                                                  reinterpret_cast<const __half*>(w->pos_emb),
                                                  reinterpret_cast<const __half*>(w->type_emb), w->emb_ln_g, w->emb_ln_b,
                                                  w->ln_eps, T, H, hidden);
  RL_CUDA_CHECK(cudaGetLastError());
  const size_t att_smem = (size_t)((max_len + 63) / 64 * 64) * kAttPitch * 2 * sizeof(__half);
  RL_REQUIRE(att_smem <= 200 * 1024, RL_EUNSUPPORTED, "rl_xenc_score: max_len=%d too long for the attention kernel", max_len);
  RL_CUDA_CHECK(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)att_smem));
  const float scale = 1.4426950408889634f / sqrtf(32.f);  // softmax in the exp2 domain
  for (int l = 0; l < w->n_layers; ++l) {
    const rl_xenc_layer& L = w->layers[l];
    int rc = launch_linear(hidden, L.qkv_img, L.qkv_bias, qkv, T, 3 * H, H, 0, sms, stream);
    if (rc != RL_OK) return rc;
    attention_kernel<<<dim3(P, nh), 128, att_smem, stream>>>(qkv, cu_seqlens, H, nh, scale, ctx);
    RL_CUDA_CHECK(cudaGetLastError());
    rc = launch_linear(ctx, L.o_img, L.o_bias, tmp, T, H, H, 0, sms, stream);
    if (rc != RL_OK) return rc;
    add_ln_kernel<<<tok_blocks, 256, 0, stream>>>(tmp, hidden, L.ln1_g, L.ln1_b, w->ln_eps, T, H, hidden);
    RL_CUDA_CHECK(cudaGetLastError());
    rc = launch_linear(hidden, L.up_img, L.up_bias, ffn, T, F, H, 1, sms, stream);
    if (rc != RL_OK) return rc;
    rc = launch_linear(ffn, L.down_img, L.down_bias, tmp, T, H, F, 0, sms, stream);
    if (rc != RL_OK) return rc;
    add_ln_kernel<<<tok_blocks, 256, 0, stream>>>(tmp, hidden, L.ln2_g, L.ln2_b, w->ln_eps, T, H, hidden);
    RL_CUDA_CHECK(cudaGetLastError());
  }
  cls_head_kernel<<<(P + 3) / 4, 128, (size_t)4 * H * sizeof(float), stream>>>(hidden, cu_seqlens, w->pooler_w, w->pooler_b,
                                                                                w->cls_w, w->cls_b, P, H, out_logit, out_score);
  RL_CUDA_CHECK(cudaGetLastError());
  return RL_OK;
}
