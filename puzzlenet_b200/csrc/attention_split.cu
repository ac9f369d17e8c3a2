// Offset-attention core of the SPLIT path (PZ_PREC_SPLIT) on tcgen05: per cloud (L = 256 tokens, d_k = 64, C = 256)
//   r = x - softmax(q k^T / sqrt(d_k)) v                      (model5_b.py:67-75 and :98)
// with every operand as fp16 hi / lo planes and every product as three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulation):
// the logits, the probabilities and P v all keep ~22 mantissa bits, so the attention map and r match the fp32 reference
// to ~1e-6 -- the bf16 kernel's map is off by 0.14 of its maximum with peaky logits.
//
// One CTA per (cloud, block of 128 queries): 2 CTAs per cloud.  S = q k^T lands with the query rows on the TMEM lanes,
// so each of the 128 softmax threads owns one full row (max / exp2 / sum in-thread).  The un-normalised probabilities go
// back to shared memory as the K-major A operand of O = P v (both planes), whose B operand v^T (written transposed by
// the v projection's epilogue) streams through a 3-stage ring of [128 channels x 64 keys] x 2 planes while the softmax
// runs; O is produced as two 128-channel halves so the epilogue of the first overlaps the MMAs of the second.
// Warps 0-3 softmax + epilogue, warp 4 the MMA-issuing thread, warps 5-7 the v^T producers.
// Shared memory: region A 128 KB (q, k planes -> P planes) + ring 96 KB.  TMEM: S [0,256), O halves [256,384), [384,512).
#include <cuda_fp16.h>

#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int AS_THREADS = 256;
constexpr int AS_L = 256, AS_C = 256;
constexpr uint32_t T16 = 128 * 128;                 // one [128 x 64] fp16 tile
constexpr uint32_t REGA = 8 * T16;                  // 128 KB
constexpr int AS_NST = 3;
constexpr uint32_t AS_STAGE = 2 * T16;              // hi + lo of [128 ch x 64 keys]
constexpr int AS_PROD = 96;

__host__ __device__ constexpr uint32_t idesc_f16(int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void split2h(float a, float b, uint32_t& hi, uint32_t& lo) {
  uint32_t h;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(b), "f"(a));
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&h));
  uint32_t l;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(b - hf.y), "f"(a - hf.x));
  hi = h;
  lo = l;
}
__device__ __forceinline__ void umma3(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b_hi, uint64_t b_lo, uint32_t idesc,
                                      uint32_t acc) {
  umma_bf16(d, a_hi, b_hi, idesc, acc);
  umma_bf16(d, a_hi, b_lo, idesc, 1);
  umma_bf16(d, a_lo, b_hi, idesc, 1);
}
}  // namespace

__global__ void __launch_bounds__(AS_THREADS, 1) attention_split_kernel(const AttnSplit p) {
  extern __shared__ __align__(1024) uint8_t as_smem_raw[];
  const uint32_t base = (smem_u32(as_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = as_smem_raw + (base - smem_u32(as_smem_raw));
  const uint32_t ring = base + REGA;
  const uint32_t bars = ring + AS_NST * AS_STAGE;
  const uint32_t bar_s = bars, bar_p = bars + 8, bar_o = bars + 16 /* 2 */, full_bar = bars + 32, empty_bar = full_bar + 8 * AS_NST;
  const uint32_t tmem_slot = empty_bar + 8 * AS_NST;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.x >> 1, qb = blockIdx.x & 1;
  const size_t row0 = (size_t)cloud * AS_L;
  const __half* qk_hi = static_cast<const __half*>(p.qk_hi);
  const __half* qk_lo = static_cast<const __half*>(p.qk_lo);
  const __half* vT_hi = static_cast<const __half*>(p.vT_hi) + row0 * AS_L;
  const __half* vT_lo = static_cast<const __half*>(p.vT_lo) + row0 * AS_L;

  // ---- q (this block's 128 rows) and k (all 256 rows), both planes -> region A
  const uint32_t q_hi_s = base, q_lo_s = base + T16, k_hi_s = base + 2 * T16, k_lo_s = base + 4 * T16;
  for (int id = tid; id < 128 * 8; id += AS_THREADS) {
    const int c = id & 7, i = id >> 3;
    const size_t off = (row0 + qb * 128 + i) * 128 + c * 8;
    cp_async16(q_hi_s + sw128(i, c), qk_hi + off);
    cp_async16(q_lo_s + sw128(i, c), qk_lo + off);
  }
  for (int id = tid; id < 256 * 8; id += AS_THREADS) {
    const int c = id & 7, i = id >> 3;
    const size_t off = (row0 + i) * 128 + 64 + c * 8;
    cp_async16(k_hi_s + sw128(i, c), qk_hi + off);
    cp_async16(k_lo_s + sw128(i, c), qk_lo + off);
  }
  cp_async_commit();
  if (tid == 0) {
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_o + 8, 1);
    for (int s = 0; s < AS_NST; ++s) {
      mbar_init(full_bar + 8 * s, AS_PROD);
      mbar_init(empty_bar + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (warp >= 5) {
    // =========================================================== v^T producers: 8 stage loads (2 channel halves x 4 key blocks)
    const int pt = tid - 5 * 32;
    uint32_t arrived = 0;
    for (uint32_t it = 0; it < 8; ++it) {
      const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
      const int h = it >> 2, kb = it & 3;
      mbar_wait(empty_bar + 8 * s, ph ^ 1);
      const uint32_t st = ring + s * AS_STAGE;
      for (int id = pt; id < 128 * 8; id += AS_PROD) {
        const int c = id & 7, r = id >> 3;
        const size_t off = (size_t)(h * 128 + r) * AS_L + kb * 64 + c * 8;
        cp_async16(st + sw128(r, c), vT_hi + off);
        cp_async16(st + T16 + sw128(r, c), vT_lo + off);
      }
      cp_async_commit();
      if (it - arrived >= 1) {       // keep two stage loads in flight
        cp_async_wait<1>();
        fence_proxy_async();
        mbar_arrive(full_bar + 8 * (arrived % AS_NST));
        ++arrived;
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (arrived < 8) {
      mbar_arrive(full_bar + 8 * (arrived % AS_NST));
      ++arrived;
    }
  } else if (warp == 4) {
    // =========================================================== MMA issuer
    if (lane == 0) {
      {  // S[i, j] = sum_d q[i, d] k[j, d]
        const uint32_t idesc = idesc_f16(256);
        const uint64_t a_hi = make_desc(q_hi_s), a_lo = make_desc(q_lo_s), b_hi = make_desc(k_hi_s), b_lo = make_desc(k_lo_s);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) umma3(tmem, a_hi + 2 * k4, a_lo + 2 * k4, b_hi + 2 * k4, b_lo + 2 * k4, idesc, k4 != 0);
        umma_commit(bar_s);
      }
      mbar_wait(bar_p, 0);          // both planes of P are in region A
      tc_fence_after();
      const uint32_t idesc = idesc_f16(128);
      for (uint32_t it = 0; it < 8; ++it) {
        const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
        const int h = it >> 2, kb = it & 3;
        mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = ring + s * AS_STAGE;
        const uint64_t a_hi = make_desc(base + kb * T16), a_lo = make_desc(base + 4 * T16 + kb * T16);
        const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + T16);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)
          umma3(tmem + 256 + h * 128, a_hi + 2 * k4, a_lo + 2 * k4, b_hi + 2 * k4, b_lo + 2 * k4, idesc, (kb | k4) != 0);
        umma_commit(empty_bar + 8 * s);
        if (kb == 3) umma_commit(bar_o + 8 * h);
      }
    }
  } else {
    // =========================================================== softmax + epilogue: thread = query row
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16);
    const int prow = warp * 32 + lane;
    const size_t grow = row0 + qb * 128 + prow;
    const float cexp = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(d_k)
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float m = -INFINITY;
#pragma unroll 1
    for (int c32 = 0; c32 < 8; ++c32) {
      float v[32];
      tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
    }
    const float mc = m * cexp;
    float sum = 0.f;
#pragma unroll 1
    for (int c32 = 0; c32 < 8; ++c32) {
      float v[32];
      tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = exp2f(fmaf(v[i], cexp, -mc));
        sum += v[i];
      }
      // un-normalised probabilities -> K-major operand planes: k-block c32 / 2, chunks (c32 & 1) * 4 .. +3
      const uint32_t pk = base + (c32 >> 1) * T16;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 oh, ol;
        split2h(v[q4 * 8 + 0], v[q4 * 8 + 1], oh.x, ol.x);
        split2h(v[q4 * 8 + 2], v[q4 * 8 + 3], oh.y, ol.y);
        split2h(v[q4 * 8 + 4], v[q4 * 8 + 5], oh.z, ol.z);
        split2h(v[q4 * 8 + 6], v[q4 * 8 + 7], oh.w, ol.w);
        const uint32_t off = sw128(prow, (c32 & 1) * 4 + q4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pk + off), "r"(oh.x), "r"(oh.y), "r"(oh.z), "r"(oh.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pk + 4 * T16 + off), "r"(ol.x), "r"(ol.y), "r"(ol.z), "r"(ol.w) : "memory");
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_p);
    const float inv = 1.0f / sum;
    if (p.attn_mode != 0) {  // attention map (need=True): mean of the four layers' maps, model5_b.py:468-469
      float* ag = p.attn + grow * AS_L;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 a;
          a.x = exp2f(fmaf(v[q4 * 4 + 0], cexp, -mc)) * inv;
          a.y = exp2f(fmaf(v[q4 * 4 + 1], cexp, -mc)) * inv;
          a.z = exp2f(fmaf(v[q4 * 4 + 2], cexp, -mc)) * inv;
          a.w = exp2f(fmaf(v[q4 * 4 + 3], cexp, -mc)) * inv;
          float4* dst = reinterpret_cast<float4*>(ag + c32 * 32 + q4 * 4);
          if (p.attn_mode != 1) {
            const float4 o = *dst;
            a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
            if (p.attn_mode == 3) { a.x *= 0.25f; a.y *= 0.25f; a.z *= 0.25f; a.w *= 0.25f; }
          }
          *dst = a;
        }
      }
    }
    // r = x - O / sum, one 128-channel half at a time (the second half's MMAs run under the first half's epilogue)
    const __half* xh = static_cast<const __half*>(p.x_hi) + grow * p.ldx;
    const __half* xl = static_cast<const __half*>(p.x_lo) + grow * p.ldx;
    __half* rh = static_cast<__half*>(p.r_hi) + grow * AS_C;
    __half* rl = static_cast<__half*>(p.r_lo) + grow * AS_C;
    for (int h = 0; h < 2; ++h) {
      mbar_wait(bar_o + 8 * h, 0);
      tc_fence_after();
#pragma unroll 1
      for (int c32 = 0; c32 < 4; ++c32) {
        const int cb = h * 128 + c32 * 32;
        uint4 xhv[4], xlv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          xhv[q4] = *reinterpret_cast<const uint4*>(xh + cb + q4 * 8);
          xlv[q4] = *reinterpret_cast<const uint4*>(xl + cb + q4 * 8);
        }
        float v[32];
        tmem_ld32(t_row + 256 + cb, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t* hp = reinterpret_cast<const uint32_t*>(&xhv[q4]);
          const uint32_t* lp = reinterpret_cast<const uint32_t*>(&xlv[q4]);
          uint4 oh, ol;
          uint32_t* ohp = reinterpret_cast<uint32_t*>(&oh);
          uint32_t* olp = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&hp[e]));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&lp[e]));
            split2h((a.x + b.x) - v[q4 * 8 + 2 * e] * inv, (a.y + b.y) - v[q4 * 8 + 2 * e + 1] * inv, ohp[e], olp[e]);
          }
          *reinterpret_cast<uint4*>(rh + cb + q4 * 8) = oh;
          *reinterpret_cast<uint4*>(rl + cb + q4 * 8) = ol;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

int launch_attention_split(const AttnSplit& p, int clouds, cudaStream_t st) {
  PZ_REQUIRE(p.qk_hi && p.qk_lo && p.vT_hi && p.vT_lo && p.x_hi && p.x_lo && p.r_hi && p.r_lo, PZ_ERR_ARG, "attention_split: null pointer");
  PZ_REQUIRE(p.ldx % 8 == 0 && (((uintptr_t)p.x_hi | (uintptr_t)p.x_lo | (uintptr_t)p.r_hi | (uintptr_t)p.r_lo | (uintptr_t)p.qk_hi |
                                 (uintptr_t)p.qk_lo | (uintptr_t)p.vT_hi | (uintptr_t)p.vT_lo) & 15) == 0,
             PZ_ERR_ARG, "attention_split: rows must be 16-byte aligned");
  PZ_REQUIRE(p.attn_mode == 0 || p.attn, PZ_ERR_ARG, "attention_split: attention map requested without a buffer");
  const size_t smem = 1024 + REGA + AS_NST * AS_STAGE + 8 * (4 + 2 * AS_NST) + 32;
  static_assert(1024 + REGA + AS_NST * AS_STAGE + 8 * (4 + 2 * AS_NST) + 32 <= 232448, "attention_split: shared memory budget");
  PZ_CUDA(cudaFuncSetAttribute(attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_split_kernel<<<2 * clouds, AS_THREADS, smem, st>>>(p);
  PZ_LAUNCH_CHECK();
  return 0;
}

}  // namespace pz
