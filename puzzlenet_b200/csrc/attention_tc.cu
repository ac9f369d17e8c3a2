// Offset-attention core on tcgen05 for the bf16 path: per cloud (L = 256 tokens, d_k = 64, C = 256)
//   r = x - softmax(q k^T / sqrt(d_k)) v                      (model5_b.py:67-75 and :98)
// One CTA per cloud.  S = q k^T is produced with query rows on the TMEM lanes, so each of the 128
// row-threads owns one full softmax row: max / exp2 / sum are in-thread over TMEM columns, no shuffles.
// The un-normalised probabilities go back to shared memory as the K-major bf16 A-operand of the second
// MMA, O = P v, whose B-operand is v^T -- written transposed by the q|k|v projection's epilogue
// (TcGemm::YT) so that it too is K-major.  1/sum is applied in the epilogue together with the offset
// subtraction.  TMEM: S of query block 0 in columns [0,256), block 1 in [256,512); O overwrites S in place.
#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int AL = 256, ADK = 64, ACV = 256;
constexpr uint32_t VT_BYTES = 4 * 256 * 128;        // 4 k-blocks of [256 ch x 64 keys]
constexpr uint32_t QK_BYTES = 2 * 256 * 128;        // Q tile + K tile; later P: 4 k-blocks of [128 x 64]
constexpr int ATT_THREADS = 256;
}  // namespace

__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const __nv_bfloat16* __restrict__ qk,
                                                                      const __nv_bfloat16* __restrict__ vT,
                                                                      const __nv_bfloat16* __restrict__ x, int ldx,
                                                                      __nv_bfloat16* __restrict__ r,
                                                                      float* __restrict__ attn, int attn_mode) {
  extern __shared__ __align__(1024) uint8_t att_smem_raw[];
  const uint32_t base = (smem_u32(att_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = att_smem_raw + (base - smem_u32(att_smem_raw));
  const uint32_t vt_s = base, q_s = base + VT_BYTES, k_s = q_s + 256 * 128, p_s = q_s;
  const uint32_t bar = base + VT_BYTES + QK_BYTES, tmem_slot = bar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.x;
  const size_t row0 = (size_t)cloud * AL;

  // ---- async loads: group 0 = Q and K tiles, group 1 = v^T
  for (int id = tid; id < 256 * 8; id += ATT_THREADS) {
    const int c = id & 7, i = id >> 3;
    const __nv_bfloat16* src = qk + (row0 + i) * 128 + c * 8;
    cp_async16(q_s + sw128(i, c), src);
    cp_async16(k_s + sw128(i, c), src + 64);
  }
  cp_async_commit();
  for (int id = tid; id < 4 * 256 * 8; id += ATT_THREADS) {
    const int c = id & 7, ch = (id >> 3) & 255, kb = id >> 11;
    cp_async16(vt_s + kb * (256 * 128) + sw128(ch, c), vT + (row0 + ch) * AL + kb * 64 + c * 8);
  }
  cp_async_commit();
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  cp_async_wait<1>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const uint32_t idesc = make_idesc(256);

  if (tid == 0) {  // S_qb[i, j] = sum_d q[qb*128+i, d] k[j, d]
#pragma unroll
    for (int qb = 0; qb < 2; ++qb) {
      const uint64_t ad = make_desc(q_s + qb * (128 * 128)), bd = make_desc(k_s);
#pragma unroll
      for (int k4 = 0; k4 < ADK / 16; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, idesc, k4 != 0);
    }
    umma_commit(bar);
  }
  uint32_t phase = 0;
  mbar_wait(bar, phase);
  phase ^= 1;
  tc_fence_after();

  const float cexp = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(d_k)
  for (int qb = 0; qb < 2; ++qb) {
    float inv = 0.f;
    const uint32_t t_row = tmem + ((uint32_t)(warp * 32) << 16) + qb * 256;
    const size_t grow = row0 + qb * 128 + warp * 32 + lane;   // this thread's token (warps 0-3 only)
    if (warp < 4) {
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
      const int prow = warp * 32 + lane;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          v[i] = exp2f(fmaf(v[i], cexp, -mc));
          sum += v[i];
        }
        // un-normalised probabilities -> K-major bf16 operand: k-block = c32 / 2, chunks (c32 & 1) * 4 .. +3
        uint8_t* pk = gen + (p_s - base) + (c32 >> 1) * (128 * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 o;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[q4 * 8 + 0], v[q4 * 8 + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[q4 * 8 + 2], v[q4 * 8 + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q4 * 8 + 4], v[q4 * 8 + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[q4 * 8 + 6], v[q4 * 8 + 7]);
          o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
          o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(pk + sw128(prow, (c32 & 1) * 4 + q4)) = o;
        }
      }
      inv = 1.0f / sum;
      if (attn_mode != 0) {  // attention map (need=True): mean of the four layers' maps, model5_b.py:468-469
        float* ag = attn + grow * AL;
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
            if (attn_mode != 1) {
              const float4 o = *dst;
              a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
              if (attn_mode == 3) { a.x *= 0.25f; a.y *= 0.25f; a.z *= 0.25f; a.w *= 0.25f; }
            }
            *dst = a;
          }
        }
      }
      fence_proxy_async();
    }
    if (qb == 0) {
      cp_async_wait<0>();   // v^T has landed
      fence_proxy_async();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {  // O[i, c] = sum_j P[i, j] vT[c, j]   (overwrites S_qb)
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t ad = make_desc(p_s + kb * (128 * 128)), bd = make_desc(vt_s + kb * (256 * 128));
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, idesc, (kb | k4) != 0);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (warp < 4) {  // r = x - O / sum
      const __nv_bfloat16* xr = x + grow * ldx;
      __nv_bfloat16* rr = r + grow * ACV;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        float v[32];
        uint4 xv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) xv[q4] = *reinterpret_cast<const uint4*>(xr + c32 * 32 + q4 * 8);
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv[q4]);
          uint4 o;
          uint32_t* op = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 xf = __bfloat1622float2(xp[h]);
            __nv_bfloat162 t = __floats2bfloat162_rn(xf.x - v[q4 * 8 + 2 * h] * inv, xf.y - v[q4 * 8 + 2 * h + 1] * inv);
            op[h] = *reinterpret_cast<uint32_t*>(&t);
          }
          *reinterpret_cast<uint4*>(rr + c32 * 32 + q4 * 8) = o;
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // P region and the O columns are free again
    tc_fence_after();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

int launch_attention_tc(const __nv_bfloat16* qk, const __nv_bfloat16* vT, const __nv_bfloat16* x, int ldx, int clouds,
                        __nv_bfloat16* r, float* attn, int attn_mode, cudaStream_t st) {
  PZ_REQUIRE(qk && vT && x && r, PZ_ERR_ARG, "attention_tc: null pointer");
  PZ_REQUIRE(ldx % 8 == 0 && ((uintptr_t)x & 15) == 0, PZ_ERR_ARG, "attention_tc: x rows must be 16-byte aligned");
  const size_t smem = 1024 + VT_BYTES + QK_BYTES + 64;
  PZ_CUDA(cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_tc_kernel<<<clouds, ATT_THREADS, smem, st>>>(qk, vT, x, ldx, r, attn, attn_mode);
  PZ_LAUNCH_CHECK();
  return 0;
}

}  // namespace pz
