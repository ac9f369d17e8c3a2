// One offset-attention layer (layerAttention.forward, model5_b.py:92-101) as ONE tcgen05 kernel per cloud:
//   q|k = x Wqk^T + b,  v = x Wv^T + b,  A = softmax(q k^T / sqrt(64)),  r = x - A v,  out = x + relu(r Wo^T + bo)
// for L = 256 tokens, C = 256 channels, d_k = 64 (bf16 operands, fp32 accumulation in TMEM).
// Nothing but x (in) and out (written into its 256-column slice of att_cat) touches HBM: q, k, v^T, P and r live
// in shared memory as K-major SWIZZLE_128B MMA operands; the weights stream through a 2-stage ring of
// [128 x 64] tiles.  Replaces 4 launches per layer (q|k GEMM, v^T GEMM, attention core, out-proj GEMM).
//
// Shared memory (224 KB of operands):  R_A 128 KB: x [256 tok x 256 ch] as 4 k-blocks, later v^T [256 ch x 256 tok];
//                                      R_B  64 KB: q,k tiles -> P (per query block) -> r (per query block);
//                                      R_W  32 KB: weight-tile ring.
// TMEM (512 columns): q|k accumulators (2 x 128) -> v^T (2 x 256) -> S (2 x 256) -> O (in place) -> out (in place).
// Epilogue mappings follow the operand they produce: thread = token row for q|k, S, O, out; thread = channel for v^T.
#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int FL = 256, FC = 256, FDK = 64;
constexpr uint32_t RA_BYTES = 4 * 256 * 128, RB_BYTES = 64 * 1024, WTILE = 128 * 128, RW_BYTES = 2 * WTILE;
constexpr int FT = 256;  // threads

// stream T weight tiles ([128 rows x 64 k], row stride ld elements) through the 2-stage ring; tile t's MMAs are
// issued by thread 0 once the tile has landed, and free the stage through wbar when they complete
template <class SrcFn, class MmaFn>
__device__ __forceinline__ void stream_phase(int T, int ld, SrcFn src, MmaFn mma, uint32_t rw, uint32_t wbar,
                                             uint32_t (&wuse)[2], int tid) {
  for (int t = 0; t <= T; ++t) {
    if (t < T) {
      const int st = t & 1;
      if (wuse[st] > 0) mbar_wait(wbar + 8 * st, (wuse[st] - 1) & 1);   // the stage's previous tile has been consumed
      const __nv_bfloat16* wp = src(t);
      for (int id = tid; id < 128 * 8; id += FT) {
        const int r = id >> 3, c = id & 7;
        cp_async16(rw + st * WTILE + sw128(r, c), wp + (size_t)r * ld + c * 8);
      }
      cp_async_commit();
    }
    if (t >= 1) {
      if (t < T) cp_async_wait<1>(); else cp_async_wait<0>();
      fence_proxy_async();
      __syncthreads();
      const int st = (t - 1) & 1;
      if (tid == 0) {
        tc_fence_after();
        mma(t - 1, rw + st * WTILE);
        umma_commit(wbar + 8 * st);
      }
      ++wuse[st];
    }
  }
}

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
  return o;
}
}  // namespace

__global__ void __launch_bounds__(FT, 1) attention_layer_tc_kernel(const AttnLayerTc p) {
  extern __shared__ __align__(1024) uint8_t fl_smem_raw[];
  const uint32_t base = (smem_u32(fl_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = fl_smem_raw + (base - smem_u32(fl_smem_raw));
  const uint32_t ra = base, rb = ra + RA_BYTES, rw = rb + RB_BYTES;
  const uint32_t tab_s = rw + RW_BYTES, inv_s = tab_s + 1024, wbar = inv_s + 512, dbar = wbar + 16, tmem_slot = dbar + 8;
  float* tab = reinterpret_cast<float*>(gen + (tab_s - base));
  float* invs = reinterpret_cast<float*>(gen + (inv_s - base));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, quarter = warp & 3, half = warp >> 2;
  const int cloud = blockIdx.x, set = cloud / p.clouds_per_set;
  const size_t row0 = (size_t)cloud * FL;
  const __nv_bfloat16* __restrict__ wqkv = p.wqkv[set];
  const __nv_bfloat16* __restrict__ wo = p.wo[set];
  const float* __restrict__ bqkv = p.bqkv[set];
  const float* __restrict__ bo = p.bo[set];

  // ---- x -> R_A (4 k-blocks of [256 tokens x 64 ch])
  for (int id = tid; id < 256 * 32; id += FT) {
    const int row = id >> 5, c32 = id & 31, kb = c32 >> 3, c = c32 & 7;
    cp_async16(ra + kb * (256 * 128) + sw128(row, c), p.x + (row0 + row) * p.ldx + kb * 64 + c * 8);
  }
  cp_async_commit();
  if (tid < 128) tab[tid] = bqkv[tid];
  if (tid == 0) {
    mbar_init(wbar, 1);
    mbar_init(wbar + 8, 1);
    mbar_init(dbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
  uint32_t wuse[2] = {0u, 0u};
  uint32_t dphase = 0;
  auto all_mma_done = [&]() {   // thread 0 has issued every MMA of the phase: wait for their completion
    if (tid == 0) umma_commit(dbar);
    mbar_wait(dbar, dphase);
    dphase ^= 1;
    tc_fence_after();
  };

  // ================= phase 1: q|k[tok, 0:128] = x Wqk^T   (rows on lanes; two 128-token blocks; N = 128)
  {
    const uint32_t idesc = make_idesc(128);
    stream_phase(4, FC, [&](int t) { return wqkv + t * 64; },
                 [&](int t, uint32_t wt) {
                   const uint64_t bd = make_desc(wt);
#pragma unroll
                   for (int blk = 0; blk < 2; ++blk) {
                     const uint64_t ad = make_desc(ra + t * (256 * 128) + blk * (128 * 128));
#pragma unroll
                     for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + blk * 128, ad + 2 * k4, bd + 2 * k4, idesc, (t | k4) != 0);
                   }
                 },
                 rw, wbar, wuse, tid);
    all_mma_done();
    // epilogue: warp half h owns token block h; q -> R_B[0:32K) as [256 x 64], k -> R_B[32K:64K)
    const int row = half * 128 + quarter * 32 + lane;
#pragma unroll 1
    for (int c32 = 0; c32 < 4; ++c32) {
      float v[32];
      tmem_ld32(tmem + lane_base + half * 128 + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += tab[c32 * 32 + i];
      uint8_t* dst = gen + (rb - base) + (c32 >> 1) * (256 * 128);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(dst + sw128(row, (c32 & 1) * 4 + q4)) = pack8(v + q4 * 8);
    }
    tc_fence_before();
    __syncthreads();   // q|k accumulators drained
    tc_fence_after();
  }

  // ================= phase 2: v^T[ch, tok] = Wv x^T   (channels on lanes; two 128-channel blocks; N = 256 tokens)
  {
    const uint32_t idesc = make_idesc(256);
    stream_phase(8, FC, [&](int t) { return wqkv + (size_t)(128 + (t >> 2) * 128) * FC + (t & 3) * 64; },
                 [&](int t, uint32_t wt) {
                   const int chb = t >> 2, kb = t & 3;
                   const uint64_t ad = make_desc(wt), bd = make_desc(ra + kb * (256 * 128));
#pragma unroll
                   for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + chb * 256, ad + 2 * k4, bd + 2 * k4, idesc, (kb | k4) != 0);
                 },
                 rw, wbar, wuse, tid);
    all_mma_done();   // x in R_A is dead from here on: v^T takes its place
    const int ch = half * 128 + quarter * 32 + lane;
    const float bv = bqkv[128 + ch];
#pragma unroll 1
    for (int c32 = 0; c32 < 8; ++c32) {
      float v[32];
      tmem_ld32(tmem + lane_base + half * 256 + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += bv;
      uint8_t* dst = gen + (ra - base) + (c32 >> 1) * (256 * 128);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(dst + sw128(ch, (c32 & 1) * 4 + q4)) = pack8(v + q4 * 8);
    }
    if (tid < 256) tab[tid] = bo[tid];   // the q|k bias table is dead; out-proj bias takes its place
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  // ================= phase 3: S_qb = q_qb k^T for both query blocks (N = 256 keys, K = 64)
  {
    const uint32_t idesc = make_idesc(256);
    if (tid == 0) {
#pragma unroll
      for (int qb = 0; qb < 2; ++qb) {
        const uint64_t ad = make_desc(rb + qb * (128 * 128)), bd = make_desc(rb + 256 * 128);
#pragma unroll
        for (int k4 = 0; k4 < FDK / 16; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, idesc, k4 != 0);
      }
    }
    all_mma_done();   // q,k tiles in R_B are dead: P / r reuse the region
  }

  const float cexp = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(d_k)
  for (int qb = 0; qb < 2; ++qb) {
    const uint32_t t_row = tmem + lane_base + qb * 256;
    const int lrow = quarter * 32 + lane;                   // row inside the query block
    const size_t grow = row0 + qb * 128 + lrow;
    // ---- softmax rows -> un-normalised P (bf16, K-major) in R_B.  All 8 warps: the two warps of a lane quarter
    // split the 256 key columns (4 chunks each) and exchange row max / row sum through the idle weight ring.
    float* xch = reinterpret_cast<float*>(gen + (rw - base));   // [2 halves][128 rows] max, then [2][128] sums
    float mloc = -INFINITY;
#pragma unroll 1
    for (int c32 = half * 4; c32 < half * 4 + 4; ++c32) {
      float v[32];
      tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) mloc = fmaxf(mloc, v[i]);
    }
    xch[half * 128 + lrow] = mloc;
    __syncthreads();
    const float mc = fmaxf(xch[lrow], xch[128 + lrow]) * cexp;
    float sum = 0.f;
#pragma unroll 1
    for (int c32 = half * 4; c32 < half * 4 + 4; ++c32) {
      float v[32];
      tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = exp2f(fmaf(v[i], cexp, -mc));
        sum += v[i];
      }
      uint8_t* pk = gen + (rb - base) + (c32 >> 1) * (128 * 128);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) *reinterpret_cast<uint4*>(pk + sw128(lrow, (c32 & 1) * 4 + q4)) = pack8(v + q4 * 8);
    }
    xch[256 + half * 128 + lrow] = sum;
    __syncthreads();
    const float inv_row = 1.0f / (xch[256 + lrow] + xch[384 + lrow]);
    if (half == 0) invs[lrow] = inv_row;
    if (p.attn_mode != 0) {  // attention map (need=True): mean of the four layers' maps, model5_b.py:468-469
      float* ag = p.attn + grow * FL;
#pragma unroll 1
      for (int c32 = half * 4; c32 < half * 4 + 4; ++c32) {
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          float4 a;
          a.x = exp2f(fmaf(v[q4 * 4 + 0], cexp, -mc)) * inv_row;
          a.y = exp2f(fmaf(v[q4 * 4 + 1], cexp, -mc)) * inv_row;
          a.z = exp2f(fmaf(v[q4 * 4 + 2], cexp, -mc)) * inv_row;
          a.w = exp2f(fmaf(v[q4 * 4 + 3], cexp, -mc)) * inv_row;
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
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- O = P v  (A = P [128 x 256 keys], B = v^T [256 ch x 256 keys]); overwrites S_qb
    if (tid == 0) {
      const uint32_t idesc = make_idesc(256);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t ad = make_desc(rb + kb * (128 * 128)), bd = make_desc(ra + kb * (256 * 128));
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, idesc, (kb | k4) != 0);
      }
    }
    all_mma_done();   // P is dead: r takes its place
    // ---- r = x - O / sum  -> R_B as the A operand of the out-projection; the two warp halves take alternate chunks
    {
      const float inv = invs[lrow];
      const __nv_bfloat16* xr = p.x + grow * p.ldx;
#pragma unroll 1
      for (int c32 = half; c32 < 8; c32 += 2) {
        uint4 xv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) xv[q4] = *reinterpret_cast<const uint4*>(xr + c32 * 32 + q4 * 8);
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
        uint8_t* dst = gen + (rb - base) + (c32 >> 1) * (128 * 128);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv[q4]);
          float rr[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 xf = __bfloat1622float2(xp[h]);
            rr[2 * h] = xf.x - v[q4 * 8 + 2 * h] * inv;
            rr[2 * h + 1] = xf.y - v[q4 * 8 + 2 * h + 1] * inv;
          }
          *reinterpret_cast<uint4*>(dst + sw128(lrow, (c32 & 1) * 4 + q4)) = pack8(rr);
        }
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();   // r complete, O_qb drained
    tc_fence_after();
    // ---- out[tok, ch] = r Wo^T : two 128-channel halves into the columns O_qb occupied
    {
      const uint32_t idesc = make_idesc(128);
      stream_phase(8, FC, [&](int t) { return wo + (size_t)((t >> 2) * 128) * FC + (t & 3) * 64; },
                   [&](int t, uint32_t wt) {
                     const int chh = t >> 2, kb = t & 3;
                     const uint64_t ad = make_desc(rb + kb * (128 * 128)), bd = make_desc(wt);
#pragma unroll
                     for (int k4 = 0; k4 < 4; ++k4)
                       umma_bf16(tmem + qb * 256 + chh * 128, ad + 2 * k4, bd + 2 * k4, idesc, (kb | k4) != 0);
                   },
                   rw, wbar, wuse, tid);
      all_mma_done();
    }
    // ---- out = x + relu(acc + bo): thread = token row, 128-bit stores into the layer's slice of att_cat
    {
      const __nv_bfloat16* xr = p.x + grow * p.ldx;
      __nv_bfloat16* yb = p.yb + grow * p.ldyb;
      float* yf = p.yf ? p.yf + grow * p.ldyf : nullptr;
#pragma unroll 1
      for (int c32 = half; c32 < 8; c32 += 2) {
        uint4 xv[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) xv[q4] = *reinterpret_cast<const uint4*>(xr + c32 * 32 + q4 * 8);
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv[q4]);
          float o8[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 xf = __bfloat1622float2(xp[h]);
            const int c = c32 * 32 + q4 * 8 + 2 * h;
            o8[2 * h] = xf.x + fmaxf(v[q4 * 8 + 2 * h] + tab[c], 0.f);
            o8[2 * h + 1] = xf.y + fmaxf(v[q4 * 8 + 2 * h + 1] + tab[c + 1], 0.f);
          }
          *reinterpret_cast<uint4*>(yb + c32 * 32 + q4 * 8) = pack8(o8);
          if (yf) {
            *reinterpret_cast<float4*>(yf + c32 * 32 + q4 * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4*>(yf + c32 * 32 + q4 * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // R_B (r) and the out columns are free for the next query block
    tc_fence_after();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

int launch_attention_layer_tc(const AttnLayerTc& p, int clouds, cudaStream_t st) {
  PZ_REQUIRE(p.x && p.wqkv[0] && p.wo[0] && p.bqkv[0] && p.bo[0] && p.yb, PZ_ERR_ARG, "attention_layer_tc: null pointer");
  PZ_REQUIRE(p.ldx % 8 == 0 && p.ldyb % 8 == 0 && ((uintptr_t)p.x & 15) == 0 && ((uintptr_t)p.yb & 15) == 0 &&
                 (!p.yf || (p.ldyf % 4 == 0 && ((uintptr_t)p.yf & 15) == 0)),
             PZ_ERR_ARG, "attention_layer_tc: rows must be 16-byte aligned");
  const size_t smem = 1024 + RA_BYTES + RB_BYTES + RW_BYTES + 1024 + 512 + 64;
  PZ_CUDA(cudaFuncSetAttribute(attention_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  attention_layer_tc_kernel<<<clouds, FT, smem, st>>>(p);
  PZ_LAUNCH_CHECK();
  return 0;
}

}  // namespace pz
