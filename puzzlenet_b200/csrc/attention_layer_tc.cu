// A stack of 1-4 offset-attention layers (layerAttention.forward, model5_b.py:92-101; the encoder applies four in a row,
// model5_b.py:463-466) as ONE tcgen05 kernel, one CTA per cloud.  Per layer:
//   q|k = x Wqk^T + b,  v = x Wv^T + b,  A = softmax(q k^T / sqrt(64)),  r = x - A v,  out = x + relu(r Wo^T + bo)
// for L = 256 tokens, C = 256 channels, d_k = 64 (bf16 operands, fp32 accumulation in TMEM); layer l + 1 only needs
// layer l of the same cloud, so a CTA walks the layers back to back.
// Nothing but each layer's x (in) and out (written into its 256-column slice of att_cat) touches HBM: q, k, v^T, P and
// r live in shared memory as K-major SWIZZLE_128B MMA operands.  x enters through TMA tensor loads ([128 tok x 64 ch]
// boxes, SWIZZLE_128B): as the projections' operand (layer 0 only -- afterwards the output tile of layer l, formed in
// R_A in operand layout, IS the x of layer l + 1), then again straight into the slots where r and the output tile are
// formed in place (the residuals); the output leaves through TMA tensor stores: a thread owns a token row (its TMEM
// lane), so direct global accesses would touch 32 cache lines per warp instruction.
//
// Warp-specialised: warps 0-7 are the epilogue group (TMEM -> registers -> shared-memory operands / global output),
// warp 8 lane 0 is the weight producer, warp 9 lane 0 issues every MMA.  The three roles only meet through
// mbarriers that complete exactly once per layer (layer l waits with parity l & 1):
//   full[t]  weight tile t has landed (bulk-copy complete_tx)     cons[t]  the MMAs reading tile t have completed
//   c_*      tcgen05.commit of an MMA phase (accumulator ready, operands dead)
//   g_*      256 epilogue arrivals: operand written to shared memory and the accumulator columns drained (g_out: the
//            layer's output tiles are stored and in place as the next x)
// so the q|k epilogue runs under the v^T MMAs, weight tiles stream under epilogues, and no phase starts with a cold
// weight ring.
//
// Weights arrive as pre-swizzled 16 KB tile images ([128 rows x 64 k] bf16, SWIZZLE_128B byte order, written once by
// the weight pack): one cp.async.bulk per tile, 20 tiles per layer:
//   t 0-3   Wqk  k-block t          (B operand of phase 1)      -> slots B0-B3 (R_B is idle until q|k are written)
//   t 4-11  Wv   (ch block, k-block) (A operand of phase 2)      -> ring W0/W1
//   t 12-19 Wo   (ch half, k-block)  (B operand of the out-projection, streamed ONCE for both query blocks)
//                                                                -> W0/W1, and A4-A7 once v^T is dead
// Shared memory, 14 slots of 16 KB:  R_A = A0-A7: x [256 tok x 256 ch] as 4 k-blocks -> v^T [256 ch x 256 tok]
//                                          -> r of query block 1 (A0-A3) + Wo tiles (A4-A7) -> x again -> out = next x;
//                                    R_B = B0-B3: Wqk tiles -> q,k -> P0 -> P1 -> r of query block 0;   R_W = W0,W1.
// TMEM (512 columns): q|k [0,256) | v^T ch 0-127 [256,512) -> v^T ch 128-255 [0,256) -> S0 [0,256) S1 [256,512)
//                     -> O in place -> out in place.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)

#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int FL = 256, FC = 256, FDK = 64;
constexpr uint32_t SLOT = 128 * 128;                       // one [128 x 64] bf16 tile
constexpr uint32_t RA_BYTES = 8 * SLOT, RB_BYTES = 4 * SLOT, RW_BYTES = 2 * SLOT;
constexpr int NEPI = 256, FT = 320;                        // epilogue threads, all threads
constexpr int NTILES = 20;

// barrier indices (8 bytes each)
constexpr int B_FULL = 0, B_CONS = NTILES - 4, B_C = 2 * NTILES - 4;    // cons[t] exists for t >= 4; c_qk c_v c_s c_pv0 c_pv1 c_out
constexpr int C_QK = B_C, C_V = B_C + 1, C_S = B_C + 2, C_PV0 = B_C + 3, C_PV1 = B_C + 4, C_OUT = B_C + 5;
constexpr int F_X = B_C + 6;                       // x k-block kb landed (TMA): for the projections, for r, for out
constexpr int F_XR = F_X + 4, F_XO = F_X + 8;
constexpr int G_QK = F_X + 12, G_V = G_QK + 1, G_P0 = G_QK + 2, G_R = G_QK + 4, G_OUT = G_QK + 5;   // G_P1 = G_P0 + 1
constexpr int NBARS = G_OUT + 1;
constexpr uint32_t MISC_BYTES = 1024 /*xch*/ + 512 /*q|k bias*/ + NBARS * 8 + 16;

__device__ __forceinline__ uint4 pack8(const float* v) {
  uint4 o;
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
  o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
  return o;
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// pz_profile_attention_timeline: CTA 0 stamps %globaltimer-free SM clocks at its phase boundaries
__device__ __forceinline__ void stamp(long long* prof, int slot) {
  if (prof != nullptr && blockIdx.x == 0) prof[slot] = clock64();
}
__device__ __forceinline__ void expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one [128 rows x 64 ch] bf16 box at (column c0, row c1) <-> a 16 KB SWIZZLE_128B slot
__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store(const CUtensorMap* tm, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(NEPI) : "memory"); }
}  // namespace

struct alignas(64) AttnMaps {
  CUtensorMap m[5];   // m[0]: input of layer 0; m[l + 1]: output of layer l = input of layer l + 1
};

__global__ void __launch_bounds__(FT, 1) attention_layer_tc_kernel(const AttnLayerTc p, const __grid_constant__ AttnMaps maps) {
  extern __shared__ __align__(1024) uint8_t fl_smem_raw[];
  const uint32_t base = (smem_u32(fl_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = fl_smem_raw + (base - smem_u32(fl_smem_raw));
  const uint32_t ra = base, rb = ra + RA_BYTES, rw = rb + RB_BYTES;
  const uint32_t xch_s = rw + RW_BYTES, tab_s = xch_s + 1024, bars = tab_s + 512, tmem_slot = bars + NBARS * 8;
  float* xch = reinterpret_cast<float*>(gen + (xch_s - base));   // [2 halves][128 rows]
  float* tab = reinterpret_cast<float*>(gen + (tab_s - base));   // q|k bias
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.x, set = cloud / p.clouds_per_set;
  const size_t row0 = (size_t)cloud * FL;
  stamp(tid == 0 ? p.prof : nullptr, 15);
  if (p.prof != nullptr && tid == 0) {   // every CTA: %globaltimer (ns) at entry / exit, slots 64 + 2 cta, 65 + 2 cta
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.prof[64 + 2 * blockIdx.x] = (long long)ns;
  }

  auto slot_of = [&](int t) -> uint32_t {
    if (t < 4) return rb + t * SLOT;
    if (t >= 14 && t < 18) return ra + (4 + t - 14) * SLOT;
    return rw + (t & 1) * SLOT;
  };
  // layer l of the stack: weight images, biases and tensor maps advance per layer; every barrier completes exactly once
  // per layer, so layer l waits with parity l & 1
  const uint8_t* wimg0 = reinterpret_cast<const uint8_t*>(p.wimg[set]);
  auto load_w = [&](int l, int t) {
    bulk_load(slot_of(t), wimg0 + ((size_t)l * NTILES + t) * SLOT, SLOT, bar(B_FULL + t));
  };
  const CUtensorMap* tmx0 = &maps.m[0];

  if (tid < 128) tab[tid] = p.bqkv[set][tid];
  if (tid == NEPI) {
    // the producer thread initialises the barriers and starts the first loads before anyone else needs them:
    // x -> R_A (4 k-blocks of [256 tokens x 64 ch] = 2 boxes each), Wqk -> B0-B3, the first two Wv tiles -> W0/W1
    for (int i = 0; i < NBARS; ++i) mbar_init(bar(i), i >= G_QK ? NEPI : 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    fence_proxy_async();
    for (int kb = 0; kb < 4; ++kb) {
      expect_tx(bar(F_X + kb), 2 * SLOT);
      tma_load(ra + kb * (2 * SLOT), tmx0, kb * 64, (int)row0, bar(F_X + kb));
      tma_load(ra + kb * (2 * SLOT) + SLOT, tmx0, kb * 64, (int)row0 + 128, bar(F_X + kb));
      load_w(0, kb);
    }
    load_w(0, 4);
    load_w(0, 5);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (warp == 8) {
    // ================= weight producer
    if (lane == 0) {
#pragma unroll 1
      for (int l = 0; l < p.nlayers; ++l) {
        const uint32_t par = l & 1;
        const CUtensorMap* tmx = &maps.m[l];
        for (int t = 6; t < 14; ++t) {
          mbar_wait(bar(B_CONS + t - 2), par);
          load_w(l, t);
        }
        // P and v^T are dead: x comes back into the slots where r is formed in place (block 0 in R_B, block 1 in
        // A0-A3), and the upper half of R_A takes four Wo tiles
        mbar_wait(bar(C_PV1), par);
        for (int kb = 0; kb < 4; ++kb) {
          expect_tx(bar(F_XR + kb), 2 * SLOT);
          tma_load(rb + kb * SLOT, tmx, kb * 64, (int)row0, bar(F_XR + kb));
          tma_load(ra + kb * SLOT, tmx, kb * 64, (int)row0 + 128, bar(F_XR + kb));
        }
        for (int t = 14; t < 18; ++t) load_w(l, t);
        for (int t = 18; t < 20; ++t) {
          mbar_wait(bar(B_CONS + t - 6), par);
          load_w(l, t);
        }
        // r and every weight tile are dead: x once more, for the output residual, into R_A in operand layout (the
        // output tile formed over it is the next layer's x), and the next layer's first weight tiles
        mbar_wait(bar(C_OUT), par);
        for (int kb = 0; kb < 4; ++kb) {
          expect_tx(bar(F_XO + kb), 2 * SLOT);
          tma_load(ra + kb * (2 * SLOT), tmx, kb * 64, (int)row0, bar(F_XO + kb));
          tma_load(ra + kb * (2 * SLOT) + SLOT, tmx, kb * 64, (int)row0 + 128, bar(F_XO + kb));
        }
        if (l + 1 < p.nlayers)
          for (int t = 0; t < 6; ++t) load_w(l + 1, t);
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ================= MMA issuer
    if (lane == 0) {
      const uint32_t id128 = make_idesc(128), id256 = make_idesc(256);
#pragma unroll 1
      for (int l = 0; l < p.nlayers; ++l) {
      const uint32_t par = l & 1;
      stamp(p.prof, 32);
      // phase 1: q|k[tok, 0:128] = x Wqk^T   (rows on lanes; two 128-token blocks; N = 128) -> cols [0,256)
      if (l > 0) {   // x = the previous layer's output tiles, formed in R_A; its accumulators are drained
        mbar_wait(bar(G_OUT), par ^ 1);
        tc_fence_after();
      }
      for (int t = 0; t < 4; ++t) {
        if (l == 0) mbar_wait(bar(F_X + t), 0);
        mbar_wait(bar(B_FULL + t), par);
        tc_fence_after();
        const uint64_t bd = make_desc(slot_of(t));
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint64_t ad = make_desc(ra + t * (2 * SLOT) + blk * SLOT);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + blk * 128, ad + 2 * k4, bd + 2 * k4, id128, (t | k4) != 0);
        }
      }
      umma_commit(bar(C_QK));
      stamp(p.prof, 33);
      // phase 2: v^T[ch, tok] = Wv x^T  (channels on lanes; N = 256 tokens): ch block 0 -> cols [256,512),
      // ch block 1 -> cols [0,256) once the q|k accumulators are drained
      for (int j = 0; j < 8; ++j) {
        const int t = 4 + j, chb = j >> 2, kb = j & 3;
        if (j == 4) {
          stamp(p.prof, 34);
          mbar_wait(bar(G_QK), par);
          tc_fence_after();
          stamp(p.prof, 35);
        }
        mbar_wait(bar(B_FULL + t), par);
        tc_fence_after();
        const uint64_t ad = make_desc(slot_of(t)), bd = make_desc(ra + kb * (2 * SLOT));
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + (1 - chb) * 256, ad + 2 * k4, bd + 2 * k4, id256, (kb | k4) != 0);
        umma_commit(bar(B_CONS + t));
      }
      umma_commit(bar(C_V));
      stamp(p.prof, 36);
      // phase 3: S_qb = q_qb k^T for both query blocks (N = 256 keys, K = 64)
      mbar_wait(bar(G_V), par);
      tc_fence_after();
      stamp(p.prof, 37);
#pragma unroll
      for (int qb = 0; qb < 2; ++qb) {
        const uint64_t ad = make_desc(rb + qb * SLOT), bd = make_desc(rb + 2 * SLOT);
#pragma unroll
        for (int k4 = 0; k4 < FDK / 16; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, id256, k4 != 0);
      }
      umma_commit(bar(C_S));
      // O_qb = P_qb v  (A = P [128 x 256 keys], B = v^T [256 ch x 256 keys]); overwrites S_qb
      for (int qb = 0; qb < 2; ++qb) {
        mbar_wait(bar(G_P0 + qb), par);
        tc_fence_after();
        stamp(p.prof, 38 + qb);
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t ad = make_desc(rb + kb * SLOT), bd = make_desc(ra + kb * (2 * SLOT));
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) umma_bf16(tmem + qb * 256, ad + 2 * k4, bd + 2 * k4, id256, (kb | k4) != 0);
        }
        umma_commit(bar(C_PV0 + qb));
      }
      // out[tok, ch] = r Wo^T for both query blocks per weight tile, into the columns O occupied
      mbar_wait(bar(G_R), par);
      tc_fence_after();
      stamp(p.prof, 40);
      for (int j = 0; j < 8; ++j) {
        const int t = 12 + j, chh = j >> 2, kb = j & 3;
        mbar_wait(bar(B_FULL + t), par);
        tc_fence_after();
        const uint64_t bd = make_desc(slot_of(t));
#pragma unroll
        for (int blk = 0; blk < 2; ++blk) {
          const uint64_t ad = make_desc((blk == 0 ? rb : ra) + kb * SLOT);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem + blk * 256 + chh * 128, ad + 2 * k4, bd + 2 * k4, id128, (kb | k4) != 0);
        }
        umma_commit(bar(B_CONS + t));
      }
      umma_commit(bar(C_OUT));
      stamp(p.prof, 41);
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue group (256 threads): thread = TMEM lane quarter*32 + lane, two warps per quarter.
    // Accumulators are read 64 columns (= one 128-byte operand row) per tcgen05.ld.
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const int lrow = quarter * 32 + lane;
    long long* eprof = tid == 0 ? p.prof : nullptr;
#pragma unroll 1
    for (int l = 0; l < p.nlayers; ++l) {
    const uint32_t par = l & 1;
    const float* __restrict__ bqkv = p.bqkv[set] + (size_t)l * 384;
    const float* __restrict__ bo = p.bo[set][l];
    const int attn_mode = p.attn_mode[l];
    auto publish = [&](int b) {   // operand written / accumulator drained -> MMA issuer
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar(b));
    };
    auto acquire = [&](int b) {   // an MMA phase has completed
      mbar_wait(bar(b), par);
      tc_fence_after();
    };
    stamp(eprof, 0);
    if (l > 0) {   // this layer's q|k bias (the table's readers, the previous q|k epilogue, are long done)
      if (tid < 128) tab[tid] = bqkv[tid];
      epi_bar();
    }

    // ---- q|k epilogue (under the v^T MMAs of channel block 0): warp half h owns token block h;
    // q -> R_B[0:32K) as [256 x 64], k -> R_B[32K:64K)
    acquire(C_QK);
    stamp(eprof, 2);
    {
      const int row = half * 128 + lrow;
#pragma unroll 1
      for (int c64 = 0; c64 < 2; ++c64) {
        float v[64];
        tmem_ld64(tmem + lane_base + half * 128 + c64 * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] += tab[c64 * 64 + i];
        uint8_t* dst = gen + (rb - base) + c64 * (2 * SLOT);
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) *reinterpret_cast<uint4*>(dst + sw128(row, q8)) = pack8(v + q8 * 8);
      }
    }
    publish(G_QK);
    stamp(eprof, 3);

    // ---- v^T epilogue: thread = channel; x in R_A is dead, v^T takes its place
    acquire(C_V);
    stamp(eprof, 4);
    {
      const int ch = half * 128 + lrow;
      const float bv = bqkv[128 + ch];
      const uint32_t t_v = tmem + lane_base + (1 - half) * 256;
#pragma unroll 1
      for (int c64 = 0; c64 < 4; ++c64) {
        float v[64];
        tmem_ld64(t_v + c64 * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] += bv;
        uint8_t* dst = gen + (ra - base) + c64 * (2 * SLOT);
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) *reinterpret_cast<uint4*>(dst + sw128(ch, q8)) = pack8(v + q8 * 8);
      }
    }
    publish(G_V);
    stamp(eprof, 5);

    // ---- softmax rows -> un-normalised P (bf16, K-major) in R_B.  The two warps of a lane quarter split the 256
    // key columns (128 each) and exchange row max / row sum through xch: a thread publishes in its own slot
    // first (max), then in its partner's slot (sum), so the 1 KB is reused without a further barrier.
    const float cexp = 1.4426950408889634f / 8.0f;  // log2(e) / sqrt(d_k)
    float inv0 = 0.f, inv1 = 0.f;
    acquire(C_S);   // q,k tiles in R_B are dead
    stamp(eprof, 6);
#pragma unroll 1
    for (int qb = 0; qb < 2; ++qb) {
      const uint32_t t_row = tmem + lane_base + qb * 256;
      const size_t grow = row0 + qb * 128 + lrow;
      float mloc = -INFINITY;
#pragma unroll 1
      for (int c64 = half * 2; c64 < half * 2 + 2; ++c64) {
        float v[64];
        tmem_ld64(t_row + c64 * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) mloc = fmaxf(mloc, v[i]);
      }
      xch[half * 128 + lrow] = mloc;
      epi_bar();
      const float mc = fmaxf(mloc, xch[(1 - half) * 128 + lrow]) * cexp;
      if (qb == 1) {
        stamp(eprof, 8);
        acquire(C_PV0);   // P0 has been consumed: R_B takes P1
        stamp(eprof, 9);
      }
      float sum = 0.f;
#pragma unroll 1
      for (int c64 = half * 2; c64 < half * 2 + 2; ++c64) {
        float v[64];
        tmem_ld64(t_row + c64 * 64, v);
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          v[i] = exp2f(fmaf(v[i], cexp, -mc));
          sum += v[i];
        }
        uint8_t* pk = gen + (rb - base) + c64 * SLOT;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) *reinterpret_cast<uint4*>(pk + sw128(lrow, q8)) = pack8(v + q8 * 8);
      }
      xch[(1 - half) * 128 + lrow] = sum;
      epi_bar();
      const float inv_row = 1.0f / (sum + xch[half * 128 + lrow]);
      if (qb == 0) inv0 = inv_row; else inv1 = inv_row;
      if (attn_mode != 0) {  // attention map (need=True): mean of the four layers' maps, model5_b.py:468-469
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
            if (attn_mode != 1) {
              const float4 o = *dst;
              a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
              if (attn_mode == 3) { a.x *= 0.25f; a.y *= 0.25f; a.z *= 0.25f; a.w *= 0.25f; }
            }
            *dst = a;
          }
        }
      }
      publish(G_P0 + qb);
      stamp(eprof, qb == 0 ? 7 : 10);
    }

    // ---- from here on warp half h owns query block h
    const int qb = half;
    const size_t grow = row0 + qb * 128 + lrow;
    const uint32_t t_row = tmem + lane_base + qb * 256;
    const uint32_t rslots = qb == 0 ? rb : ra;            // the block's 4 slots: x -> r in place, x -> out in place
    uint8_t* rdst = gen + (rslots - base);
    // r = x - O / sum  -> the A operand of the out-projection, formed over the x tiles the producer brought back
    acquire(C_PV1);
    stamp(eprof, 11);   // both P v products are complete (commits complete in order): P and v^T are dead
    {
      const float inv = qb == 0 ? inv0 : inv1;
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        float v[64];
        tmem_ld64(t_row + kb * 64, v);
        mbar_wait(bar(F_XR + kb), par);
        if (kb == 0) stamp(eprof, 1);
        uint8_t* dst = rdst + kb * SLOT;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          uint4* slot16 = reinterpret_cast<uint4*>(dst + sw128(lrow, q8));
          const uint4 xv = *slot16;
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv);
          float rr[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 xf = __bfloat1622float2(xp[h]);
            rr[2 * h] = xf.x - v[q8 * 8 + 2 * h] * inv;
            rr[2 * h + 1] = xf.y - v[q8 * 8 + 2 * h + 1] * inv;
          }
          *slot16 = pack8(rr);
        }
      }
    }
    publish(G_R);
    stamp(eprof, 12);
    epi_bar();            // every thread is past the softmax exchange: xch takes the out-projection bias
    xch[tid] = bo[tid];
    epi_bar();

    // ---- out = x + relu(acc + bo), formed in place over x (R_A, operand layout: it is the next layer's x) and handed
    // to TMA one [128 x 64] tile at a time
    acquire(C_OUT);
    stamp(eprof, 13);
    {
      float* yf = p.yf ? p.yf + grow * p.ldyf + (size_t)l * p.yf_layer_stride : nullptr;
      const CUtensorMap* tmy = &maps.m[l + 1];
#pragma unroll 1
      for (int kb = 0; kb < 4; ++kb) {
        float v[64];
        tmem_ld64(t_row + kb * 64, v);
        mbar_wait(bar(F_XO + kb), par);
        if (kb == 0) stamp(eprof, 16);
        uint8_t* dst = gen + (ra - base) + kb * (2 * SLOT) + qb * SLOT;
#pragma unroll
        for (int q8 = 0; q8 < 8; ++q8) {
          uint4* slot16 = reinterpret_cast<uint4*>(dst + sw128(lrow, q8));
          const uint4 xv = *slot16;
          const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xv);
          const float* bb = xch + kb * 64 + q8 * 8;
          float o8[8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 xf = __bfloat1622float2(xp[h]);
            o8[2 * h] = xf.x + fmaxf(v[q8 * 8 + 2 * h] + bb[2 * h], 0.f);
            o8[2 * h + 1] = xf.y + fmaxf(v[q8 * 8 + 2 * h + 1] + bb[2 * h + 1], 0.f);
          }
          *slot16 = pack8(o8);
          if (yf) {
            *reinterpret_cast<float4*>(yf + kb * 64 + q8 * 8) = make_float4(o8[0], o8[1], o8[2], o8[3]);
            *reinterpret_cast<float4*>(yf + kb * 64 + q8 * 8 + 4) = make_float4(o8[4], o8[5], o8[6], o8[7]);
          }
        }
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(2 + half) : "memory");
        if (lrow == 0) {
          tma_store(tmy, kb * 64, (int)row0 + qb * 128, ra + kb * (2 * SLOT) + qb * SLOT);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
    }
    stamp(eprof, 17);
    // the stores are complete (the next layer reloads this output from global for its residuals) before the tiles and
    // the accumulator columns are handed to the next layer
    if (lrow == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    publish(G_OUT);
    stamp(eprof, 14);
    }
  }
  __syncthreads();   // every TMEM read has retired, the TMA stores have read their tiles
  if (p.prof != nullptr && tid == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    p.prof[65 + 2 * blockIdx.x] = (long long)ns;
  }
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// fp32 [rows, 256] weight block -> pre-swizzled [128 x 64] bf16 tile images; image tile = (row / 128) * 4 + col / 64
__global__ void __launch_bounds__(256) attn_weight_image_kernel(const float* __restrict__ src, int rows, int row_off,
                                                                __nv_bfloat16* __restrict__ img) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= rows * FC) return;
  const int r = e / FC + row_off, c = e % FC;
  const int tile = (r >> 7) * 4 + (c >> 6), rr = r & 127, cc = c & 63;
  const uint32_t off = sw128(rr, cc >> 3) + (cc & 7) * 2;
  img[(size_t)tile * (SLOT / 2) + off / 2] = __float2bfloat16_rn(src[(size_t)(e / FC) * FC + c]);
}

int launch_attn_weight_image(const float* src, int rows, int row_off, __nv_bfloat16* img, cudaStream_t st) {
  attn_weight_image_kernel<<<(rows * FC + 255) / 256, 256, 0, st>>>(src, rows, row_off, img);
  PZ_LAUNCH_CHECK();
  return 0;
}

static long long* g_attn_timeline = nullptr;
static long long g_attn_timeline_slots = 0;
// the registered buffer when it holds at least min_slots int64, else null (a short buffer switches that stamp set off)
long long* kernel_timeline_buffer(long long min_slots) {
  return g_attn_timeline_slots >= min_slots ? g_attn_timeline : nullptr;
}

// [rows, 256] bf16 view with a row stride of ld elements, traversed in [128 rows x 64 ch] SWIZZLE_128B boxes
static int make_tile_map(const __nv_bfloat16* ptr, int ld, size_t rows, CUtensorMap* out) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      fn = nullptr;
    return reinterpret_cast<EncodeFn>(fn);
  }();
  PZ_REQUIRE(encode != nullptr, PZ_ERR_UNSUPPORTED, "attention_layer_tc: the driver does not export cuTensorMapEncodeTiled");
  const cuuint64_t dims[2] = {(cuuint64_t)FC, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<__nv_bfloat16*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PZ_REQUIRE(r == CUDA_SUCCESS, PZ_ERR_ARG, "attention_layer_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int launch_attention_layer_tc(const AttnLayerTc& p_in, int clouds, cudaStream_t st) {
  AttnLayerTc p = p_in;
  p.prof = kernel_timeline_buffer(64 + 2 * (long long)clouds);
  PZ_REQUIRE(p.x && p.wimg[0] && p.bqkv[0] && p.bo[0][0] && p.yb, PZ_ERR_ARG, "attention_layer_tc: null pointer");
  PZ_REQUIRE(p.nlayers >= 1 && p.nlayers <= 4, PZ_ERR_ARG, "attention_layer_tc: %d layers (1..4)", p.nlayers);
  for (int l = 0; l < p.nlayers; ++l)
    PZ_REQUIRE(p.bo[0][l] && (p.clouds_per_set >= clouds || p.bo[1][l]), PZ_ERR_ARG, "attention_layer_tc: missing bias of layer %d", l);
  PZ_REQUIRE(p.ldx % 8 == 0 && p.ldyb % 8 == 0 && ((uintptr_t)p.x & 15) == 0 && ((uintptr_t)p.yb & 15) == 0 &&
                 (!p.yf || (p.ldyf % 4 == 0 && ((uintptr_t)p.yf & 15) == 0)) && ((uintptr_t)p.wimg[0] & 15) == 0,
             PZ_ERR_ARG, "attention_layer_tc: rows must be 16-byte aligned");
  const size_t smem = 1024 + RA_BYTES + RB_BYTES + RW_BYTES + MISC_BYTES;
  static_assert(1024 + RA_BYTES + RB_BYTES + RW_BYTES + MISC_BYTES <= 232448, "attention layer: shared memory budget");
  PZ_CUDA(cudaFuncSetAttribute(attention_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AttnMaps maps;
  PZ_TRY(make_tile_map(p.x, p.ldx, (size_t)clouds * FL, &maps.m[0]));
  for (int l = 0; l < 4; ++l)   // unused layers repeat the last output: every map handed to the kernel is valid
    PZ_TRY(make_tile_map(p.yb + (size_t)(l < p.nlayers ? l : p.nlayers - 1) * p.yb_layer_stride, p.ldyb, (size_t)clouds * FL,
                         &maps.m[l + 1]));
  attention_layer_tc_kernel<<<clouds, FT, smem, st>>>(p, maps);
  PZ_LAUNCH_CHECK();
  return 0;
}

}  // namespace pz

// diagnostics: a device buffer of n_slots int64 that CTA 0 of every following fused attention-layer launch fills with SM
// clock stamps (slots 0-14, 16, 17 epilogue thread 0, 15 kernel entry, 32-41 the MMA issuer; needs n_slots >= 64 + 2 * clouds)
// and CTA 0 of the stage-1 gather GEMM with its own (slots 1024..1455; needs n_slots >= 1456); null switches it off
extern "C" int pz_profile_attention_timeline(long long* device_buf_or_null, long long n_slots) {
  PZ_REQUIRE(device_buf_or_null == nullptr || n_slots >= 64, PZ_ERR_ARG,
             "pz_profile_attention_timeline: the buffer must hold at least 64 int64 (got %lld)", n_slots);
  pz::g_attn_timeline = device_buf_or_null;
  pz::g_attn_timeline_slots = device_buf_or_null ? n_slots : 0;
  return 0;
}
