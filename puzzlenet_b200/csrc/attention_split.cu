// Offset-attention core of the SPLIT path (PZ_PREC_SPLIT) on tcgen05: per cloud (L = 256 tokens, d_k = 64, C = 256)
//   r = x - softmax(q k^T / sqrt(d_k)) v                      (model5_b.py:67-75 and :98)
// with every operand as fp16 hi / lo planes and every product as three MMAs (hi*hi + hi*lo + lo*hi, fp32 accumulation):
// the logits, the probabilities and P v all keep ~22 mantissa bits, so the attention map and r match the fp32 reference
// to ~1e-6 -- the bf16 kernel's map is off by 0.14 of its maximum with peaky logits.
//
// One CTA per (cloud, block of 128 queries): 2 CTAs per cloud.  S = q k^T lands with the query rows on the TMEM lanes,
// so each of the 128 softmax threads owns one full row (max / exp2 / sum in-thread).  The un-normalised probabilities go
// back to shared memory as the K-major A operand of O = P v (both planes), whose B operand v^T (written transposed by
// the v projection's epilogue) streams by TMA through a 2-stage ring of [128 channels x 64 keys] x 2 planes while the softmax
// runs; O is produced as two 128-channel halves so the epilogue of the first overlaps the MMAs of the second.
// Warps 0-3 softmax + epilogue, warp 4 the MMA-issuing thread, one thread of warp 5 issues every TMA load of v^T (q and k
// arrive by TMA as well, issued by thread 0 before the CTA-wide barrier).
// Shared memory: region A 128 KB (q, k planes -> P planes -> r planes) + ring 64 KB + 16 KB staging tiles of the epilogues.
// FUSED OUT-PROJECTION (AttnSplit::wo_hi set; the split encoder's default): r is not written to HBM -- the epilogue writes
// its hi / lo planes over P as the K-major A operand of  out = x + relu(Wo r + bo), Wo streams through the same ring (k-block
// major, so the first two k-blocks run under the epilogue of r's second half), the accumulators land over S, and a second
// epilogue adds bias / ReLU / residual and stores the layer's output planes (and fp32 rows): one launch and 67 MB of HBM
// traffic per layer less than a separate out-projection GEMM.
// CHAINED PROJECTIONS (AttnSplit::wqkv_hi set; layers 0-2 of the split encoder): the second epilogue also writes the output
// planes over r as the A operand of the NEXT layer's [q | k | v] = out Wqkv^T + b, Wqkv streams through the ring (12 more
// stages; the first six under the epilogue of the output's second half, the rest block-major so that every block's epilogue
// overlaps the remaining MMAs), q|k accumulates over the drained output half, v over the O halves, and a third epilogue
// stores the next layer's q|k planes and v^T planes into the OTHER buffer set (the CTAs of this launch still read the
// current one): two more launches and one more read of the layer's output per layer less.  TMEM: S [0,256), O halves [256,384), [384,512).
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_fp16.h>
#include <stdlib.h>

#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int AS_THREADS = 256;
constexpr int AS_L = 256, AS_C = 256;
constexpr uint32_t T16 = 128 * 128;                 // one [128 x 64] fp16 tile
constexpr uint32_t REGA = 8 * T16;                  // 128 KB
constexpr int AS_NST = 2;                            // two stages: the freed 32 KB hold the epilogue's staging tiles
constexpr uint32_t AS_STAGE = 2 * T16;              // hi + lo of [128 ch x 64 keys]
#define AS_VSLOT(it) ((it) == 2 ? 6u : 2u * ((it) - 3u))   // first region-A tile of v^T stage it (2..5): 6, 0, 2, 4

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
// A operand from TENSOR memory (lane = row, a 32-bit column = two consecutive K elements)
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// 32 consecutive 32-bit columns of this thread's TMEM lane <- 32 registers
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void as_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// one [128 rows x 64 columns] fp16 box at (column c0, row c1) -> a 16 KB SWIZZLE_128B tile
__device__ __forceinline__ void as_tma_load(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// timeline stamps of CTA 0 (diagnostics, pz_profile_attention_timeline): globaltimer ns at slots 3072..: 0 entry, 1 q|k
// landed, 2 S in TMEM, 3 P written, 4 / 6 O half 0 / 1 complete, 5 / 7 r half 0 / 1 formed, 8 / 10 out half accumulated, 9 / 11
// stored, 12-14 next q|k / v block accumulated, 15 all stored
__device__ __forceinline__ void as_stamp(long long* prof, int slot) {
  if (prof != nullptr && blockIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    prof[3072 + slot] = (long long)t;
  }
}
struct alignas(64) AsMaps {
  CUtensorMap qk[2];   // q|k hi / lo planes: [rows, 128]
  CUtensorMap vT[2];   // v^T hi / lo planes: [clouds * 256 channels, 256 keys]
  CUtensorMap wo[2][2];   // fused out-projection: [weight set][hi / lo] Wo planes [256 out channels, 256 k]
  CUtensorMap wn[2][2];   // chained projections: [weight set][hi / lo] next layer's Wqkv planes [384, 256]
};
}  // namespace

__global__ void __launch_bounds__(AS_THREADS, 1) attention_split_kernel(const AttnSplit p, const __grid_constant__ AsMaps maps, long long* prof) {
  extern __shared__ __align__(1024) uint8_t as_smem_raw[];
  const uint32_t base = (smem_u32(as_smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = as_smem_raw + (base - smem_u32(as_smem_raw));
  const uint32_t ring = base + REGA;
  const uint32_t bars = ring + AS_NST * AS_STAGE;
  const uint32_t bar_s = bars, bar_p = bars + 8, bar_o = bars + 16 /* 2 */, full_bar = bars + 32, empty_bar = full_bar + 8 * AS_NST;
  const uint32_t bar_qk = empty_bar + 8 * AS_NST;
  const uint32_t bar_r = bar_qk + 8 /* 2 */, bar_y = bar_r + 16 /* 2 */;   // fused out-projection: r half h written / out half h accumulated
  const uint32_t bar_pk = bar_y + 16 /* 4 */;                  // P key block kb (64 keys, both planes) written
  const uint32_t bar_x2 = bar_pk + 32 /* 2 */, bar_z = bar_x2 + 16 /* 3 */;   // chained projections: out half h in region A / block accumulated
  const uint32_t bar_va = bar_z + 24 /* 4 */;                  // v^T stages 2..5 landed in region A (slots of two tiles: 6, 0, 2, 4)
  const uint32_t tmem_slot = bar_va + 32;
  const uint32_t bo_s = (tmem_slot + 16 + 15) & ~15u;         // [256] out-projection bias of this CTA's weight set, then [384] next q|k|v bias
  const uint32_t stg_all = (bo_s + 1024 + 1536 + 127) & ~127u;   // 4 x 4 KB staging tiles of the epilogues (reused by all phases)
  const bool fuse = p.wo_hi[0] != nullptr;
  const bool chain = fuse && p.wqkv_hi[0] != nullptr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cloud = blockIdx.x >> 1, qb = blockIdx.x & 1;
  const size_t row0 = (size_t)cloud * AS_L;
  const int wset = (fuse && p.clouds_per_set > 0 && cloud >= p.clouds_per_set) ? 1 : 0;

  // ---- q (this block's 128 rows) and k (all 256 rows), both planes -> region A, by TMA (six 16 KB tiles on one mbarrier)
  const uint32_t q_hi_s = base, q_lo_s = base + T16, k_hi_s = base + 2 * T16, k_lo_s = base + 4 * T16;
  if (tid == 0) {
    as_stamp(prof, 0);
    mbar_init(bar_s, 1);
    mbar_init(bar_p, 128);
    mbar_init(bar_o, 1);
    mbar_init(bar_o + 8, 1);
    mbar_init(bar_qk, 1);
    mbar_init(bar_r, 128);
    mbar_init(bar_r + 8, 128);
    for (int kb = 0; kb < 4; ++kb) mbar_init(bar_pk + 8 * kb, 128);
    mbar_init(bar_x2, 128);
    mbar_init(bar_x2 + 8, 128);
    for (int b3 = 0; b3 < 3; ++b3) mbar_init(bar_z + 8 * b3, 1);
    for (int a = 0; a < 4; ++a) mbar_init(bar_va + 8 * a, 1);
    mbar_init(bar_y, 1);
    mbar_init(bar_y + 8, 1);
    for (int s = 0; s < AS_NST; ++s) {
      mbar_init(full_bar + 8 * s, 1);
      mbar_init(empty_bar + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the inits are visible to the TMA unit
    if (p.dep_flags == nullptr) {
      pdl_wait();                   // the whole previous grid (pz_common.cuh: programmatic dependent launch)
    } else {                        // both CTAs of this cloud in the previous layer's launch have stored their last row
      const int* f = p.dep_flags + cloud;
      int v;
      do {
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      } while (v < 2);
      asm volatile("fence.proxy.async;" ::: "memory");   // their generic-proxy stores, before this CTA's TMA reads
    }
    const int r = (int)row0;
    as_expect_tx(bar_qk, 6 * T16);
    as_tma_load(q_hi_s, &maps.qk[0], 0, r + qb * 128, bar_qk);
    as_tma_load(q_lo_s, &maps.qk[1], 0, r + qb * 128, bar_qk);
    as_tma_load(k_hi_s, &maps.qk[0], 64, r, bar_qk);
    as_tma_load(k_hi_s + T16, &maps.qk[0], 64, r + 128, bar_qk);
    as_tma_load(k_lo_s, &maps.qk[1], 64, r, bar_qk);
    as_tma_load(k_lo_s + T16, &maps.qk[1], 64, r + 128, bar_qk);
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.dep_flags == nullptr) pdl_wait();
  if (fuse) reinterpret_cast<float*>(gen + (bo_s - base))[tid] = p.bo[wset][tid];   // AS_THREADS == 256 == channels
  if (chain)
    for (int i = tid; i < 384; i += AS_THREADS) reinterpret_cast<float*>(gen + (bo_s - base))[256 + i] = p.bqkv[wset][i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // every CTA of this launch has passed its own wait (grid or cloud flag) before the next launch may become resident: its
  // CTAs then never run ahead of anything older than this launch (the flags they poll were zeroed long before)
  pdl_trigger();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  // r = x - O / sum for one 128-channel half `eh` of the block's 128 rows, by the four warps (quarters wq) of one group
  auto r_epilogue = [&](const int eh, const int wq, float inv) {
    const uint32_t tq_row = tmem + ((uint32_t)(wq * 32) << 16);
    // Global traffic is coalesced through a 4 KB per-warp staging tile (two planes of [32 rows][64 B], 16-byte pieces XOR-
    // swizzled so that both the row-per-thread and the row-segment-per-4-lanes accesses are conflict-free): a thread owns a
    // row, and direct 16-byte accesses touch 32 cache lines per warp instruction -- measured, the r epilogue took 13 of the
    // CTA's 21 us on L1 wavefronts alone.  The x block of chunk c + 1 is fetched (coalesced role) while chunk c is finished.
    const uint32_t stg = stg_all + (uint32_t)wq * 4096;
    const int cr = lane >> 2, cp = lane & 3;                    // coalesced role: row 8 j + cr, piece cp
    const uint32_t own16 = stg + (uint32_t)lane * 64, sw_own = (uint32_t)((lane >> 1) & 3);
    const size_t rowbase = row0 + qb * 128 + wq * 32;
    const __half* xhb = static_cast<const __half*>(p.x_hi);
    const __half* xlb = static_cast<const __half*>(p.x_lo);
    __half* rhb = static_cast<__half*>(p.r_hi);
    __half* rlb = static_cast<__half*>(p.r_lo);
    uint4 pre[8];
    auto x_fetch = [&](int cb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t off = (rowbase + 8 * j + cr) * p.ldx + cb + cp * 8;
        pre[j] = *reinterpret_cast<const uint4*>(xhb + off);
        pre[4 + j] = *reinterpret_cast<const uint4*>(xlb + off);
      }
    };
    x_fetch(eh * 128);
    {
      const int h = eh;
      mbar_wait(bar_o + 8 * h, 0);
      if (fuse) mbar_wait(bar_o + 8, 0);   // r goes over P in region A: every P v MMA must have completed
      tc_fence_after();
      if (lane == 0 && wq == 0) as_stamp(prof, 4 + 2 * h);
#pragma unroll 1
      for (int c32 = 0; c32 < 4; ++c32) {
        const int cb = h * 128 + c32 * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // the prefetched x block -> staging
          const int r = 8 * j + cr;
          const uint32_t a = stg + (uint32_t)r * 64 + (uint32_t)((cp ^ ((r >> 1) & 3)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pre[j].x), "r"(pre[j].y), "r"(pre[j].z), "r"(pre[j].w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(pre[4 + j].x), "r"(pre[4 + j].y), "r"(pre[4 + j].z), "r"(pre[4 + j].w) : "memory");
        }
        __syncwarp();
        if (c32 + 1 < 4) x_fetch(cb + 32);
        float v[32];
        tmem_ld32(tq_row + 256 + cb, v);
        uint4 oh[4], ol[4];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 xh4, xl4;
          const uint32_t a = own16 + (((uint32_t)q4 ^ sw_own) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xh4.x), "=r"(xh4.y), "=r"(xh4.z), "=r"(xh4.w) : "r"(a));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xl4.x), "=r"(xl4.y), "=r"(xl4.z), "=r"(xl4.w) : "r"(a + 2048));
          const uint32_t* hp = reinterpret_cast<const uint32_t*>(&xh4);
          const uint32_t* lp = reinterpret_cast<const uint32_t*>(&xl4);
          uint32_t* ohp = reinterpret_cast<uint32_t*>(&oh[q4]);
          uint32_t* olp = reinterpret_cast<uint32_t*>(&ol[q4]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&hp[e]));
            const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&lp[e]));
            split2h((a2.x + b2.x) - v[q4 * 8 + 2 * e] * inv, (a2.y + b2.y) - v[q4 * 8 + 2 * e + 1] * inv, ohp[e], olp[e]);
          }
        }
        __syncwarp();                 // every lane has read its x row: the tile takes the r block
        if (fuse) {                   // r stays on chip: K-major A operand of the out-projection, k-block cb / 64 of region A
          const uint32_t rk = base + (uint32_t)(cb >> 6) * T16;
          const int prow_ = wq * 32 + lane;
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const uint32_t off = sw128(prow_, ((cb & 63) >> 3) + q4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rk + off), "r"(oh[q4].x), "r"(oh[q4].y), "r"(oh[q4].z), "r"(oh[q4].w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rk + 4 * T16 + off), "r"(ol[q4].x), "r"(ol[q4].y), "r"(ol[q4].z), "r"(ol[q4].w) : "memory");
          }
          continue;
        }
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const uint32_t a = own16 + (((uint32_t)q4 ^ sw_own) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(oh[q4].x), "r"(oh[q4].y), "r"(oh[q4].z), "r"(oh[q4].w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(ol[q4].x), "r"(ol[q4].y), "r"(ol[q4].z), "r"(ol[q4].w) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 8 * j + cr;
          const uint32_t a = stg + (uint32_t)r * 64 + (uint32_t)((cp ^ ((r >> 1) & 3)) << 4);
          uint4 hh, ll;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hh.x), "=r"(hh.y), "=r"(hh.z), "=r"(hh.w) : "r"(a));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ll.x), "=r"(ll.y), "=r"(ll.z), "=r"(ll.w) : "r"(a + 2048));
          const size_t off = (rowbase + r) * AS_C + cb + cp * 8;
          *reinterpret_cast<uint4*>(rhb + off) = hh;
          *reinterpret_cast<uint4*>(rlb + off) = ll;
        }
        __syncwarp();
      }
      if (lane == 0 && wq == 0) as_stamp(prof, 5 + 2 * h);
    }
  };
  // fused out-projection epilogue: out = x + relu(acc + bo) for one 128-channel half of the block's rows -> y planes (+ fp32)
  auto y_epilogue = [&](const int h, const int wq) {
    const uint32_t tq_row = tmem + ((uint32_t)(wq * 32) << 16);
    const uint32_t stg = stg_all + (uint32_t)wq * 4096;
    const int cr = lane >> 2, cp = lane & 3, fr = lane >> 3, fp = lane & 7;
    const uint32_t own16 = stg + (uint32_t)lane * 64, sw_own = (uint32_t)((lane >> 1) & 3);
    const size_t rowbase = row0 + qb * 128 + wq * 32;
    const __half* xhb = static_cast<const __half*>(p.x_hi);
    const __half* xlb = static_cast<const __half*>(p.x_lo);
    __half* yhb = static_cast<__half*>(p.y_hi);
    __half* ylb = static_cast<__half*>(p.y_lo);
    const float* bsm = reinterpret_cast<const float*>(gen + (bo_s - base));
    uint4 pre[8];
    auto x_fetch = [&](int cb) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const size_t off = (rowbase + 8 * j + cr) * p.ldx + cb + cp * 8;
        pre[j] = *reinterpret_cast<const uint4*>(xhb + off);
        pre[4 + j] = *reinterpret_cast<const uint4*>(xlb + off);
      }
    };
    x_fetch(h * 128);
    mbar_wait(bar_y + 8 * h, 0);
    if (chain) mbar_wait(bar_y + 8, 0);   // the output goes over r in region A: every out-projection MMA must have completed
    tc_fence_after();
    if (lane == 0 && wq == 0) as_stamp(prof, 8 + 2 * h);
#pragma unroll 1
    for (int c32 = 0; c32 < 4; ++c32) {
      const int cb = h * 128 + c32 * 32;
#pragma unroll
      for (int j = 0; j < 4; ++j) {   // the prefetched residual block -> staging
        const int r = 8 * j + cr;
        const uint32_t a = stg + (uint32_t)r * 64 + (uint32_t)((cp ^ ((r >> 1) & 3)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pre[j].x), "r"(pre[j].y), "r"(pre[j].z), "r"(pre[j].w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(pre[4 + j].x), "r"(pre[4 + j].y), "r"(pre[4 + j].z), "r"(pre[4 + j].w) : "memory");
      }
      __syncwarp();
      if (c32 + 1 < 4) x_fetch(cb + 32);
      float v[32];
      tmem_ld32(tq_row + cb, v);        // the out-projection accumulators sit over S: columns [0, 256)
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i] + bsm[cb + i], 0.f);
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 xh4, xl4;
        const uint32_t a = own16 + (((uint32_t)q4 ^ sw_own) << 4);
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xh4.x), "=r"(xh4.y), "=r"(xh4.z), "=r"(xh4.w) : "r"(a));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(xl4.x), "=r"(xl4.y), "=r"(xl4.z), "=r"(xl4.w) : "r"(a + 2048));
        const uint32_t* hp = reinterpret_cast<const uint32_t*>(&xh4);
        const uint32_t* lp = reinterpret_cast<const uint32_t*>(&xl4);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a2 = __half22float2(*reinterpret_cast<const __half2*>(&hp[e]));
          const float2 b2 = __half22float2(*reinterpret_cast<const __half2*>(&lp[e]));
          v[q4 * 8 + 2 * e] += a2.x + b2.x;
          v[q4 * 8 + 2 * e + 1] += a2.y + b2.y;
        }
      }
      __syncwarp();                   // every lane has read its residual row: the tile takes the output block
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        uint4 oh, ol;
        split2h(v[q4 * 8 + 0], v[q4 * 8 + 1], oh.x, ol.x);
        split2h(v[q4 * 8 + 2], v[q4 * 8 + 3], oh.y, ol.y);
        split2h(v[q4 * 8 + 4], v[q4 * 8 + 5], oh.z, ol.z);
        split2h(v[q4 * 8 + 6], v[q4 * 8 + 7], oh.w, ol.w);
        const uint32_t a = own16 + (((uint32_t)q4 ^ sw_own) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(oh.x), "r"(oh.y), "r"(oh.z), "r"(oh.w) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(ol.x), "r"(ol.y), "r"(ol.z), "r"(ol.w) : "memory");
        if (chain) {   // ... and the same planes as the K-major A operand of the next layer's projections (region A, over r)
          const uint32_t xk = base + (uint32_t)(cb >> 6) * T16 + sw128(wq * 32 + lane, ((cb & 63) >> 3) + q4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xk), "r"(oh.x), "r"(oh.y), "r"(oh.z), "r"(oh.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(xk + 4 * T16), "r"(ol.x), "r"(ol.y), "r"(ol.z), "r"(ol.w) : "memory");
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = 8 * j + cr;
        const uint32_t a = stg + (uint32_t)r * 64 + (uint32_t)((cp ^ ((r >> 1) & 3)) << 4);
        uint4 hh, ll;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hh.x), "=r"(hh.y), "=r"(hh.z), "=r"(hh.w) : "r"(a));
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ll.x), "=r"(ll.y), "=r"(ll.z), "=r"(ll.w) : "r"(a + 2048));
        const size_t off = (rowbase + r) * p.ldy + cb + cp * 8;
        *reinterpret_cast<uint4*>(yhb + off) = hh;
        *reinterpret_cast<uint4*>(ylb + off) = ll;
      }
      __syncwarp();
      if (p.yf) {                     // fp32 rows (need=True): [32 rows][128 B] through the same tile
#pragma unroll
        for (int p8 = 0; p8 < 8; ++p8) {
          const uint32_t a = stg + (uint32_t)lane * 128 + (uint32_t)((p8 ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v[p8 * 4]), "f"(v[p8 * 4 + 1]), "f"(v[p8 * 4 + 2]), "f"(v[p8 * 4 + 3]) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int r = 4 * j + fr;
          const uint32_t a = stg + (uint32_t)r * 128 + (uint32_t)((fp ^ (r & 7)) << 4);
          float4 o4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o4.x), "=f"(o4.y), "=f"(o4.z), "=f"(o4.w) : "r"(a));
          *reinterpret_cast<float4*>(p.yf + (rowbase + r) * p.ldyf + cb + fp * 4) = o4;
        }
        __syncwarp();
      }
    }
    if (lane == 0 && wq == 0) as_stamp(prof, 9 + 2 * h);
  };
  // chained projections epilogue, block b3 = 0: next q|k planes [rows, 128]; b3 = 1, 2: next v^T planes (channels
  // (b3 - 1) * 128 ..), transposed: the 32 lanes of a warp are 32 consecutive tokens = 64 contiguous bytes per channel
  auto z_epilogue = [&](const int b3, const int wq) {
    const uint32_t tcol = b3 == 0 ? 0u : 128u + 128u * (uint32_t)b3;   // accumulator columns: [0,128), [256,384), [384,512)
    const uint32_t tq_row = tmem + ((uint32_t)(wq * 32) << 16) + tcol;
    const uint32_t stg = stg_all + (uint32_t)wq * 4096;
    const int cr = lane >> 2, cp = lane & 3;
    const uint32_t own16 = stg + (uint32_t)lane * 64, sw_own = (uint32_t)((lane >> 1) & 3);
    const size_t rowbase = row0 + qb * 128 + wq * 32;
    const float* bsm = reinterpret_cast<const float*>(gen + (bo_s - base)) + 256 + b3 * 128;
    mbar_wait(bar_z + 8 * b3, 0);
    tc_fence_after();
    if (lane == 0 && wq == 0) as_stamp(prof, 12 + b3);
#pragma unroll 1
    for (int c32 = 0; c32 < 4; ++c32) {
      float v[32];
      tmem_ld32(tq_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += bsm[c32 * 32 + i];
      if (b3 == 0) {
        __half* qh = static_cast<__half*>(p.qk2_hi);
        __half* ql = static_cast<__half*>(p.qk2_lo);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint4 oh, ol;
          split2h(v[q4 * 8 + 0], v[q4 * 8 + 1], oh.x, ol.x);
          split2h(v[q4 * 8 + 2], v[q4 * 8 + 3], oh.y, ol.y);
          split2h(v[q4 * 8 + 4], v[q4 * 8 + 5], oh.z, ol.z);
          split2h(v[q4 * 8 + 6], v[q4 * 8 + 7], oh.w, ol.w);
          const uint32_t a = own16 + (((uint32_t)q4 ^ sw_own) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(oh.x), "r"(oh.y), "r"(oh.z), "r"(oh.w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048), "r"(ol.x), "r"(ol.y), "r"(ol.z), "r"(ol.w) : "memory");
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = 8 * j + cr;
          const uint32_t a = stg + (uint32_t)r * 64 + (uint32_t)((cp ^ ((r >> 1) & 3)) << 4);
          uint4 hh, ll;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(hh.x), "=r"(hh.y), "=r"(hh.z), "=r"(hh.w) : "r"(a));
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ll.x), "=r"(ll.y), "=r"(ll.z), "=r"(ll.w) : "r"(a + 2048));
          const size_t off = (rowbase + r) * 128 + c32 * 32 + cp * 8;
          *reinterpret_cast<uint4*>(qh + off) = hh;
          *reinterpret_cast<uint4*>(ql + off) = ll;
        }
        __syncwarp();
      } else {
        const int ch0 = (b3 - 1) * 128 + c32 * 32;
        const size_t tok = (size_t)qb * 128 + wq * 32 + lane;
        __half* dh = static_cast<__half*>(p.vT2_hi) + ((size_t)cloud * AS_C + ch0) * AS_L + tok;
        __half* dl = static_cast<__half*>(p.vT2_lo) + ((size_t)cloud * AS_C + ch0) * AS_L + tok;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          uint32_t hi, lo;
          split2h(v[i], v[i + 1], hi, lo);
          const __half2 h2 = *reinterpret_cast<const __half2*>(&hi), l2 = *reinterpret_cast<const __half2*>(&lo);
          dh[(size_t)i * AS_L] = __low2half(h2);
          dh[(size_t)(i + 1) * AS_L] = __high2half(h2);
          dl[(size_t)i * AS_L] = __low2half(l2);
          dl[(size_t)(i + 1) * AS_L] = __high2half(l2);
        }
      }
    }
  };
  if (warp >= 5) {
    // =========================================================== v^T producer: ONE thread, 8 stage loads (2 channel halves
    // x 4 key blocks), two TMA tiles (hi, lo) per stage
    if (warp == 5 && lane == 0) {
      const int vrow = cloud * AS_C;
      // P lives in tensor memory (over S), so region A is free for v^T while q k^T and the softmax run: stages 0, 1 and 6, 7
      // go through the ring, stage 2 into tiles 6, 7 at once, stages 3..5 into the q / k tiles as soon as S is complete --
      // six of the eight stages are in flight or landed before the first P v MMA can be issued
      for (uint32_t it = 0; it < 8; ++it) {
        const int h = it >> 2, kb = it & 3;
        uint32_t st, fb;
        if (it >= 2 && it < 6) {
          if (it == 3) mbar_wait(bar_s, 0);   // q and k have been read
          st = base + AS_VSLOT(it) * T16;
          fb = bar_va + 8 * (it - 2);
        } else {
          const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          st = ring + s * AS_STAGE;
          fb = full_bar + 8 * s;
        }
        as_expect_tx(fb, AS_STAGE);
        as_tma_load(st, &maps.vT[0], kb * 64, vrow + h * 128, fb);
        as_tma_load(st + T16, &maps.vT[1], kb * 64, vrow + h * 128, fb);
      }
      if (fuse) {   // 8 more stage loads: Wo [128 out channels x 64 k] x 2 planes per (channel half, k-block)
        for (uint32_t it = 8; it < 16; ++it) {
          const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
          const int kb = (it - 8) >> 1, h = (it - 8) & 1;   // k-block major: k-blocks 0, 1 only need r half 0
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          const uint32_t st = ring + s * AS_STAGE;
          as_expect_tx(full_bar + 8 * s, AS_STAGE);
          as_tma_load(st, &maps.wo[wset][0], kb * 64, h * 128, full_bar + 8 * s);
          as_tma_load(st + T16, &maps.wo[wset][1], kb * 64, h * 128, full_bar + 8 * s);
        }
      }
      if (chain) {   // 12 more: the next layer's Wqkv [128 rows x 64 k] x 2 planes per (k-block, block of 128 output rows)
        for (uint32_t it = 16; it < 28; ++it) {
          const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
          // k-blocks 0, 1 block-interleaved (they run under the second out epilogue); k-blocks 2, 3 block-major, so that a
          // block completes every two stages and its epilogue overlaps the remaining MMAs
          const int j = (int)it - 16, kb = j < 6 ? j / 3 : 2 + ((j - 6) & 1), b3 = j < 6 ? j % 3 : (j - 6) >> 1;
          mbar_wait(empty_bar + 8 * s, ph ^ 1);
          const uint32_t st = ring + s * AS_STAGE;
          as_expect_tx(full_bar + 8 * s, AS_STAGE);
          as_tma_load(st, &maps.wn[wset][0], kb * 64, b3 * 128, full_bar + 8 * s);
          as_tma_load(st + T16, &maps.wn[wset][1], kb * 64, b3 * 128, full_bar + 8 * s);
        }
      }
    }
  } else if (warp == 4) {
    // =========================================================== MMA issuer
    if (elect_one()) {
      {  // S[i, j] = sum_d q[i, d] k[j, d]
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        as_stamp(prof, 1);
        const uint32_t idesc = idesc_f16(256);
        const uint64_t a_hi = make_desc(q_hi_s), a_lo = make_desc(q_lo_s), b_hi = make_desc(k_hi_s), b_lo = make_desc(k_lo_s);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) umma3(tmem, a_hi + 2 * k4, a_lo + 2 * k4, b_hi + 2 * k4, b_lo + 2 * k4, idesc, k4 != 0);
        umma_commit(bar_s);
      }
      const uint32_t idesc = idesc_f16(128);
      for (uint32_t it = 0; it < 8; ++it) {
        const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
        const int h = it >> 2, kb = it & 3;
        const bool in_a = it >= 2 && it < 6;
        if (h == 0) mbar_wait(bar_pk + 8 * kb, 0);   // P's key block kb is in tensor memory (the softmax is still writing later ones)
        if (in_a) mbar_wait(bar_va + 8 * (it - 2), 0);
        else mbar_wait(full_bar + 8 * s, ph);
        tc_fence_after();
        const uint32_t st = in_a ? base + AS_VSLOT(it) * T16 : ring + s * AS_STAGE;
        const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + T16);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          // keys 64 kb + 16 k4 .. + 15 of P: chunk 2 kb + (k4 >> 1) of 32 columns = [16 hi columns | 16 lo columns]
          const uint32_t a_hi = tmem + (uint32_t)(2 * kb + (k4 >> 1)) * 32 + (uint32_t)(k4 & 1) * 8, a_lo = a_hi + 16;
          umma_ts(tmem + 256 + h * 128, a_hi, b_hi + 2 * k4, idesc, (kb | k4) != 0);
          umma_ts(tmem + 256 + h * 128, a_hi, b_lo + 2 * k4, idesc, 1);
          umma_ts(tmem + 256 + h * 128, a_lo, b_hi + 2 * k4, idesc, 1);
        }
        if (!in_a) umma_commit(empty_bar + 8 * s);
        if (kb == 3) umma_commit(bar_o + 8 * h);
      }
      if (fuse) {   // out[i, c] = sum_k r[i, k] Wo[c, k]: r planes in region A (over P), accumulators over S
        for (uint32_t it = 8; it < 16; ++it) {
          const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
          const int kb = (it - 8) >> 1, h = (it - 8) & 1;
          if (it == 8) mbar_wait(bar_r, 0);        // r channels 0..127 = k-blocks 0, 1 (under the epilogue of r half 1)
          if (it == 12) mbar_wait(bar_r + 8, 0);   // r channels 128..255 = k-blocks 2, 3
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = ring + s * AS_STAGE;
          const uint64_t a_hi = make_desc(base + kb * T16), a_lo = make_desc(base + 4 * T16 + kb * T16);
          const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + T16);
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma3(tmem + h * 128, a_hi + 2 * k4, a_lo + 2 * k4, b_hi + 2 * k4, b_lo + 2 * k4, idesc, (kb | k4) != 0);
          umma_commit(empty_bar + 8 * s);
          if (kb == 3) umma_commit(bar_y + 8 * h);
        }
      }
      if (chain) {   // [q | k | v][i, c] = sum_k out[i, k] Wqkv[c, k]: out planes in region A (over r); accumulators: q|k over
                     // columns [0, 128) (the out half the second epilogue has drained), v over the O halves [256, 512)
        for (uint32_t it = 16; it < 28; ++it) {
          const uint32_t s = it % AS_NST, ph = (it / AS_NST) & 1;
          const int j = (int)it - 16, kb = j < 6 ? j / 3 : 2 + ((j - 6) & 1), b3 = j < 6 ? j % 3 : (j - 6) >> 1;
          if (it == 16) mbar_wait(bar_x2, 0);        // out channels 0..127 = k-blocks 0, 1
          if (it == 22) mbar_wait(bar_x2 + 8, 0);    // out channels 128..255 = k-blocks 2, 3
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t st = ring + s * AS_STAGE;
          const uint64_t a_hi = make_desc(base + kb * T16), a_lo = make_desc(base + 4 * T16 + kb * T16);
          const uint64_t b_hi = make_desc(st), b_lo = make_desc(st + T16);
          const uint32_t dcol = b3 == 0 ? 0u : 128u + 128u * (uint32_t)b3;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma3(tmem + dcol, a_hi + 2 * k4, a_lo + 2 * k4, b_hi + 2 * k4, b_lo + 2 * k4, idesc, (kb | k4) != 0);
          umma_commit(empty_bar + 8 * s);
          if (kb == 3) umma_commit(bar_z + 8 * b3);
        }
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
    if (tid == 0) as_stamp(prof, 2);
    float m = -INFINITY;
#pragma unroll 1
    for (int c32 = 0; c32 < 8; ++c32) {
      float v[32];
      tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
    }
    const float mc = m * cexp;
    if (p.attn_mode != 0) {  // attention map (need=True): mean of the four layers' maps, model5_b.py:468-469.  Two extra passes over
      // S BEFORE the pass below overwrites it with P: the row sum (same order of additions as below), then the map
      float s0 = 0.f;
#pragma unroll 1
      for (int c32 = 0; c32 < 8; ++c32) {
        float v[32];
        tmem_ld32(t_row + c32 * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) s0 += exp2f(fmaf(v[i], cexp, -mc));
      }
      const float inv = 1.0f / s0;
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
      // un-normalised probabilities -> the A operand of P v, IN PLACE over the 32 S columns just read: 16 columns of hi
      // pairs (keys 2 i, 2 i + 1 in column i), then 16 columns of lo pairs
      uint32_t pr[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) split2h(v[2 * i], v[2 * i + 1], pr[i], pr[16 + i]);
      tmem_st32(t_row + c32 * 32, pr);
      if (c32 & 1) {   // key block c32 / 2 complete: the P v MMAs over it may start under the rest of the pass
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        mbar_arrive(bar_pk + 8 * (c32 >> 1));
      }
    }
    if (tid == 0) as_stamp(prof, 3);
    const float inv = 1.0f / sum;
    // r = x - O / sum, one 128-channel half at a time (the second half's MMAs run under the first half's epilogue)
    // (a second group of four warps taking the other half was measured: both halves slow down to the same total -- the
    // phase moves 256 KB per CTA at ~5.4 TB/s over the chip, it is HBM-bound)
    r_epilogue(0, warp, inv);
    if (fuse) {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_r);
    }
    r_epilogue(1, warp, inv);
    if (fuse) {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bar_r + 8);
      y_epilogue(0, warp);
      if (chain) {
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_x2);
      }
      y_epilogue(1, warp);
      if (chain) {
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bar_x2 + 8);
        z_epilogue(0, warp);
        z_epilogue(1, warp);
        z_epilogue(2, warp);
        if (tid == 0) as_stamp(prof, 15);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && p.sig_flags != nullptr) {   // every store of this CTA is ordered before the barrier: release them to the next layer
    __threadfence();
    asm volatile("red.release.gpu.global.add.s32 [%0], 1;" ::"l"(p.sig_flags + cloud) : "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
  }
}

// 2-D fp16 view [rows, cols] (row stride ld elements) traversed in [128 rows x 64 columns] SWIZZLE_128B boxes
static int as_make_map(const void* ptr, int ld, size_t rows, int cols, CUtensorMap* out) {
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
  PZ_REQUIRE(encode != nullptr, PZ_ERR_UNSUPPORTED, "attention_split: the driver does not export cuTensorMapEncodeTiled");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(__half)};
  const cuuint32_t box[2] = {64, 128}, estr[2] = {1, 1};
  const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PZ_REQUIRE(r == CUDA_SUCCESS, PZ_ERR_ARG, "attention_split: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int launch_attention_split(const AttnSplit& p, int clouds, cudaStream_t st) {
  const bool fuse = p.wo_hi[0] != nullptr;
  PZ_REQUIRE(p.qk_hi && p.qk_lo && p.vT_hi && p.vT_lo && p.x_hi && p.x_lo && (fuse || (p.r_hi && p.r_lo)), PZ_ERR_ARG, "attention_split: null pointer");
  if (fuse)
    PZ_REQUIRE(p.wo_lo[0] && p.bo[0] && p.y_hi && p.y_lo && p.ldy % 8 == 0 && (((uintptr_t)p.y_hi | (uintptr_t)p.y_lo | (uintptr_t)p.wo_hi[0] | (uintptr_t)p.wo_lo[0]) & 15) == 0 &&
                   (p.clouds_per_set <= 0 || p.clouds_per_set >= clouds || (p.wo_hi[1] && p.wo_lo[1] && p.bo[1])) &&
                   (!p.yf || (p.ldyf % 4 == 0 && ((uintptr_t)p.yf & 15) == 0)),
               PZ_ERR_ARG, "attention_split: fused out-projection needs Wo planes, bias and 16-byte aligned output planes");
  PZ_REQUIRE(p.ldx % 8 == 0 && (((uintptr_t)p.x_hi | (uintptr_t)p.x_lo | (uintptr_t)p.r_hi | (uintptr_t)p.r_lo | (uintptr_t)p.qk_hi |
                                 (uintptr_t)p.qk_lo | (uintptr_t)p.vT_hi | (uintptr_t)p.vT_lo) & 15) == 0,
             PZ_ERR_ARG, "attention_split: rows must be 16-byte aligned");
  PZ_REQUIRE(p.attn_mode == 0 || p.attn, PZ_ERR_ARG, "attention_split: attention map requested without a buffer");
  const bool chain = fuse && p.wqkv_hi[0] != nullptr;
  if (chain)
    PZ_REQUIRE(p.wqkv_lo[0] && p.bqkv[0] && p.qk2_hi && p.qk2_lo && p.vT2_hi && p.vT2_lo &&
                   (((uintptr_t)p.wqkv_hi[0] | (uintptr_t)p.wqkv_lo[0] | (uintptr_t)p.qk2_hi | (uintptr_t)p.qk2_lo | (uintptr_t)p.vT2_hi | (uintptr_t)p.vT2_lo) & 15) == 0 &&
                   (p.clouds_per_set <= 0 || p.clouds_per_set >= clouds || (p.wqkv_hi[1] && p.wqkv_lo[1] && p.bqkv[1])) &&
                   p.qk2_hi != p.qk_hi && p.vT2_hi != p.vT_hi,
               PZ_ERR_ARG, "attention_split: chained projections need next-layer weights, bias and separate output planes");
  const size_t smem = 1024 + REGA + AS_NST * AS_STAGE + 8 * (22 + 2 * AS_NST) + 32 + 16 + 1024 + 1536 + 128 + 4 * 4096;
  static_assert(1024 + REGA + AS_NST * AS_STAGE + 8 * (22 + 2 * AS_NST) + 32 + 16 + 1024 + 1536 + 128 + 4 * 4096 <= 232448, "attention_split: shared memory budget");
  PZ_CUDA(cudaFuncSetAttribute(attention_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AsMaps maps;
  const size_t rows = (size_t)clouds * AS_L;
  PZ_TRY(as_make_map(p.qk_hi, 128, rows, 128, &maps.qk[0]));
  PZ_TRY(as_make_map(p.qk_lo, 128, rows, 128, &maps.qk[1]));
  PZ_TRY(as_make_map(p.vT_hi, AS_L, (size_t)clouds * AS_C, AS_L, &maps.vT[0]));
  PZ_TRY(as_make_map(p.vT_lo, AS_L, (size_t)clouds * AS_C, AS_L, &maps.vT[1]));
  for (int ws = 0; ws < 2; ++ws) {   // without fusion (or with one weight set) the maps repeat a valid tensor and are never used
    const void* wh = fuse ? (p.wo_hi[ws] ? p.wo_hi[ws] : p.wo_hi[0]) : p.vT_hi;
    const void* wl = fuse ? (p.wo_lo[ws] ? p.wo_lo[ws] : p.wo_lo[0]) : p.vT_lo;
    PZ_TRY(as_make_map(wh, AS_C, AS_C, AS_C, &maps.wo[ws][0]));
    PZ_TRY(as_make_map(wl, AS_C, AS_C, AS_C, &maps.wo[ws][1]));
    const void* nh = chain ? (p.wqkv_hi[ws] ? p.wqkv_hi[ws] : p.wqkv_hi[0]) : wh;
    const void* nl = chain ? (p.wqkv_lo[ws] ? p.wqkv_lo[ws] : p.wqkv_lo[0]) : wl;
    PZ_TRY(as_make_map(nh, AS_C, chain ? 384 : AS_C, AS_C, &maps.wn[ws][0]));
    PZ_TRY(as_make_map(nl, AS_C, chain ? 384 : AS_C, AS_C, &maps.wn[ws][1]));
  }
  static const bool tl_chain_only = getenv("PZ_AS_TL_CHAIN") != nullptr;   // diagnostics: only launches with chained projections stamp
  long long* prof = (tl_chain_only && !chain) ? nullptr : kernel_timeline_buffer(3072 + 16);
  PZ_CUDA(launch_pdl(attention_split_kernel, dim3(2 * clouds), dim3(AS_THREADS), smem, st, p, maps, prof));
  PZ_LAUNCH_CHECK();
  return 0;
}

}  // namespace pz
