// tcgen05 GEMM with ROW-MAJOR epilogue for the plain nn.Linear layers of the bf16 path:
//   Y[row, ch] = act(X[row,:] . W[ch,:] + b[ch] (+ W1x[ch,:] . xyz[row]) (+ rowbias[row / rb_rows, ch])) (+ R[row, ch])
// Same operands as gemm_tc.cu (both K-major SWIZZLE_128B), but with the roles swapped: activation rows on the
// UMMA M dimension (TMEM lanes -> one epilogue thread per ROW), output channels on UMMA N (TMEM columns).
// A thread therefore holds 32 consecutive channels of its row per tcgen05.ld and writes / reads them with
// 128-bit accesses (64 B of bf16 per chunk), instead of one 2-byte store per element in the channel-per-thread
// orientation -- that orientation is kept for the max-pool epilogues (gemm_tc.cu), this one for plain stores.
// Per-channel constants (bias, the three xyz weights) are staged once per tile as one float4 per channel in
// shared memory and read as broadcasts.
#include "pz_common.cuh"
#include "tc_common.cuh"

namespace pz {

using namespace tc;

namespace {
constexpr int RG_THREADS = 13 * 32;   // 8 epilogue warps, 1 MMA warp, 4 producer warps
constexpr int RG_EPI = 8, RG_PROD_THREADS = 128;
constexpr int RG_ROWS = 128;          // UMMA M
}  // namespace

template <int NCOLS, int NST>
__global__ void __launch_bounds__(RG_THREADS, 1) tc_rowgemm_kernel(const TcGemm g) {
  extern __shared__ __align__(1024) uint8_t rg_smem_raw[];
  const uint32_t smem_base = (smem_u32(rg_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = rg_smem_raw + (smem_base - smem_u32(rg_smem_raw));
  constexpr uint32_t STAGE_X = RG_ROWS * 128, STAGE_W = NCOLS * 128, STAGE = STAGE_X + STAGE_W;
  const uint32_t bars = smem_base + NST * STAGE;
  const uint32_t full_bar = bars, empty_bar = bars + 8 * NST, accf_bar = bars + 16 * NST, acce_bar = accf_bar + 16;
  const uint32_t tmem_slot = acce_bar + 16;
  const uint32_t chan_s = tmem_slot + 16;           // [NCOLS] float4 (w1x, w1y, w1z, bias), 16-byte aligned
  float4* chan = reinterpret_cast<float4*>(smem_gen + (chan_s - smem_base));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kblocks = g.K / KB;
  const int nsets = (g.rows_per_wset > 0 && g.M > g.rows_per_wset) ? 2 : 1;
  const int row_tiles = g.M / RG_ROWS, col_tiles = g.Nout / NCOLS;
  const int tiles_per_set = (row_tiles / nsets) * col_tiles;
  const int ctas_per_set = gridDim.x / nsets;
  const int wset = min((int)blockIdx.x / ctas_per_set, nsets - 1);
  const int rank_in_set = blockIdx.x - wset * ctas_per_set;
  const int step = (wset == nsets - 1) ? (int)gridDim.x - wset * ctas_per_set : ctas_per_set;
  const int tile_begin = wset * tiles_per_set;
  const __nv_bfloat16* __restrict__ W = g.W[wset];
  const float* __restrict__ bias = g.bias[wset];

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(full_bar + 8 * s, RG_PROD_THREADS);
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accf_bar + 8 * b, 1);
      mbar_init(acce_bar + 8 * b, RG_EPI * 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == RG_EPI) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));

  if (warp > RG_EPI) {
    // ================================================= producers: cp.async both operands, 2 stages in flight
    const int pt = tid - (RG_EPI + 1) * 32;
    uint32_t issued = 0, arrived = 0;
    for (int t = tile_begin + rank_in_set; t < tile_begin + tiles_per_set; t += step) {
      const int ct = t % col_tiles, rt = t / col_tiles;
      const int row0 = rt * RG_ROWS, col0 = ct * NCOLS;
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t s = issued % NST, ph = (issued / NST) & 1;
        mbar_wait(empty_bar + 8 * s, ph ^ 1);
        const uint32_t st_addr = smem_base + s * STAGE;
        for (int id = pt; id < RG_ROWS * 8; id += RG_PROD_THREADS) {
          const int c = id & 7, r = id >> 3;
          cp_async16(st_addr + sw128(r, c), g.X + (size_t)(row0 + r) * g.ldx + kb * KB + c * 8);
        }
        for (int id = pt; id < NCOLS * 8; id += RG_PROD_THREADS) {
          const int c = id & 7, r = id >> 3;
          cp_async16(st_addr + STAGE_X + sw128(r, c), W + (size_t)(col0 + r) * g.ldw + kb * KB + c * 8);
        }
        cp_async_commit();
        ++issued;
        if (issued - arrived > 2) {
          cp_async_wait<2>();
          fence_proxy_async();
          mbar_arrive(full_bar + 8 * (arrived % NST));
          ++arrived;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    while (arrived < issued) {
      mbar_arrive(full_bar + 8 * (arrived % NST));
      ++arrived;
    }
  } else if (warp == RG_EPI) {
    // ================================================= MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc(NCOLS);
      uint32_t it = 0, tcn = 0;
      for (int t = tile_begin + rank_in_set; t < tile_begin + tiles_per_set; t += step, ++tcn) {
        const uint32_t buf = tcn & 1, aph = (tcn >> 1) & 1;
        mbar_wait(acce_bar + 8 * buf, aph ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < kblocks; ++kb, ++it) {
          const uint32_t s = it % NST, ph = (it / NST) & 1;
          mbar_wait(full_bar + 8 * s, ph);
          tc_fence_after();
          const uint32_t st_addr = smem_base + s * STAGE;
          const uint64_t adesc = make_desc(st_addr), bdesc = make_desc(st_addr + STAGE_X);
#pragma unroll
          for (int k4 = 0; k4 < KB / 16; ++k4)
            umma_bf16(tmem_base + buf * NCOLS, adesc + 2 * k4, bdesc + 2 * k4, idesc, (kb | k4) != 0);
          umma_commit(empty_bar + 8 * s);
        }
        umma_commit(accf_bar + 8 * buf);
      }
    }
  } else {
    // ================================================= epilogue: thread = row, 32 channels per TMEM load
    const int quarter = warp & 3, half = warp >> 2;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const bool has_xyz = g.xyz != nullptr, relu = g.relu != 0;
    const int nvalid = g.n_valid > 0 ? g.n_valid : g.Nout;
    uint32_t tcn = 0;
    int staged_ct = -1;
    for (int t = tile_begin + rank_in_set; t < tile_begin + tiles_per_set; t += step, ++tcn) {
      const int ct = t % col_tiles, rt = t / col_tiles;
      const int col0 = ct * NCOLS;
      const size_t row = (size_t)rt * RG_ROWS + quarter * 32 + lane;
      if (ct != staged_ct) {   // per-channel constants of this column tile -> smem (only changes when col_tiles > 1)
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = tid; c < NCOLS; c += RG_EPI * 32) {
          const int ch = col0 + c;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch < nvalid) {
            if (bias) v.w = bias[ch];
            if (has_xyz) {
              const float* wp = g.W1x[wset] + (size_t)ch * g.ldw1x;
              v.x = wp[0]; v.y = wp[1]; v.z = wp[2];
            }
          }
          chan[c] = v;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        staged_ct = ct;
      }
      float px = 0.f, py = 0.f, pz = 0.f;
      if (has_xyz) {
        const float* p = g.xyz + row * 3;
        px = p[0]; py = p[1]; pz = p[2];
      }
      const float* rbp = g.rowbias ? g.rowbias + (row / g.rb_rows) * g.rb_ld : nullptr;
      const uint32_t buf = tcn & 1, aph = (tcn >> 1) & 1;
      mbar_wait(accf_bar + 8 * buf, aph);
      tc_fence_after();
#pragma unroll 1
      for (int c32 = half; c32 < NCOLS / 32; c32 += 2) {
        const int cb = col0 + c32 * 32;              // first channel of this chunk
        if (cb >= nvalid) break;                     // zero-padded channels: nothing to store
        uint4 rv[4];
        if (g.Rb) {
          const uint4* rp = reinterpret_cast<const uint4*>(g.Rb + row * g.ldrb + cb);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) rv[q4] = rp[q4];
        }
        float v[32];
        tmem_ld32(tmem_base + lane_base + buf * NCOLS + c32 * 32, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float4 cc = chan[c32 * 32 + i];
          float x = v[i] + cc.w;
          if (has_xyz) x = fmaf(cc.x, px, fmaf(cc.y, py, fmaf(cc.z, pz, x)));
          v[i] = x;
        }
        if (rbp) {
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(rbp + cb + q4 * 4);
            v[q4 * 4] += b4.x; v[q4 * 4 + 1] += b4.y; v[q4 * 4 + 2] += b4.z; v[q4 * 4 + 3] += b4.w;
          }
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (g.Rb) {
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&rv[q4]);
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const float2 f = __bfloat1622float2(xp[h]);
              v[q4 * 8 + 2 * h] += f.x;
              v[q4 * 8 + 2 * h + 1] += f.y;
            }
          }
        }
        if (g.Rf) {
          const float4* rp = reinterpret_cast<const float4*>(g.Rf + row * g.ldrf + cb);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 b4 = rp[q4];
            v[q4 * 4] += b4.x; v[q4 * 4 + 1] += b4.y; v[q4 * 4 + 2] += b4.z; v[q4 * 4 + 3] += b4.w;
          }
        }
        if (g.Yb) {
          uint4* yp = reinterpret_cast<uint4*>(g.Yb + row * g.ldyb + cb);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint4 o;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[q4 * 8 + 0], v[q4 * 8 + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[q4 * 8 + 2], v[q4 * 8 + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[q4 * 8 + 4], v[q4 * 8 + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[q4 * 8 + 6], v[q4 * 8 + 7]);
            o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
            o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
            yp[q4] = o;
          }
        }
        if (g.Yf) {
          float4* yp = reinterpret_cast<float4*>(g.Yf + row * g.ldyf + cb);
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) yp[q4] = make_float4(v[q4 * 4], v[q4 * 4 + 1], v[q4 * 4 + 2], v[q4 * 4 + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(acce_bar + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RG_EPI) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int NCOLS, int NST>
static int rowgemm_launch(const TcGemm& g, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)NST * (RG_ROWS * 128 + NCOLS * 128) + 8 * (2 * NST + 4) + 32 + (size_t)NCOLS * 16;
  auto kern = tc_rowgemm_kernel<NCOLS, NST>;
  PZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int nsets = (g.rows_per_wset > 0 && g.M > g.rows_per_wset) ? 2 : 1;
  const int tiles = (g.M / RG_ROWS) * (g.Nout / NCOLS);
  int grid = tiles < kNumSMs ? tiles : kNumSMs;
  if (nsets == 2 && (grid & 1)) --grid;
  if (grid < nsets) grid = nsets;
  kern<<<grid, RG_THREADS, smem, st>>>(g);
  PZ_LAUNCH_CHECK();
  return 0;
}

// Row-major-output tcgen05 GEMM.  Supported: M % (128*sets) == 0, K % 64 == 0, Nout % 128 == 0, 16-byte aligned rows
// of every operand / output / residual; epi must be 0 (plain store); no YT / Ymax (those live in gemm_tc.cu).
int launch_tc_rowgemm(const TcGemm& g, cudaStream_t st) {
  PZ_REQUIRE(g.W[0] && g.X && (g.Yf || g.Yb), PZ_ERR_ARG, "tc_rowgemm: null operand");
  PZ_REQUIRE(g.epi == 0 && !g.YT && !g.Ymax && !g.rows, PZ_ERR_ARG, "tc_rowgemm: plain-store epilogue only");
  PZ_REQUIRE(g.K % 64 == 0 && g.Nout % 128 == 0 && g.ldx % 8 == 0 && g.ldw % 8 == 0, PZ_ERR_UNSUPPORTED,
             "tc_rowgemm: needs K %% 64 == 0, Nout %% 128 == 0 and 16-byte aligned rows (K=%d Nout=%d)", g.K, g.Nout);
  PZ_REQUIRE((!g.Yb || (g.ldyb % 8 == 0 && ((uintptr_t)g.Yb & 15) == 0)) && (!g.Yf || (g.ldyf % 4 == 0 && ((uintptr_t)g.Yf & 15) == 0)) &&
                 (!g.Rb || (g.ldrb % 8 == 0 && ((uintptr_t)g.Rb & 15) == 0)) && (!g.Rf || (g.ldrf % 4 == 0 && ((uintptr_t)g.Rf & 15) == 0)) &&
                 (!g.rowbias || (g.rb_ld % 4 == 0 && ((uintptr_t)g.rowbias & 15) == 0)),
             PZ_ERR_ARG, "tc_rowgemm: outputs / residuals / rowbias must have 16-byte aligned rows");
  PZ_REQUIRE(((uintptr_t)g.X & 15) == 0 && ((uintptr_t)g.W[0] & 15) == 0 && (!g.W[1] || ((uintptr_t)g.W[1] & 15) == 0),
             PZ_ERR_ARG, "tc_rowgemm: operands must be 16-byte aligned");
  const int nsets = (g.rows_per_wset > 0 && g.M > g.rows_per_wset) ? 2 : 1;
  if (nsets == 2) PZ_REQUIRE(g.W[1] && g.M == 2 * g.rows_per_wset, PZ_ERR_ARG, "tc_rowgemm: two weight sets need M == 2*rows_per_wset");
  PZ_REQUIRE(g.M % (RG_ROWS * nsets) == 0, PZ_ERR_UNSUPPORTED, "tc_rowgemm: M=%d must be a multiple of %d", g.M, RG_ROWS * nsets);
  if (g.n_valid > 0) PZ_REQUIRE(g.n_valid % 32 == 0, PZ_ERR_ARG, "tc_rowgemm: n_valid must be a multiple of 32");
  if (g.Nout % 256 == 0) return rowgemm_launch<256, 4>(g, st);
  return rowgemm_launch<128, 4>(g, st);
}

}  // namespace pz
