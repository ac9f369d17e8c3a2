// Approximate earth-mover distance (auction-style soft matching) for sm_100a.
// Reference semantics: PyTorchEMD/cuda/emd_kernel.cu:25-158 (approxmatch), :200-243 (matchcost),
// :286-355 (matchcostgrad1/2).  The reference runs 32 CTAs that each loop over batch items and
// read-modify-write the n*m `match` matrix once per annealing level (10 x 4 MB at 1024^2).
//
// B200 design: every batch item gets its own thread-block cluster (1, 2 or 4 CTAs, chosen so the
// grid fills the 148 SMs); the CTAs split the outer index of each pass and exchange the small
// remain/ratio vectors through global scratch + cluster barriers.  `match` is NOT touched inside
// the level loop: the per-level ratio vectors (10 x (n+m) floats) are kept, and match is written
// exactly once at the end as  sum_j exp(level_j d^2) ratioL_j[k] ratioR_j[l]  in the reference's
// level order, which is the same sequence of fp32 additions the reference performs on memory.
#include <cooperative_groups.h>

#include "pz_common.cuh"

namespace cg = cooperative_groups;

namespace pz {

constexpr int EMD_T = 512;
constexpr int EMD_TILE = 1024;  // inner points staged per smem tile (emd_kernel.cu:36 `Block`)
constexpr int EMD_LEVELS = 10;  // j = 7 .. -2 (emd_kernel.cu:46)

__device__ __forceinline__ float emd_level(int j) {  // emd_kernel.cu:47-50
  return j == -2 ? 0.f : -powf(4.0f, (float)j);
}

__device__ __forceinline__ float emd_d2(float x1, float y1, float z1, float x2, float y2, float z2) {
  return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1) + (z2 - z1) * (z2 - z1);
}

// scratch per item: remainL[n] remainR[m] ratioL[n] ratioR[m] histL[10][n] histR[10][m]
__global__ void __launch_bounds__(EMD_T) emd_approxmatch_kernel(int n, int m, const float* __restrict__ xyz1,
                                                                const float* __restrict__ xyz2,
                                                                float* __restrict__ match,
                                                                float* __restrict__ scratch, int cl) {
  __shared__ float4 buf[EMD_TILE];
  cg::cluster_group cluster = cg::this_cluster();
  const int item = blockIdx.x / cl;
  const int rank = blockIdx.x - item * cl;
  const int t = threadIdx.x;
  const float* p1 = xyz1 + (size_t)item * n * 3;
  const float* p2 = xyz2 + (size_t)item * m * 3;
  float* sc = scratch + (size_t)item * (size_t)(12 * (n + m));
  float* remainL = sc;
  float* remainR = remainL + n;
  float* ratioL = remainR + m;
  float* ratioR = ratioL + n;
  float* histL = ratioR + m;               // [10][n]
  float* histR = histL + (size_t)EMD_LEVELS * n;  // [10][m]
  float* mt = match + (size_t)item * n * m;

  float multiL, multiR;  // integer division as in emd_kernel.cu:29-35
  if (n >= m) { multiL = 1.f; multiR = (float)(n / m); } else { multiL = (float)(m / n); multiR = 1.f; }
  // this CTA's slice of the outer indices
  const int kchunk = (n + cl - 1) / cl, kbeg = rank * kchunk, kend = min(n, kbeg + kchunk);
  const int lchunk = (m + cl - 1) / cl, lbeg = rank * lchunk, lend = min(m, lbeg + lchunk);
  for (int k = kbeg + t; k < kend; k += EMD_T) remainL[k] = multiL;
  for (int l = lbeg + t; l < lend; l += EMD_T) remainR[l] = multiR;
  cluster.sync();

  for (int lv = 0; lv < EMD_LEVELS; ++lv) {
    const float level = emd_level(7 - lv);
    // ---- pass 1: ratioL[k] = remainL[k] / (1e-9 + sum_l e * remainR[l])        (:51-84)
    for (int k0 = kbeg; k0 < kend; k0 += EMD_T) {
      const int k = k0 + t;
      float x1 = 0, y1 = 0, z1 = 0;
      if (k < kend) { x1 = p1[k * 3]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; }
      float suml = 1e-9f;
      for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int cnt = min(m, l0 + EMD_TILE) - l0;
        __syncthreads();
        for (int l = t; l < cnt; l += EMD_T)
          buf[l] = make_float4(p2[(l0 + l) * 3], p2[(l0 + l) * 3 + 1], p2[(l0 + l) * 3 + 2], remainR[l0 + l]);
        __syncthreads();
#pragma unroll 4
        for (int l = 0; l < cnt; ++l) {
          const float4 q = buf[l];
          suml += __expf(level * emd_d2(x1, y1, z1, q.x, q.y, q.z)) * q.w;
        }
      }
      if (k < kend) ratioL[k] = remainL[k] / suml;
    }
    cluster.sync();
    // ---- pass 2: consumption on the right side                                      (:86-119)
    for (int l0 = lbeg; l0 < lend; l0 += EMD_T) {
      const int l = l0 + t;
      float x2 = 0, y2 = 0, z2 = 0;
      if (l < lend) { x2 = p2[l * 3]; y2 = p2[l * 3 + 1]; z2 = p2[l * 3 + 2]; }
      float sumr = 0.f;
      for (int k0 = 0; k0 < n; k0 += EMD_TILE) {
        const int cnt = min(n, k0 + EMD_TILE) - k0;
        __syncthreads();
        for (int k = t; k < cnt; k += EMD_T)
          buf[k] = make_float4(p1[(k0 + k) * 3], p1[(k0 + k) * 3 + 1], p1[(k0 + k) * 3 + 2], ratioL[k0 + k]);
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
          const float4 q = buf[k];
          sumr += __expf(level * emd_d2(q.x, q.y, q.z, x2, y2, z2)) * q.w;
        }
      }
      if (l < lend) {
        const float rr = remainR[l];
        sumr *= rr;
        const float consumption = fminf(rr / (sumr + 1e-9f), 1.0f);
        const float v = consumption * rr;
        ratioR[l] = v;
        histR[(size_t)lv * m + l] = v;
        remainR[l] = fmaxf(0.0f, rr - sumr);
      }
    }
    cluster.sync();
    // ---- pass 3: remainL[k] -= sum_l e * ratioL[k] * ratioR[l]                     (:121-154)
    for (int k0 = kbeg; k0 < kend; k0 += EMD_T) {
      const int k = k0 + t;
      float x1 = 0, y1 = 0, z1 = 0, rl = 0;
      if (k < kend) { x1 = p1[k * 3]; y1 = p1[k * 3 + 1]; z1 = p1[k * 3 + 2]; rl = ratioL[k]; }
      float suml = 0.f;
      for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
        const int cnt = min(m, l0 + EMD_TILE) - l0;
        __syncthreads();
        for (int l = t; l < cnt; l += EMD_T)
          buf[l] = make_float4(p2[(l0 + l) * 3], p2[(l0 + l) * 3 + 1], p2[(l0 + l) * 3 + 2], ratioR[l0 + l]);
        __syncthreads();
#pragma unroll 4
        for (int l = 0; l < cnt; ++l) {
          const float4 q = buf[l];
          suml += __expf(level * emd_d2(x1, y1, z1, q.x, q.y, q.z)) * rl * q.w;
        }
      }
      if (k < kend) {
        histL[(size_t)lv * n + k] = rl;
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
    cluster.sync();
  }

  // ---- match[l][k] = sum over levels of e * ratioL_j[k] * ratioR_j[l], written once, k contiguous
  for (int l0 = lbeg; l0 < lend; l0 += 8) {
    const int lcnt = min(8, lend - l0);
    __syncthreads();
    if (t < lcnt * EMD_LEVELS) {
      const int li = t / EMD_LEVELS, lv = t - li * EMD_LEVELS;
      reinterpret_cast<float*>(buf)[64 + li * EMD_LEVELS + lv] = histR[(size_t)lv * m + l0 + li];
    }
    if (t < lcnt * 3) reinterpret_cast<float*>(buf)[t] = p2[(size_t)l0 * 3 + t];
    __syncthreads();
    const float* sb = reinterpret_cast<const float*>(buf);
    for (int k = t; k < n; k += EMD_T) {
      const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
      float hl[EMD_LEVELS];
#pragma unroll
      for (int lv = 0; lv < EMD_LEVELS; ++lv) hl[lv] = histL[(size_t)lv * n + k];
      for (int li = 0; li < lcnt; ++li) {
        const float d2 = emd_d2(x1, y1, z1, sb[li * 3], sb[li * 3 + 1], sb[li * 3 + 2]);
        float acc = 0.f;
#pragma unroll
        for (int lv = 0; lv < EMD_LEVELS; ++lv)
          acc += __expf(emd_level(7 - lv) * d2) * hl[lv] * sb[64 + li * EMD_LEVELS + lv];
        mt[(size_t)(l0 + li) * n + k] = acc;
      }
    }
  }
}

// cost[i] = sum_{k,l} d2(k,l) * match[i,l,k]     (emd_kernel.cu:200-243).  One CTA per item
// (deterministic summation order, no scratch); thread -> k so match rows are read coalesced.
__global__ void __launch_bounds__(1024) emd_matchcost_kernel(int n, int m, const float* __restrict__ xyz1,
                                                             const float* __restrict__ xyz2,
                                                             const float* __restrict__ match,
                                                             float* __restrict__ cost) {
  __shared__ float red[32];
  __shared__ float q[EMD_TILE * 3];
  const int item = blockIdx.x, t = threadIdx.x;
  const float* p1 = xyz1 + (size_t)item * n * 3;
  const float* p2 = xyz2 + (size_t)item * m * 3;
  const float* mt = match + (size_t)item * n * m;
  float sub = 0.f;
  for (int l0 = 0; l0 < m; l0 += EMD_TILE) {
    const int cnt = min(m, l0 + EMD_TILE) - l0;
    __syncthreads();
    for (int i = t; i < cnt * 3; i += 1024) q[i] = p2[(size_t)l0 * 3 + i];
    __syncthreads();
    for (int k = t; k < n; k += 1024) {
      const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
#pragma unroll 16   // 16 independent 4 KB row reads in flight per CTA pass: the loop is bound by HBM latency x concurrency
      for (int l = 0; l < cnt; ++l)
        sub += emd_d2(x1, y1, z1, q[l * 3], q[l * 3 + 1], q[l * 3 + 2]) * mt[(size_t)(l0 + l) * n + k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sub += __shfl_xor_sync(0xffffffffu, sub, o);
  if ((t & 31) == 0) red[t >> 5] = sub;
  __syncthreads();
  if (t < 32) {
    float s = red[t];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (t == 0) cost[item] = s;
  }
}

// grad1[i,k,:] = gc[i] * sum_l 2 match[i,l,k] (x1_k - x2_l)    (emd_kernel.cu:333-355)
__global__ void __launch_bounds__(128) emd_grad1_kernel(int n, int m, const float* __restrict__ gc,
                                                        const float* __restrict__ xyz1,
                                                        const float* __restrict__ xyz2,
                                                        const float* __restrict__ match, float* __restrict__ grad1) {
  const int item = blockIdx.y;
  const int k = blockIdx.x * 128 + threadIdx.x;
  if (k >= n) return;
  const float* p1 = xyz1 + (size_t)item * n * 3;
  const float* p2 = xyz2 + (size_t)item * m * 3;
  const float* mt = match + (size_t)item * n * m;
  const float x1 = p1[k * 3], y1 = p1[k * 3 + 1], z1 = p1[k * 3 + 2];
  float dx = 0, dy = 0, dz = 0;
#pragma unroll 16
  for (int l = 0; l < m; ++l) {
    const float d = mt[(size_t)l * n + k] * 2;
    dx += (x1 - p2[l * 3]) * d;
    dy += (y1 - p2[l * 3 + 1]) * d;
    dz += (z1 - p2[l * 3 + 2]) * d;
  }
  const float g = gc[item];
  float* o = grad1 + ((size_t)item * n + k) * 3;
  o[0] = dx * g; o[1] = dy * g; o[2] = dz * g;
}

// grad2[i,l,:] = gc[i] * sum_k 2 match[i,l,k] (x2_l - x1_k)    (emd_kernel.cu:286-327); warp per l
__global__ void __launch_bounds__(256) emd_grad2_kernel(int n, int m, const float* __restrict__ gc,
                                                        const float* __restrict__ xyz1,
                                                        const float* __restrict__ xyz2,
                                                        const float* __restrict__ match, float* __restrict__ grad2) {
  const int item = blockIdx.y;
  const int l = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (l >= m) return;
  const float* p1 = xyz1 + (size_t)item * n * 3;
  const float* p2 = xyz2 + ((size_t)item * m + l) * 3;
  const float* mr = match + (size_t)item * n * m + (size_t)l * n;
  const float x2 = p2[0], y2 = p2[1], z2 = p2[2];
  float sx = 0, sy = 0, sz = 0;
#pragma unroll 8
  for (int k = lane; k < n; k += 32) {
    const float d = mr[k] * 2;
    sx += (x2 - p1[k * 3]) * d;
    sy += (y2 - p1[k * 3 + 1]) * d;
    sz += (z2 - p1[k * 3 + 2]) * d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) {
    const float g = gc[item];
    float* o = grad2 + ((size_t)item * m + l) * 3;
    o[0] = sx * g; o[1] = sy * g; o[2] = sz * g;
  }
}

}  // namespace pz

using namespace pz;

static size_t emd_scratch_floats(int b, int n, int m) {
  return (size_t)b * 12 * ((size_t)n + m);
}

extern "C" size_t pz_emd_workspace_bytes(int b, int n, int m) {
  if (b < 1 || n < 1 || m < 1) return 0;
  return align_up(emd_scratch_floats(b, n, m) * sizeof(float), 256);
}

extern "C" int pz_emd_approxmatch(const float* xyz1, const float* xyz2, int b, int n, int m, float* match,
                                  void* workspace, size_t workspace_bytes, pz_stream_t stream) {
  PZ_REQUIRE(xyz1 && xyz2 && match, PZ_ERR_ARG, "pz_emd_approxmatch: null pointer");
  PZ_REQUIRE(b >= 0 && n >= 1 && m >= 1, PZ_ERR_ARG, "pz_emd_approxmatch: bad sizes b=%d n=%d m=%d", b, n, m);
  if (b == 0) return 0;
  PZ_REQUIRE(workspace && workspace_bytes >= pz_emd_workspace_bytes(b, n, m), PZ_ERR_WORKSPACE,
             "pz_emd_approxmatch: workspace %zu B < required %zu B", workspace_bytes, pz_emd_workspace_bytes(b, n, m));
  // cluster size: enough CTAs to cover the SMs, but no more than one CTA per 128 outer points
  int cl = 1;
  while (cl < 4 && b * cl * 2 <= kNumSMs + 20 && (n > 256 * cl || m > 256 * cl)) cl *= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(b * cl);
  cfg.blockDim = dim3(EMD_T);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PZ_CUDA(cudaLaunchKernelEx(&cfg, emd_approxmatch_kernel, n, m, xyz1, xyz2, match, static_cast<float*>(workspace), cl));
  count_launch();
  return 0;
}

extern "C" int pz_emd_matchcost(const float* xyz1, const float* xyz2, const float* match, int b, int n, int m,
                                float* cost, pz_stream_t stream) {
  PZ_REQUIRE(xyz1 && xyz2 && match && cost, PZ_ERR_ARG, "pz_emd_matchcost: null pointer");
  PZ_REQUIRE(b >= 0 && n >= 1 && m >= 1, PZ_ERR_ARG, "pz_emd_matchcost: bad sizes");
  if (b == 0) return 0;
  emd_matchcost_kernel<<<b, 1024, 0, as_stream(stream)>>>(n, m, xyz1, xyz2, match, cost);
  PZ_LAUNCH_CHECK();
  return 0;
}

extern "C" int pz_emd_matchcost_grad(const float* grad_cost, const float* xyz1, const float* xyz2,
                                     const float* match, int b, int n, int m, float* grad1, float* grad2,
                                     pz_stream_t stream) {
  PZ_REQUIRE(grad_cost && xyz1 && xyz2 && match && grad1 && grad2, PZ_ERR_ARG, "pz_emd_matchcost_grad: null pointer");
  PZ_REQUIRE(b >= 0 && n >= 1 && m >= 1, PZ_ERR_ARG, "pz_emd_matchcost_grad: bad sizes");
  if (b == 0) return 0;
  PZ_REQUIRE(b <= 65535, PZ_ERR_UNSUPPORTED, "pz_emd_matchcost_grad: b > 65535");
  emd_grad1_kernel<<<dim3((n + 127) / 128, b), 128, 0, as_stream(stream)>>>(n, m, grad_cost, xyz1, xyz2, match, grad1);
  PZ_LAUNCH_CHECK();
  emd_grad2_kernel<<<dim3((m + 7) / 8, b), 256, 0, as_stream(stream)>>>(n, m, grad_cost, xyz1, xyz2, match, grad2);
  PZ_LAUNCH_CHECK();
  return 0;
}
