// Post-forward epilogue and loss-side kernels:
//   * chamfer_loss              model5_b.py:1495-1505 (same code in dataset.py:1135-1145)
//   * comp(g, igt)              model5_b.py:1512-1519
//   * softmax -> top-128 boundary selection, gather, se3.transform, boundary chamfer, IoU counts and the
//     isotropic pose errors of test_step, model5_b.py:1314-1358 with metrics.py:54-84 -- fused into ONE launch
//     (one CTA per pair) by pz_pair_score; the reference issues ~60 tiny launches and two [B,128,128] bmm's.
// Everything here is latency-bound scalar work (a few hundred KB per batch); the design goal is launch count.
#include <math.h>

#include "pz_common.cuh"
#include "se3_math.cuh"

namespace pz {
namespace {

// a 3-term dot product the way a K=3 GEMM accumulates it (the reference gets these from torch.bmm)
__device__ __forceinline__ float dot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fmaf(az, bz, fmaf(ay, by, ax * bx));
}

// ------------------------------------------------------------------ chamfer_loss
// P[i,j] = (|x_i|^2 + |y_j|^2) - 2 x_i.y_j ; blockIdx.y = 0: min over i for every j, 1: min over j for every i.
constexpr int CH_TILE = 1024;
__global__ void __launch_bounds__(256) chamfer_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                      int n, int m, float* __restrict__ min_over_x,
                                                      float* __restrict__ min_over_y, int* __restrict__ arg_x,
                                                      int* __restrict__ arg_y) {
  __shared__ float4 ref[CH_TILE];
  const int b = blockIdx.z, dir = blockIdx.y;
  const int nq = dir == 0 ? m : n, nr = dir == 0 ? n : m;
  if (blockIdx.x * 256 >= nq) return;
  const float* Q = dir == 0 ? y + (size_t)b * m * 3 : x + (size_t)b * n * 3;
  const float* R = dir == 0 ? x + (size_t)b * n * 3 : y + (size_t)b * m * 3;
  const int q = blockIdx.x * 256 + threadIdx.x;
  float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
  if (q < nq) {
    qx = Q[q * 3], qy = Q[q * 3 + 1], qz = Q[q * 3 + 2];
    qn = dot3(qx, qy, qz, qx, qy, qz);
  }
  float best = INFINITY;
  int bi = 0;
  for (int r0 = 0; r0 < nr; r0 += CH_TILE) {
    const int cnt = min(CH_TILE, nr - r0);
    __syncthreads();
    for (int t = threadIdx.x; t < cnt; t += 256) {
      const float rx = R[(r0 + t) * 3], ry = R[(r0 + t) * 3 + 1], rz = R[(r0 + t) * 3 + 2];
      ref[t] = make_float4(rx, ry, rz, dot3(rx, ry, rz, rx, ry, rz));
    }
    __syncthreads();
    if (q < nq) {
#pragma unroll 4
      for (int t = 0; t < cnt; ++t) {
        const float4 r = ref[t];
        const float p = (r.w + qn) - 2.f * dot3(r.x, r.y, r.z, qx, qy, qz);
        if (p < best) { best = p; bi = r0 + t; }
      }
    }
  }
  if (q < nq) {
    float* o = dir == 0 ? min_over_x : min_over_y;
    int* a = dir == 0 ? arg_x : arg_y;
    o[(size_t)b * nq + q] = best;
    if (a) a[(size_t)b * nq + q] = bi;
  }
}

// d(sum_j w1_j min_i P_ij + sum_i w2_i min_j P_ij) / d(x, y) given the arg-mins of the forward.
// One CTA per batch item; fp32 atomics on that item's rows (summation order, hence the last bit, is not fixed).
__global__ void __launch_bounds__(256) chamfer_grad_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                           int n, int m, const int* __restrict__ arg_x,
                                                           const int* __restrict__ arg_y, const float* __restrict__ gx_w,
                                                           const float* __restrict__ gy_w, float* __restrict__ gx,
                                                           float* __restrict__ gy) {
  // gx_w [B,m] = dL/d(min_over_x[b,j]); gy_w [B,n] = dL/d(min_over_y[b,i])
  const int b = blockIdx.x;
  const float* X = x + (size_t)b * n * 3;
  const float* Y = y + (size_t)b * m * 3;
  float* GX = gx + (size_t)b * n * 3;
  float* GY = gy + (size_t)b * m * 3;
  for (int t = threadIdx.x; t < n * 3; t += 256) GX[t] = 0.f;
  for (int t = threadIdx.x; t < m * 3; t += 256) GY[t] = 0.f;
  __syncthreads();
  // term 1: for every j, i* = arg_x[j]:  dP/dx_i* = 2 x - 2 y_j ; dP/dy_j = 2 y_j - 2 x
  for (int j = threadIdx.x; j < m; j += 256) {
    const float w = gx_w[(size_t)b * m + j];
    const int i = arg_x[(size_t)b * m + j];
    for (int c = 0; c < 3; ++c) {
      const float d = 2.f * (X[i * 3 + c] - Y[j * 3 + c]) * w;
      atomicAdd(&GX[i * 3 + c], d);
      atomicAdd(&GY[j * 3 + c], -d);
    }
  }
  for (int i = threadIdx.x; i < n; i += 256) {
    const float w = gy_w[(size_t)b * n + i];
    const int j = arg_y[(size_t)b * n + i];
    for (int c = 0; c < 3; ++c) {
      const float d = 2.f * (X[i * 3 + c] - Y[j * 3 + c]) * w;
      atomicAdd(&GX[i * 3 + c], d);
      atomicAdd(&GY[j * 3 + c], -d);
    }
  }
}

// ------------------------------------------------------------------ comp(g, igt)
// mse(g . igt, I, reduction='mean') * 16  ==  sum_{b,r,c} (A - I)^2 / B ; one CTA, fixed summation order.
__global__ void __launch_bounds__(256) comp_kernel(const float* __restrict__ g, const float* __restrict__ igt, int B,
                                                   float* __restrict__ loss) {
  __shared__ float red[256];
  float acc = 0.f;
  for (int e = threadIdx.x; e < B * 16; e += 256) {
    const int b = e >> 4, r = (e >> 2) & 3, c = e & 3;
    const float* G = g + (size_t)b * 16 + r * 4;
    const float* H = igt + (size_t)b * 16 + c;
    float a = G[0] * H[0];
    a = fmaf(G[1], H[4], a);
    a = fmaf(G[2], H[8], a);
    a = fmaf(G[3], H[12], a);
    const float d = a - (r == c ? 1.f : 0.f);
    acc = fmaf(d, d, acc);
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] / (float)B;
}

// ------------------------------------------------------------------ se3.transform
__global__ void __launch_bounds__(256) se3_transform_kernel(const float* __restrict__ g, const float* __restrict__ pts,
                                                            int n, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* G = g + (size_t)b * 16;
  const float* p = pts + ((size_t)b * n + i) * 3;
  const float a0 = p[0], a1 = p[1], a2 = p[2];
  float* o = out + ((size_t)b * n + i) * 3;
  for (int r = 0; r < 3; ++r) o[r] = dot3(G[r * 4], G[r * 4 + 1], G[r * 4 + 2], a0, a1, a2) + G[r * 4 + 3];
}

// ------------------------------------------------------------------ softmax(dim=1)[:,1] -> top-K
// class-1 probability exactly as torch.softmax evaluates it over two classes
__device__ __forceinline__ float prob1(float l0, float l1) {
  const float mx = fmaxf(l0, l1);
  const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
  return e1 / (e0 + e1);
}

// keys[0..1024) -> sorted descending (bitonic, 1024 threads).  key = prob bits << 32 | ~index, so equal
// probabilities resolve to the LOWEST index (torch.topk leaves ties unspecified).
__device__ __forceinline__ void sort1024_desc(unsigned long long* keys) {
  const int tid = threadIdx.x;
  for (int k = 2; k <= 1024; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int ixj = tid ^ j;
      if (ixj > tid) {
        const unsigned long long a = keys[tid], c = keys[ixj];
        const bool desc = (tid & k) == 0;
        if ((a < c) == desc) { keys[tid] = c; keys[ixj] = a; }
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ unsigned long long make_key(float p, int idx) {
  return ((unsigned long long)__float_as_uint(p) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)idx);
}
__device__ __forceinline__ int key_index(unsigned long long k) { return (int)(0xFFFFFFFFu - (unsigned)(k & 0xFFFFFFFFull)); }

__global__ void __launch_bounds__(1024) boundary_topk_kernel(const float* __restrict__ logits, int N, int K,
                                                             int64_t* __restrict__ idx, float* __restrict__ prob) {
  __shared__ unsigned long long keys[1024];
  const int b = blockIdx.x, t = threadIdx.x;
  const float* L = logits + (size_t)b * 2 * N;
  keys[t] = t < N ? make_key(prob1(L[t], L[N + t]), t) : 0ull;
  __syncthreads();
  sort1024_desc(keys);
  if (t < K) {
    idx[(size_t)b * K + t] = key_index(keys[t]);
    if (prob) prob[(size_t)b * K + t] = __uint_as_float((unsigned)(keys[t] >> 32));
  }
}

// torch.topk(v, K, 1) for rows of at most 1024 floats of either sign (largest = 0 selects the K smallest, i.e.
// torch.topk(-v, K)): the float bits are mapped to an order-preserving unsigned key first.
__global__ void __launch_bounds__(1024) select_topk_kernel(const float* __restrict__ v, int N, int K, int largest,
                                                           int64_t* __restrict__ idx, float* __restrict__ vals) {
  __shared__ unsigned long long keys[1024];
  const int b = blockIdx.x, t = threadIdx.x;
  if (t < N) {
    const float x = largest ? v[(size_t)b * N + t] : -v[(size_t)b * N + t];
    unsigned u = __float_as_uint(x);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    keys[t] = ((unsigned long long)u << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)t);
  } else {
    keys[t] = 0ull;
  }
  __syncthreads();
  sort1024_desc(keys);
  if (t < K) {
    const int i = key_index(keys[t]);
    idx[(size_t)b * K + t] = i;
    if (vals) vals[(size_t)b * K + t] = v[(size_t)b * N + i];
  }
}

// ------------------------------------------------------------------ fused test_step epilogue
constexpr int NB = 128;    // boundary points per cloud (model5_b.py:1327, :1329)
constexpr int NP = 1024;   // points per cloud

struct PairScoreArgs {
  const float *out6, *de_fpcb, *de_mrpcb, *fpc, *src, *fpcb, *rpcb, *fpc_idx, *rpc_idx, *igt;
  float* scores;
  int64_t *idx_f, *idx_m;
  float *bnd_f, *bnd_m;
};

__device__ __forceinline__ float block_sum_1024(float v, float* red) {
  // fixed order: warp shuffle tree, then warp 0 over the 32 partials
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 32) {
    r = red[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  if (threadIdx.x == 0) red[32] = r;
  __syncthreads();
  return red[32];
}

__global__ void __launch_bounds__(1024) pair_score_kernel(const PairScoreArgs a) {
  __shared__ unsigned long long keys[1024];
  __shared__ float4 pts[4][NB];   // 0: de_fpcb points, 1: aligned de_mrpcb points, 2: fpcb, 3: rpcb  (w = |p|^2)
  __shared__ int sel[2][NB];
  __shared__ float mat[16];
  __shared__ float red[40];
  __shared__ float group_sum[6];
  const int b = blockIdx.x, t = threadIdx.x;
  float* S = a.scores + (size_t)b * PZ_SCORE_COLS;

  if (t == 0) {
    float tw[6];
    for (int i = 0; i < 6; ++i) tw[i] = a.out6[b * 6 + i];
    se3_exp_dev(tw, mat);
    float r_iso = 0.f, t_iso = 0.f, t_mse = 0.f, t_mae = 0.f;
    if (a.igt) {
      // compute_metrics (model5_b.py:1426-1440): gt is inverted first (metrics.py:7-10), then
      // isotropic_R_error(R, inv_R) (metrics.py:54-71) and isotropic_t_error(t, inv_t, inv_R) (:74-84)
      const float* G = a.igt + (size_t)b * 16;
      float gR[9], gt[3], it[3];
      for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) gR[r * 3 + c] = G[r * 4 + c];
        gt[r] = G[r * 4 + 3];
      }
      for (int r = 0; r < 3; ++r) it[r] = -(gR[0 * 3 + r] * gt[0] + gR[1 * 3 + r] * gt[1] + gR[2 * 3 + r] * gt[2]);
      // r1r2 = (inv_R)^T R = gR R ; only the trace is needed
      float tr = 0.f;
      for (int r = 0; r < 3; ++r) tr += gR[r * 3] * mat[0 * 4 + r] + gR[r * 3 + 1] * mat[1 * 4 + r] + gR[r * 3 + 2] * mat[2 * 4 + r];
      const float cs = fminf(fmaxf((tr - 1.f) / 2.f, -1.f), 1.f);
      r_iso = acosf(cs) / 3.14159265358979323846f * 180.f;
      // inv_R_t(inv_R, inv_t) = (gR, -gR inv_t); error = gR t + that
      float e2 = 0.f;
      for (int r = 0; r < 3; ++r) {
        const float rt = gR[r * 3] * mat[3] + gR[r * 3 + 1] * mat[7] + gR[r * 3 + 2] * mat[11];
        const float t2 = -(gR[r * 3] * it[0] + gR[r * 3 + 1] * it[1] + gR[r * 3 + 2] * it[2]);
        const float e = rt + t2;
        e2 += e * e;
        const float d = mat[r * 4 + 3] - it[r];
        t_mse += d * d;
        t_mae += fabsf(d);
      }
      t_iso = sqrtf(e2);
      t_mse /= 3.f;
      t_mae /= 3.f;
    }
    S[0] = r_iso; S[1] = t_iso; S[2] = t_mse; S[3] = t_mae;
  }

  // ---- boundary selection for both clouds
  for (int which = 0; which < 2; ++which) {
    const float* L = (which == 0 ? a.de_fpcb : a.de_mrpcb) + (size_t)b * 2 * NP;
    __syncthreads();
    keys[t] = make_key(prob1(L[t], L[NP + t]), t);
    __syncthreads();
    sort1024_desc(keys);
    if (t < NB) {
      const int i = key_index(keys[t]);
      sel[which][t] = i;
      int64_t* o = which == 0 ? a.idx_f : a.idx_m;
      if (o) o[(size_t)b * NB + t] = i;
    }
  }
  __syncthreads();

  // ---- gather (+ align) the predicted boundaries, stage the ground-truth ones
  {
    const int w = t >> 7, k = t & (NB - 1);
    if (w == 0) {
      const float* p = a.fpc + ((size_t)b * NP + sel[0][k]) * 3;
      pts[0][k] = make_float4(p[0], p[1], p[2], dot3(p[0], p[1], p[2], p[0], p[1], p[2]));
      if (a.bnd_f) for (int c = 0; c < 3; ++c) a.bnd_f[((size_t)b * NB + k) * 3 + c] = p[c];
    } else if (w == 1) {
      const float* p = a.src + ((size_t)b * NP + sel[1][k]) * 3;
      float q[3];
      for (int r = 0; r < 3; ++r) q[r] = dot3(mat[r * 4], mat[r * 4 + 1], mat[r * 4 + 2], p[0], p[1], p[2]) + mat[r * 4 + 3];
      pts[1][k] = make_float4(q[0], q[1], q[2], dot3(q[0], q[1], q[2], q[0], q[1], q[2]));
      if (a.bnd_m) for (int c = 0; c < 3; ++c) a.bnd_m[((size_t)b * NB + k) * 3 + c] = q[c];
    } else if (w == 2 || w == 3) {
      const float* base = w == 2 ? a.fpcb : a.rpcb;
      if (base) {
        const float* p = base + ((size_t)b * NB + k) * 3;
        pts[w][k] = make_float4(p[0], p[1], p[2], dot3(p[0], p[1], p[2], p[0], p[1], p[2]));
      } else {
        pts[w][k] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  __syncthreads();

  // ---- six 128x128 directed chamfer passes: (fpcb, de_fpcb), (rpcb, de_rpcb), (de_fpcb, aligned de_mrpcb)
  {
    const int w = t >> 7, k = t & (NB - 1);
    float best = 0.f;
    if (w < 6) {
      const int pairs_x[3] = {2, 3, 0}, pairs_y[3] = {0, 1, 1};
      const int xi = pairs_x[w >> 1], yi = pairs_y[w >> 1];
      const float4* Q = (w & 1) == 0 ? pts[yi] : pts[xi];   // even: min over x for every y_j
      const float4* R = (w & 1) == 0 ? pts[xi] : pts[yi];
      const float4 q = Q[k];
      best = INFINITY;
      for (int i = 0; i < NB; ++i) {
        const float4 r = R[i];
        best = fminf(best, (r.w + q.w) - 2.f * dot3(r.x, r.y, r.z, q.x, q.y, q.z));
      }
    }
    // per-group (128 threads = 4 warps) sums
    float v = best;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if ((t & 31) == 0) red[t >> 5] = v;
    __syncthreads();
    if (t < 6) group_sum[t] = ((red[t * 4] + red[t * 4 + 1]) + (red[t * 4 + 2] + red[t * 4 + 3])) / (float)NB;
    __syncthreads();
  }

  // ---- IoU counts (model5_b.py:1332-1341): |pred & gt|, |pred | gt| with |pred| = 128 distinct indices
  float inter_f = 0.f, inter_m = 0.f, gt_f = 0.f, gt_m = 0.f;
  if (a.fpc_idx) {
    gt_f = a.fpc_idx[(size_t)b * NP + t] != 0.f ? 1.f : 0.f;
    if (t < NB) inter_f = a.fpc_idx[(size_t)b * NP + sel[0][t]] != 0.f ? 1.f : 0.f;
  }
  if (a.rpc_idx) {
    gt_m = a.rpc_idx[(size_t)b * NP + t] != 0.f ? 1.f : 0.f;
    if (t < NB) inter_m = a.rpc_idx[(size_t)b * NP + sel[1][t]] != 0.f ? 1.f : 0.f;
  }
  const float If = block_sum_1024(inter_f, red), Gf = block_sum_1024(gt_f, red);
  const float Im = block_sum_1024(inter_m, red), Gm = block_sum_1024(gt_m, red);
  if (t == 0) {
    S[4] = If; S[5] = (float)NB + Gf - If;
    S[6] = Im; S[7] = (float)NB + Gm - Im;
    S[8] = a.fpcb ? group_sum[0] + group_sum[1] : 0.f;
    S[9] = a.rpcb ? group_sum[2] + group_sum[3] : 0.f;
    S[10] = group_sum[4] + group_sum[5];
    S[11] = 0.f;
  }
}

}  // namespace
}  // namespace pz

using namespace pz;

extern "C" int pz_chamfer(const float* x, const float* y, int B, int n, int m, float* min_over_x, float* min_over_y,
                          int32_t* arg_x_or_null, int32_t* arg_y_or_null, pz_stream_t stream) {
  PZ_REQUIRE(B >= 0 && n >= 0 && m >= 0, PZ_ERR_ARG, "pz_chamfer: negative size");
  if (B == 0 || (n == 0 && m == 0)) return PZ_OK;
  PZ_REQUIRE(n > 0 && m > 0, PZ_ERR_ARG, "pz_chamfer: one cloud is empty (min over an empty set)");
  PZ_REQUIRE(x && y && min_over_x && min_over_y, PZ_ERR_ARG, "pz_chamfer: null pointer");
  PZ_REQUIRE(B <= 65535, PZ_ERR_UNSUPPORTED, "pz_chamfer: B > 65535");
  dim3 grid((max(n, m) + 255) / 256, 2, B);
  chamfer_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, y, n, m, min_over_x, min_over_y, arg_x_or_null,
                                                      arg_y_or_null);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_chamfer_grad(const float* x, const float* y, int B, int n, int m, const int32_t* arg_x,
                               const int32_t* arg_y, const float* grad_min_over_x, const float* grad_min_over_y,
                               float* grad_x, float* grad_y, pz_stream_t stream) {
  PZ_REQUIRE(B >= 0 && n > 0 && m > 0, PZ_ERR_ARG, "pz_chamfer_grad: bad size");
  if (B == 0) return PZ_OK;
  PZ_REQUIRE(x && y && arg_x && arg_y && grad_min_over_x && grad_min_over_y && grad_x && grad_y, PZ_ERR_ARG,
             "pz_chamfer_grad: null pointer");
  chamfer_grad_kernel<<<B, 256, 0, as_stream(stream)>>>(x, y, n, m, arg_x, arg_y, grad_min_over_x, grad_min_over_y,
                                                        grad_x, grad_y);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_comp(const float* g, const float* igt, int B, float* loss, pz_stream_t stream) {
  PZ_REQUIRE(g && igt && loss, PZ_ERR_ARG, "pz_comp: null pointer");
  PZ_REQUIRE(B > 0, PZ_ERR_ARG, "pz_comp: B <= 0 (mean over an empty batch)");
  comp_kernel<<<1, 256, 0, as_stream(stream)>>>(g, igt, B, loss);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_se3_transform(const float* g, const float* pts, int B, int n, float* out, pz_stream_t stream) {
  PZ_REQUIRE(B >= 0 && n >= 0, PZ_ERR_ARG, "pz_se3_transform: negative size");
  if (B == 0 || n == 0) return PZ_OK;
  PZ_REQUIRE(g && pts && out, PZ_ERR_ARG, "pz_se3_transform: null pointer");
  PZ_REQUIRE(B <= 65535, PZ_ERR_UNSUPPORTED, "pz_se3_transform: B > 65535");
  se3_transform_kernel<<<dim3((n + 255) / 256, B), 256, 0, as_stream(stream)>>>(g, pts, n, out);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_boundary_topk(const float* logits, int B, int N, int K, int64_t* idx, float* prob_or_null,
                                pz_stream_t stream) {
  PZ_REQUIRE(B >= 0, PZ_ERR_ARG, "pz_boundary_topk: B < 0");
  if (B == 0) return PZ_OK;
  PZ_REQUIRE(logits && idx, PZ_ERR_ARG, "pz_boundary_topk: null pointer");
  PZ_REQUIRE(N >= 1 && N <= 1024, PZ_ERR_UNSUPPORTED, "pz_boundary_topk: N must be in [1,1024] (got %d)", N);
  PZ_REQUIRE(K >= 1 && K <= N, PZ_ERR_ARG, "pz_boundary_topk: K must be in [1,N]");
  boundary_topk_kernel<<<B, 1024, 0, as_stream(stream)>>>(logits, N, K, idx, prob_or_null);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_topk(const float* values, int B, int N, int K, int largest, int64_t* idx, float* vals_or_null,
                       pz_stream_t stream) {
  PZ_REQUIRE(B >= 0, PZ_ERR_ARG, "pz_topk: B < 0");
  if (B == 0) return PZ_OK;
  PZ_REQUIRE(values && idx, PZ_ERR_ARG, "pz_topk: null pointer");
  PZ_REQUIRE(N >= 1 && N <= 1024, PZ_ERR_UNSUPPORTED, "pz_topk: N must be in [1,1024] (got %d)", N);
  PZ_REQUIRE(K >= 1 && K <= N, PZ_ERR_ARG, "pz_topk: K must be in [1,N]");
  select_topk_kernel<<<B, 1024, 0, as_stream(stream)>>>(values, N, K, largest, idx, vals_or_null);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}

extern "C" int pz_pair_score(const float* out6, const float* de_fpcb, const float* de_mrpcb, const float* fpc,
                             const float* src, const float* fpcb_or_null, const float* rpcb_or_null,
                             const float* fpc_idx_or_null, const float* rpc_idx_or_null, const float* igt_or_null,
                             int B, float* scores, int64_t* idx_f_or_null, int64_t* idx_m_or_null,
                             float* bnd_f_or_null, float* bnd_m_or_null, pz_stream_t stream) {
  PZ_REQUIRE(B >= 0, PZ_ERR_ARG, "pz_pair_score: B < 0");
  if (B == 0) return PZ_OK;
  PZ_REQUIRE(out6 && de_fpcb && de_mrpcb && fpc && src && scores, PZ_ERR_ARG, "pz_pair_score: null pointer");
  PairScoreArgs a{out6, de_fpcb, de_mrpcb, fpc, src, fpcb_or_null, rpcb_or_null, fpc_idx_or_null, rpc_idx_or_null,
                  igt_or_null, scores, idx_f_or_null, idx_m_or_null, bnd_f_or_null, bnd_m_or_null};
  pair_score_kernel<<<B, 1024, 0, as_stream(stream)>>>(a);
  PZ_LAUNCH_CHECK();
  return PZ_OK;
}
