/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's approximate-EMD kernels.
 *
 * Follows PyTorchEMD/cuda/emd_kernel.cu of the reference: approxmatch :25-158, matchcost :200-243,
 * matchcostgrad2 :286-327, matchcostgrad1 :333-355.  Sequential fp32 C, one batch item at a time.
 * The GPU kernels use __expf (ex2.approx); this port uses expf, so parity with a GPU
 * implementation is to ~1e-4 relative on the cost, not bit-exact.
 *
 * Pinned by the only golden vector the reference holds for this path: the 2-point case in
 * PyTorchEMD/test_emd_loss.py:8-33 (per-item cost 0.71) -- see tests/test_oracle_emd.py.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static float sq3(const float* a, const float* b) {
  float dx = b[0] - a[0], dy = b[1] - a[1], dz = b[2] - a[2];
  return dx * dx + dy * dy + dz * dz;
}

/* xyz1 [b,n,3], xyz2 [b,m,3] -> match [b,m,n] (match[i][l][k] pairs xyz2 point l with xyz1 point k) */
void emd_oracle_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match) {
  float* remainL = (float*)malloc(sizeof(float) * (size_t)n);
  float* remainR = (float*)malloc(sizeof(float) * (size_t)m);
  float* ratioL = (float*)malloc(sizeof(float) * (size_t)n);
  float* ratioR = (float*)malloc(sizeof(float) * (size_t)m);
  /* integer division on purpose, emd_kernel.cu:29-35 */
  float multiL = n >= m ? 1.0f : (float)(m / n);
  float multiR = n >= m ? (float)(n / m) : 1.0f;
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    float* mt = match + (size_t)i * n * m;
    memset(mt, 0, sizeof(float) * (size_t)n * m);
    for (int k = 0; k < n; ++k) remainL[k] = multiL;
    for (int l = 0; l < m; ++l) remainR[l] = multiR;
    for (int j = 7; j >= -2; --j) {                       /* annealing levels, :46-50 */
      float level = j == -2 ? 0.0f : -powf(4.0f, (float)j);
      for (int k = 0; k < n; ++k) {                       /* :51-84 */
        float suml = 1e-9f;
        for (int l = 0; l < m; ++l) suml += expf(level * sq3(p1 + k * 3, p2 + l * 3)) * remainR[l];
        ratioL[k] = remainL[k] / suml;
      }
      for (int l = 0; l < m; ++l) {                       /* :86-119 */
        float sumr = 0.0f;
        for (int k = 0; k < n; ++k) sumr += expf(level * sq3(p1 + k * 3, p2 + l * 3)) * ratioL[k];
        sumr *= remainR[l];
        float consumption = fminf(remainR[l] / (sumr + 1e-9f), 1.0f);
        ratioR[l] = consumption * remainR[l];
        remainR[l] = fmaxf(0.0f, remainR[l] - sumr);
      }
      for (int k = 0; k < n; ++k) {                       /* :121-154 */
        float suml = 0.0f;
        for (int l = 0; l < m; ++l) {
          float w = expf(level * sq3(p1 + k * 3, p2 + l * 3)) * ratioL[k] * ratioR[l];
          mt[(size_t)l * n + k] += w;
          suml += w;
        }
        remainL[k] = fmaxf(0.0f, remainL[k] - suml);
      }
    }
  }
  free(remainL); free(remainR); free(ratioL); free(ratioR);
}

/* cost[i] = sum_{k,l} |x1_k - x2_l|^2 * match[i][l][k]   (:200-243) */
void emd_oracle_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match, float* cost) {
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    const float* mt = match + (size_t)i * n * m;
    double s = 0.0; /* the GPU kernel sums fp32 partials in a tree; double here is the neutral choice */
    for (int k = 0; k < n; ++k)
      for (int l = 0; l < m; ++l) s += (double)(sq3(p1 + k * 3, p2 + l * 3) * mt[(size_t)l * n + k]);
    cost[i] = (float)s;
  }
}

/* grad1[i][k] = gc[i] * sum_l 2 match[i][l][k] (x1_k - x2_l)   (:333-355)
 * grad2[i][l] = gc[i] * sum_k 2 match[i][l][k] (x2_l - x1_k)   (:286-327) */
void emd_oracle_matchcost_grad(int b, int n, int m, const float* gc, const float* xyz1, const float* xyz2,
                               const float* match, float* grad1, float* grad2) {
  for (int i = 0; i < b; ++i) {
    const float* p1 = xyz1 + (size_t)i * n * 3;
    const float* p2 = xyz2 + (size_t)i * m * 3;
    const float* mt = match + (size_t)i * n * m;
    for (int k = 0; k < n; ++k) {
      float d[3] = {0, 0, 0};
      for (int l = 0; l < m; ++l) {
        float w = mt[(size_t)l * n + k] * 2;
        for (int c = 0; c < 3; ++c) d[c] += (p1[k * 3 + c] - p2[l * 3 + c]) * w;
      }
      for (int c = 0; c < 3; ++c) grad1[((size_t)i * n + k) * 3 + c] = d[c] * gc[i];
    }
    for (int l = 0; l < m; ++l) {
      float d[3] = {0, 0, 0};
      for (int k = 0; k < n; ++k) {
        float w = mt[(size_t)l * n + k] * 2;
        for (int c = 0; c < 3; ++c) d[c] += (p2[l * 3 + c] - p1[k * 3 + c]) * w;
      }
      for (int c = 0; c < 3; ++c) grad2[((size_t)i * m + l) * 3 + c] = d[c] * gc[i];
    }
  }
}
