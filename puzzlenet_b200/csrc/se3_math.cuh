// se3.exp as a device function (se_math/se3.py:57-80 with so3.mat :16-26 and the sinc1/2/3 Taylor branches of
// sinc.py:6-18, :96-108, :126-138), shared by pz_se3_exp (encoder.cu) and the fused pair-score epilogue (losses.cu).
#pragma once

namespace pz {

// twist x = (omega, v) -> g row-major 4x4:  R = I + sinc1 W + sinc2 W^2,  p = (I + sinc2 W + sinc3 W^2) v
__device__ __forceinline__ void se3_exp_dev(const float* x, float* g) {
  const float w0 = x[0], w1 = x[1], w2 = x[2];
  const float t = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
  const float t2 = t * t;
  float s1, s2, s3;
  if (fabsf(t) < 0.01f) {
    s1 = 1.f - t2 / 6.f * (1.f - t2 / 20.f * (1.f - t2 / 42.f));
    s2 = 0.5f * (1.f - t2 / 12.f * (1.f - t2 / 30.f * (1.f - t2 / 56.f)));
    s3 = (1.f / 6.f) * (1.f - t2 / 20.f * (1.f - t2 / 42.f * (1.f - t2 / 72.f)));
  } else {
    const float sn = sinf(t), cs = cosf(t);
    s1 = sn / t;
    s2 = (1.f - cs) / t2;
    s3 = (t - sn) / (t2 * t);
  }
  const float W[9] = {0.f, -w2, w1, w2, 0.f, -w0, -w1, w0, 0.f};
  float S[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) S[i * 3 + j] = W[i * 3] * W[j] + W[i * 3 + 1] * W[3 + j] + W[i * 3 + 2] * W[6 + j];
  for (int i = 0; i < 3; ++i) {
    float p = 0.f;
    for (int j = 0; j < 3; ++j) {
      const float id = i == j ? 1.f : 0.f;
      g[i * 4 + j] = id + s1 * W[i * 3 + j] + s2 * S[i * 3 + j];
      p += (id + s2 * W[i * 3 + j] + s3 * S[i * 3 + j]) * x[3 + j];
    }
    g[i * 4 + 3] = p;
  }
  g[12] = 0.f; g[13] = 0.f; g[14] = 0.f; g[15] = 1.f;
}

}  // namespace pz
