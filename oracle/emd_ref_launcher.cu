// TEST INFRASTRUCTURE ONLY -- thin launcher around the REFERENCE's own EMD kernels.
//
// The reference's PyTorchEMD/cuda/emd_kernel.cu does not compile against torch 2.11 (THC headers,
// SURVEY.md D13), but its four __global__ kernels are self-contained.  oracle/Makefile extracts the
// kernel line ranges (:25-158, :200-243, :286-355) from the file where it lies under /root/reference
// into a temporary include outside the repo and compiles them here, unmodified, with the reference's
// launch shapes (<<<32,512>>>, <<<dim3(32,32),256>>>; emd_kernel.cu:188,274,392-393).  The result,
// oracle/_ref/libemd_ref.so, is the "reference kernel on the same GPU" used by tests (parity) and by
// bench_emd (kernel to beat).  No reference source is stored in this repository.
#include <cuda_runtime.h>
#include <stdint.h>

#include EMD_REF_KERNELS_INC

extern "C" int emd_ref_approxmatch(int b, int n, int m, const float* xyz1, const float* xyz2, float* match,
                                   float* temp /* [32,(n+m)*2] */, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(match, 0, sizeof(float) * (size_t)b * n * m, st);
  cudaMemsetAsync(temp, 0, sizeof(float) * (size_t)32 * (n + m) * 2, st);
  approxmatch<float><<<32, 512, 0, st>>>(b, n, m, xyz1, xyz2, match, temp);
  return (int)cudaGetLastError();
}
extern "C" int emd_ref_matchcost(int b, int n, int m, const float* xyz1, const float* xyz2, const float* match,
                                 float* cost, void* stream) {
  matchcost<float><<<32, 512, 0, (cudaStream_t)stream>>>(b, n, m, xyz1, xyz2, match, cost);
  return (int)cudaGetLastError();
}
extern "C" int emd_ref_matchcost_grad(int b, int n, int m, const float* gc, const float* xyz1, const float* xyz2,
                                      const float* match, float* grad1, float* grad2, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  matchcostgrad1<float><<<32, 512, 0, st>>>(b, n, m, gc, xyz1, xyz2, match, grad1);
  matchcostgrad2<float><<<dim3(32, 32), 256, 0, st>>>(b, n, m, gc, xyz1, xyz2, match, grad2);
  return (int)cudaGetLastError();
}
