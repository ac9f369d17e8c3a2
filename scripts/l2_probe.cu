// L2 -> SM read bandwidth probe (B200): every SM streams a buffer that fits the 126 MB L2 again and again with 16-byte
// loads (ld.global.nc.L1::no_allocate, so that L1 cannot serve repeats) and with cp.async.cg into shared memory; the same
// kernels over a 4 GB buffer give the HBM figure.  Build + run on the GPU box:
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/l2_probe scripts/l2_probe.cu && /tmp/l2_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) read_ldg(const uint4* __restrict__ buf, size_t n16, int reps, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * 4) {
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t j = i + u * stride;
        if (j < n16)
          asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                       : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(buf + j));
        else v[u] = make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  if (acc == 0x12345678u) *sink = acc;
}

__global__ void __launch_bounds__(512) read_cpasync(const uint4* __restrict__ buf, size_t n16, int reps, unsigned* sink) {
  extern __shared__ uint4 sm[];   // 4 stages x 512 threads x 4 x 16 B = 128 KB
  const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sm);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (int r = 0; r < reps; ++r) {
    int st = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride * 4, st = (st + 1) & 3) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const size_t j = i + u * stride;
        if (j < n16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + ((st * 4 + u) * 512 + threadIdx.x) * 16), "l"(buf + j) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  acc = sm[threadIdx.x].x;
  if (acc == 0x12345678u) *sink = acc;
}

// random ROW gather: each group of 8 lanes reads one 128-byte piece of a random 512-byte row (the access pattern of the
// gathered GEMM's producers), UNROLL independent pieces in flight per thread
template <int UNROLL>
__global__ void __launch_bounds__(512) gather_rows(const uint4* __restrict__ buf, const int* __restrict__ rows, size_t nidx,
                                                   int reps, unsigned* sink) {
  unsigned acc = 0;
  const size_t gstride = ((size_t)gridDim.x * blockDim.x) >> 3;   // row slots per sweep
  const size_t slot0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const int piece = threadIdx.x & 7;
  for (int r = 0; r < reps; ++r)
    for (size_t i = slot0; i + (UNROLL - 1) * gstride < nidx; i += gstride * UNROLL) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int row = rows[i + u * gstride];
        const uint4* src = buf + (size_t)row * 32 + (r & 3) * 8 + piece;   // 512-byte rows = 32 pieces; quarter r & 3
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(src));
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
  if (acc == 0x12345678u) *sink = acc;
}

template <int UNROLL>
static void run_gather(const uint4* buf, const int* rows, size_t nidx, unsigned* sink, int blocks) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 8;
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0);
    gather_rows<UNROLL><<<blocks, 512>>>(buf, rows, nidx, reps, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
  }
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  printf("random 128-byte row pieces from a 64 MB buffer, %d in flight per thread, %d x 512 threads: %.1f GB/s\n", UNROLL, blocks,
         (double)nidx * 128 * reps / ms / 1e6);
}

int main() {
  unsigned* sink;
  cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(read_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  for (size_t mb : {16, 32, 64, 96, 4096}) {
    const size_t bytes = mb << 20, n16 = bytes / 16;
    uint4* buf;
    if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc %zu MB failed\n", mb); continue; }
    cudaMemset(buf, 1, bytes);
    const int reps = mb >= 4096 ? 2 : 40;
    for (int which = 0; which < 2; ++which) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int it = 0; it < 2; ++it) {   // first pass warms L2
        cudaEventRecord(e0);
        if (which == 0) read_ldg<<<148 * 2, 512>>>(buf, n16, reps, sink);
        else read_cpasync<<<148, 512, 128 * 1024>>>(buf, n16, reps, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%5zu MB buffer, %s: %.1f GB/s (%s)\n", mb, which == 0 ? "ld.global.nc 16 B    " : "cp.async.cg -> smem ",
             (double)bytes * reps / ms / 1e6, mb <= 96 ? "L2-resident" : "HBM");
    }
    cudaFree(buf);
  }
  {
    const size_t bytes = 64ull << 20, nrows = bytes / 512, nidx = 4ull << 20;
    uint4* buf; int* rows;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 1, bytes);
    cudaMalloc(&rows, nidx * sizeof(int));
    int* h = (int*)malloc(nidx * sizeof(int));
    unsigned long long st = 88172645463325252ull;
    for (size_t i = 0; i < nidx; ++i) { st ^= st << 13; st ^= st >> 7; st ^= st << 17; h[i] = (int)(st % nrows); }
    cudaMemcpy(rows, h, nidx * sizeof(int), cudaMemcpyHostToDevice);
    run_gather<1>(buf, rows, nidx, sink, 148 * 2);
    run_gather<4>(buf, rows, nidx, sink, 148 * 2);
    run_gather<8>(buf, rows, nidx, sink, 148 * 2);
    run_gather<8>(buf, rows, nidx, sink, 148 * 4);
    run_gather<4>(buf, rows, nidx, sink, 148);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
