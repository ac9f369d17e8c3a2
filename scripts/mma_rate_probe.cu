// tcgen05.mma issue / execution rate probe (B200): one CTA per SM issues `iters` back-to-back kind::f16 MMAs of M = 128,
// K = 16 from one thread on fixed shared-memory operands (SWIZZLE_128B K-major tiles, contents irrelevant) and reports the
// SM clocks per MMA for N = 64 / 128 / 256, both operands in shared memory (SS) or A in tensor memory (TS), optionally with
// 8 other warps hammering shared memory (what the epilogue warps of the attention kernel do during the P v phase).
//     nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I puzzlenet_b200/csrc -o /tmp/mma_probe scripts/mma_rate_probe.cu && /tmp/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace pz::tc;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// mode: 0 SS, 1 TS; noise: 1 = warps 1..8 stream st.shared / ld.shared over a 64 KB region while the MMAs run
__global__ void __launch_bounds__(288, 1) probe(int n, int mode, int noise, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  const uint32_t a_s = base, b_s = base + 16384, noise_s = base + 65536, bar = base + 65536 + 65536, slot = bar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  __shared__ volatile int stop;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    stop = 0;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (uint32_t i = tid; i < 65536 / 4; i += 288) reinterpret_cast<uint32_t*>(smem + (base - smem_u32(smem)))[i] = 0x3c003c00u;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + (slot - smem_u32(smem)));
  if (warp == 0 && mode >= 2) {
    // the issuing lane chosen by elect.sync inside a convergent warp: ptxas then knows that exactly one lane runs the MMAs
    uint32_t leader;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    long long t0 = clock64();
    if (leader) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a = make_desc(a_s), b = make_desc(b_s);
      // the shape of the kernels' loops: per stage 4 k-steps x 3 products, descriptors recomputed per stage
      for (int i = 0; i < iters / 12; ++i) {
        const uint64_t a_hi = a + (uint64_t)((i & 1) * 512), b_hi = b + (uint64_t)((i & 1) * 512), a_lo = a_hi + 256, b_lo = b_hi + 256;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          mma_ss(tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, (i | k4) != 0);
          mma_ss(tmem, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1);
          mma_ss(tmem, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1);
        }
      }
      commit(bar);
      mbar_wait(bar, 0);
    }
    __syncwarp();
    long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
    if (lane == 0) stop = 1;
  } else if (warp == 0 && mode == -1) {
    if (lane == 0) {   // the same loop under `lane == 0`
      const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a = make_desc(a_s), b = make_desc(b_s);
      long long t0 = clock64();
      for (int i = 0; i < iters / 12; ++i) {
        const uint64_t a_hi = a + (uint64_t)((i & 1) * 512), b_hi = b + (uint64_t)((i & 1) * 512), a_lo = a_hi + 256, b_lo = b_hi + 256;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          mma_ss(tmem, a_hi + 2 * k4, b_hi + 2 * k4, idesc, (i | k4) != 0);
          mma_ss(tmem, a_hi + 2 * k4, b_lo + 2 * k4, idesc, 1);
          mma_ss(tmem, a_lo + 2 * k4, b_hi + 2 * k4, idesc, 1);
        }
      }
      commit(bar);
      mbar_wait(bar, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
  } else if (warp == 0) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a = make_desc(a_s), b = make_desc(b_s);
      long long t0 = clock64();
      if (mode == 0) {
        mma_ss(tmem, a, b, idesc, 0);
        for (int i = 0; i < iters; i += 8) {   // unrolled: the issuing thread's own instructions must not be the limit
#pragma unroll
          for (int u = 0; u < 8; ++u) mma_ss(tmem, a + 2 * (u & 3), b + 2 * (u & 3), idesc, 1);
        }
      } else {
        mma_ts(tmem, tmem + 384, b, idesc, 0);
        for (int i = 0; i < iters; i += 8) {
#pragma unroll
          for (int u = 0; u < 8; ++u) mma_ts(tmem, tmem + 384 + 8 * (u & 3), b + 2 * (u & 3), idesc, 1);
        }
      }
      commit(bar);
      mbar_wait(bar, 0);
      long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      stop = 1;
    }
  } else if (noise == 2) {
    // what the epilogue warps of the kernels do while the MMAs run: spin on an mbarrier (here: the one the final commit completes)
    if (warp <= 4) mbar_wait(bar, 0);
  } else if (noise) {
    uint32_t addr = noise_s + (uint32_t)(warp - 1) * 8192 + lane * 16;
    uint32_t x = tid;
    while (!stop) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr + u * 512), "r"(x) : "memory");
      }
      uint32_t y0, y1, y2, y3;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(y0), "=r"(y1), "=r"(y2), "=r"(y3) : "r"(addr));
      x += y0 + y1 + y2 + y3;
    }
    if (x == 0x12345u) out[1] = x;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const size_t smem = 1024 + 65536 + 65536 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int iters = 4096;
  for (int mode : {-1, 2})
    for (int n : {128, 256}) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        probe<<<148, 288, smem>>>(n, mode, 0, 4092, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      }
      printf("kernel-shaped loop (12 MMAs per stage), issuing lane by %s, N=%3d: %.1f clocks per MMA\n", mode == 2 ? "elect.sync" : "lane == 0", n, (double)h / 4092);
    }
  for (int n : {128, 256}) {
    long long h = 0;
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<148, 288, smem>>>(n, 0, 2, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    }
    printf("SS N=%3d with 4 warps spinning on mbarrier.try_wait: %.1f clocks per MMA\n", n, (double)h / iters);
  }
  for (int noise = 0; noise < 2; ++noise)
    for (int mode = 0; mode < 2; ++mode)
      for (int n : {64, 128, 256}) {
        long long h = 0;
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<148, 288, smem>>>(n, mode, noise, iters, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        }
        printf("%s N=%3d %s: %.1f clocks per MMA (floor 128 N / 256 = %d)\n", mode ? "TS" : "SS", n, noise ? "with shared-memory traffic from 8 warps" : "quiet", (double)h / iters, n / 2);
      }
  return 0;
}
