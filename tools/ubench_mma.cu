// How long does a tcgen05.mma (kind::f16, K = 16) occupy the tensor pipe as a function of its shape?
// 64 back-to-back instructions from one thread on resident shared-memory operands; dependent (same
// accumulator) and independent (4 accumulators round-robin) variants.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../diffusiondrive_b200/csrc/tc_ptx.cuh"
using namespace ddh;

__device__ __forceinline__ uint32_t idesc(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__global__ void rate(long long* out, int M, int N, int nacc, int count, int nthr) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  const uint32_t pad = ((a0 + 1023u) & ~1023u) - a0;
  const uint32_t sa = a0 + pad;
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + pad)[i] = 0;
  if (tid == 0) { mbar_init(smem_u32(&bar), nthr); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(smem_u32(&slot));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  __shared__ long long tstart;
  if (tid == 0) tstart = clock64();
  __syncthreads();
  if ((tid & 31) == 0 && warp < nthr) {
    const uint32_t id = idesc(M, N);
    const uint64_t ad = umma_desc_sw128(sa), bd = umma_desc_sw128(sa + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < count / nthr; ++i)
      umma_bf16(tmem + warp * (512 / 4) + (i % nacc) * (128 / nacc), ad + (uint64_t)((i & 3) * 2), bd + (uint64_t)((i & 3) * 2), id, 1u);
    const long long t1 = clock64();
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t2 = clock64();
    if (warp == 0) { out[0] = t1 - t0; }
    atomicMax((unsigned long long*)&out[1], (unsigned long long)(t2 - tstart));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}
int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
  const int shapes[][2] = {{64, 16}, {64, 64}, {128, 64}, {128, 128}};
  for (auto& sh : shapes)
    for (int nthr : {1, 2, 4}) {
      const int nacc = 1;
      for (int count : {64}) {
        long long h[2];
        for (int rep = 0; rep < 2; ++rep) {
          cudaMemset(d, 0, 16);
          rate<<<1, 128, 70000>>>(d, sh[0], sh[1], nacc, count, nthr);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("M=%3d N=%3d issuing threads=%d count=%2d: thread-0 issue %6lld cycles, all done %6lld cycles -> %.1f cycles per MMA\n", sh[0],
               sh[1], nthr, count, h[0], h[1], (double)h[1] / count);
      }
    }
  return 0;
}
