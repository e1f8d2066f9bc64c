// Probe: where do the rows of a cta_group::1, M = 64 tcgen05.mma accumulator land in TMEM?
// A[r][0] = r + 1, B[n][0] = n + 1  =>  D[r][n] = (r + 1) * (n + 1).
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
__global__ void probe(float* out, int M) {
  extern __shared__ uint8_t raw[];
  const uint32_t a0 = smem_u32(raw);
  const uint32_t pad = ((a0 + 1023u) & ~1023u) - a0;
  uint8_t* sm = raw + pad;
  const uint32_t sa = a0 + pad;
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0;
  __syncthreads();
  // A tile at 0: 128 rows x 128 B; B tile at 16384: 16 rows x 128 B
  if (tid < 128) *reinterpret_cast<__nv_bfloat16*>(sm + tid * 128 + (((0) ^ (tid & 7)) << 4)) = __float2bfloat16((float)(tid + 1));
  if (tid < 16) *reinterpret_cast<__nv_bfloat16*>(sm + 16384 + tid * 128 + (((0) ^ (tid & 7)) << 4)) = __float2bfloat16((float)(tid + 1));
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<32>(smem_u32(&slot));
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // clear accumulator lanes first (M = 128 with zero-ish: write zeros via tcgen05.st)
  {
    uint32_t z[32];
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(-1.0f);
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16), z);
    tmem_st_wait();
    tc_fence_before();
  }
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    umma_bf16(tmem, umma_desc_sw128(sa), umma_desc_sw128(sa + 16384), idesc(M, 16), 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  uint32_t u[32];
  tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), u);
  tmem_ld_wait();
  for (int c = 0; c < 16; ++c) out[(warp * 32 + lane) * 16 + c] = __uint_as_float(u[c]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<32>(tmem);
}
int main() {
  float* d;
  cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  for (int M : {128, 64}) {
    probe<<<1, 128, 40960>>>(d, M);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M=%d failed: %s\n", M, cudaGetErrorString(e)); return 1; }
    float h[128 * 16];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("M=%d: lane -> D row (from column 0; -1 = untouched), col1/col0 ratio\n", M);
    for (int l = 0; l < 128; ++l) printf("%s%d:%g(%g)", (l % 8) ? " " : "\n  ", l, h[l * 16], h[l * 16] != 0 ? h[l * 16 + 1] / h[l * 16] : 0.f);
    printf("\n");
  }
  return 0;
}
