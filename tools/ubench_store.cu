// Micro-benchmark: how fast can one CTA per SM push bytes from the SM to global memory (L2 / HBM)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_store ubench_store.cu
// Modes: 0  st.global.v4 from registers, a warp writes 512 contiguous bytes per instruction (8 warps)
//        1  the same with 16 warps
//        2  st.global.b16, a warp writes 64 contiguous bytes per instruction (8 warps)
//        3  cp.async.bulk shared -> global, 8 KiB per request, one issuing thread, <= 4 groups in flight
//        4  cp.async.bulk shared -> global, 32 KiB per request, one issuing thread
//        5  cp.async.bulk shared -> global, 8 KiB per request, 4 issuing threads
// Every CTA writes its own region of `per_cta` bytes `reps` times (the region is L2-sized or larger when
// summed over the grid: 148 x 1 MiB).  Prints bytes / clock / SM from clock64 of CTA 0 and GB/s from events.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(512, 1)
store_kernel(uint8_t* out, size_t per_cta, int reps, int mode, long long* clk) {
  extern __shared__ __align__(128) uint8_t sm[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* base = out + (size_t)blockIdx.x * per_cta;
  for (int i = tid; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i * 2654435761u;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const long long t0 = clock64();
  if (mode <= 1) {
    const int nw = mode == 0 ? 8 : 16;
    if (warp < nw) {
      const uint4 v = make_uint4(tid, warp, lane, 7u);
      for (int r = 0; r < reps; ++r)
        for (size_t off = (size_t)warp * 512; off < per_cta; off += (size_t)nw * 512)
          *reinterpret_cast<uint4*>(base + off + lane * 16) = v;
    }
  } else if (mode == 2) {
    if (warp < 8) {
      const unsigned short v = (unsigned short)tid;
      for (int r = 0; r < reps; ++r)
        for (size_t off = (size_t)warp * 64; off < per_cta; off += 8 * 64)
          *reinterpret_cast<unsigned short*>(base + off + lane * 2) = v;
    }
  } else {
    const uint32_t req = mode == 4 ? 32768u : 8192u;
    const int nthr = mode == 5 ? 4 : 1;
    if (lane == 0 && warp < nthr) {
      for (int r = 0; r < reps; ++r)
        for (size_t off = (size_t)warp * req; off < per_cta; off += (size_t)nthr * req) {
          bulk_store(base + off, smem_u32(sm) + (uint32_t)(off % 32768u / req * req % 32768u), req);
          bulk_commit();
          bulk_wait_read<3>();
        }
      bulk_wait_read<0>();
    }
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) clk[0] = clock64() - t0;
}

int main(int argc, char** argv) {
  const size_t per_cta = 1 << 20;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint8_t* out;
  long long* clk;
  cudaMalloc(&out, per_cta * sms);
  cudaMalloc(&clk, 8);
  cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const char* names[] = {"st.v4 x 8 warps", "st.v4 x 16 warps", "st.b16 x 8 warps", "bulk 8 KiB x 1 thread",
                         "bulk 32 KiB x 1 thread", "bulk 8 KiB x 4 threads"};
  for (int grid : {1, sms}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int reps = mode == 2 ? 2 : 8;
      store_kernel<<<grid, 512, 32768>>>(out, per_cta, 1, mode, clk);
      cudaEventRecord(a);
      store_kernel<<<grid, 512, 32768>>>(out, per_cta, reps, mode, clk);
      cudaEventRecord(b);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, a, b);
      long long c = 0;
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      const double bytes = (double)per_cta * reps;
      printf("grid %3d  %-24s %7.1f B/clk/SM (CTA 0)   %8.1f GB/s total\n", grid, names[mode], bytes / (double)c,
             bytes * grid / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
