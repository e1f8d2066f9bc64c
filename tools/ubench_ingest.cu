// Micro-benchmark: how fast can the CTAs of one 16-CTA cluster pull L2-resident bytes into shared
// memory?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_ingest ubench_ingest.cu
// Modes: 0 TMA 2D box {64 x 128 rows} (128-byte swizzle), every CTA the same tiles
//        1 same, every CTA its own matrix
//        2 1D bulk copy of 16 KiB, every CTA the same bytes
//        3 1D bulk copy, every CTA its own bytes
//        4 1D bulk copy with cluster multicast (CTA r issues every 16th slot for all)
//        5 cp.async 16 B (256 threads), every CTA the same bytes
//        6 cp.async 16 B, every CTA its own bytes
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../diffusiondrive_b200/csrc/tc_ptx.cuh"
using namespace ddh;

constexpr int SLOT = 32768;
constexpr int NSLOT = 6;
constexpr int NT = 288;

__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t ncta() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t c) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(c)); return r; }
__device__ __forceinline__ void bulk_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_1d_mc(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "h"(mask) : "memory");
}
__device__ __forceinline__ void arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

__global__ void __launch_bounds__(NT, 1)
ingest_kernel(const CUtensorMap* maps, const uint8_t* flat, size_t per_cta_bytes, int mode, int nslots_total,
              int K, long long* out, int sz, int nprod) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;
  const uint32_t sm = raw + pad;
  const uint32_t bar = sm + NSLOT * SLOT;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rank = (int)ctarank(), nc = (int)ncta();
  auto full = [&](int s) { return bar + s * 8; };
  auto empty = [&](int s) { return bar + (NSLOT + s) * 8; };
  auto go = [&](int s) { return bar + (2 * NSLOT + s) * 8; };
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(full(s), mode >= 5 ? 256 : 1);
      mbar_init(empty(s), 1);
      mbar_init(go(s), nc);
    }
    fence_barrier_init();
  }
  __syncthreads();
  csync();
  const bool own = (mode == 1 || mode == 3 || mode == 6);
  const uint8_t* base = flat + (own ? (size_t)rank * per_cta_bytes : 0);
  const CUtensorMap* map = maps + (own ? rank : 0);
  const int kch = K / 64;
  long long t0 = clock64();
  if (mode <= 4) {
    if (warp >= 4 && warp < 4 + nprod && (tid & 31) == 0) {   // producers: warp 4 + p takes requests q = p (mod nprod)
      for (int q = warp - 4; q < nslots_total; q += nprod) {
        const int s = q % NSLOT, use = q / NSLOT;
        if (use > 0) mbar_wait(empty(s), (use - 1) & 1);
        mbar_arrive_expect_tx(full(s), sz);
        if (mode <= 1) {
          const int rows = sz / 128;
          const int tile = q % (per_cta_bytes / sz);
          tma_load_2d(sm + s * SLOT, map, full(s), (tile % kch) * 64, (tile / kch) * rows);
        } else if (mode <= 3) {
          bulk_1d(sm + s * SLOT, base + (size_t)(q % (per_cta_bytes / sz)) * sz, sz, full(s));
        } else {
          // tell the issuer of this slot that my copy of the slot is free and armed
          const int issuer = q % nc;
          arrive_remote(mapa(go(s), issuer));
          if (issuer == rank) {
            mbar_wait(go(s), (q / (nc > NSLOT ? nc : NSLOT)) & 1);
            bulk_1d_mc(sm + s * SLOT, base + (size_t)(q % (per_cta_bytes / SLOT)) * SLOT, SLOT, full(s),
                       (uint16_t)((1u << nc) - 1u));
          }
        }
      }
    } else if (warp == 0 && tid == 0) {   // consumer
      for (int q = 0; q < nslots_total; ++q) {
        const int s = q % NSLOT;
        mbar_wait(full(s), (q / NSLOT) & 1);
        mbar_arrive(empty(s));
      }
    }
  } else {
    if (warp < 8) {
      for (int q = 0; q < nslots_total; ++q) {
        const int s = q % NSLOT, use = q / NSLOT;
        if (use > 0) mbar_wait(empty(s), (use - 1) & 1);
        const uint8_t* src = base + (size_t)(q % (per_cta_bytes / SLOT)) * SLOT;
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(sm + s * SLOT + (tid + i * 256) * 16, src + (tid + i * 256) * 16, 16);
        cp_async_mbar_arrive_noinc(full(s));
      }
    } else if (tid == 256) {
      for (int q = 0; q < nslots_total; ++q) {
        const int s = q % NSLOT;
        mbar_wait(full(s), (q / NSLOT) & 1);
        mbar_arrive(empty(s));
      }
    }
  }
  __syncthreads();
  long long t1 = clock64();
  csync();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int K = 256, N = 4096;                       // per-CTA matrix: 4096 x 256 bf16 = 2 MiB
  const size_t per = (size_t)N * K * 2;
  uint8_t* flat;
  cudaMalloc(&flat, per * 16);
  cudaMemset(flat, 1, per * 16);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  EncodeFn enc = (EncodeFn)fn;
  CUtensorMap hm[16];
  for (int r = 0; r < 16; ++r) {
    cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
    cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, 128};
    cuuint32_t es[2] = {1, 1};
    CUresult rc = enc(&hm[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, flat + r * per, gdim, gstr, box, es,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc) { printf("encode failed %d\n", (int)rc); return 1; }
  }
  CUtensorMap* dm;
  cudaMalloc(&dm, sizeof hm);
  cudaMemcpy(dm, hm, sizeof hm, cudaMemcpyHostToDevice);
  long long* out;
  cudaMalloc(&out, 16 * 8);
  const int smem = NSLOT * SLOT + 1024 + 1024;
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(ingest_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cl : {1}) {
    for (int mode : {0, 2}) {
      for (int nprod : {1, 2, 4}) for (int sz : {8192, 32768}) {
        // per-size tensor maps (box rows = sz / 128)
        for (int r = 0; r < 16; ++r) {
          cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
          cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
          cuuint32_t box[2] = {64, (cuuint32_t)(sz / 128)};
          cuuint32_t es[2] = {1, 1};
          enc(&hm[r], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, flat + r * per, gdim, gstr, box, es,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        cudaMemcpy(dm, hm, sizeof hm, cudaMemcpyHostToDevice);
        const int total = (8 << 20) / sz;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cl);
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        long long h[16];
        for (int rep = 0; rep < 3; ++rep) {
          cudaError_t e = cudaLaunchKernelEx(&cfg, ingest_kernel, (const CUtensorMap*)dm, (const uint8_t*)flat, per, mode,
                                             total, K, out, sz, nprod);
          if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); return 1; }
          e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("sync failed: %s\n", cudaGetErrorString(e)); return 1; }
        }
        cudaMemcpy(h, out, sizeof(long long) * cl, cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int i = 0; i < cl; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("cluster %2d mode %d producers %d request %5d B: %8lld cycles for 8 MiB per CTA -> %.1f B/clk per CTA, %.0f cycles per request\n",
               cl, mode, nprod, sz, mx, (double)(8 << 20) / (double)mx, (double)mx / total);
      }
    }
  }
  return 0;
}
