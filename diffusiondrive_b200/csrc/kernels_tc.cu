// Tensor-core engine of the ddh planning head for sm_100a: tcgen05.mma with fp32 accumulators
// in TMEM, weights staged by TMA (128-byte swizzle), activations / gathered BEV patches staged
// by cp.async into the same swizzled K-major layout, mbarrier pipelines between the roles.
//
//   C[m, n0:n0+256] = A[m, :] . W[n0:n0+256, :]^T        bf16 x bf16 -> fp32
//
// One CTA = NT row tiles of 128 rows x one 256-wide column block, 192 threads:
//   warps 0-3  A producers during the main loop (cp.async gather, manual 128B swizzle),
//              then the epilogue warps (tcgen05.ld -> smem staging -> shared row epilogue)
//   warp  4    TMA producer for the weight tile (one elected lane)
//   warp  5    TMEM allocator + MMA issuer (one elected lane)
//
// CONV = true is the on-demand value_proj of GridSampleCrossBEVAttention
// (modules/blocks.py:68-76,114): row r of a scene is the 3x3xC patch around the r-th unique
// sampled pixel, gathered from the NHWC bf16 BEV map with zero padding; K = 9*C ordered
// (tap, channel) to match the packed weights.
#include "kernels.h"

namespace ddh {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                        // 64 bf16 = 128 B = one swizzle span
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;     // 16 KiB
constexpr int TC_B_TILE = D * TC_BK * 2;         // 32 KiB
constexpr int TC_CS_LD = D + 4;                  // padded fp32 staging row
constexpr int TC_LAG = 2;                        // cp.async groups kept in flight

template <int NT>
struct TcCfg {
  static constexpr int kStages = (NT == 1) ? 4 : 3;
  static constexpr int kStageBytes = NT * TC_A_TILE + TC_B_TILE;
  static constexpr int kPipeBytes = kStages * kStageBytes;
  static constexpr int kStagingBytes = 4 * 32 * TC_CS_LD * 4;
  static constexpr int kBarBytes = 256;
  static constexpr int kSmemBytes =
      (kPipeBytes > kStagingBytes ? kPipeBytes : kStagingBytes) + kBarBytes + 1024;
  static constexpr int kTmemCols = NT * D;       // 256 or 512 (power of two)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, single-CTA, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor: K-major operand, 128-byte swizzle, bf16.
// Rows are 128 B apart, 8-row groups 1024 B apart (SBO = 64 in 16-byte units); LBO is unused
// for swizzled K-major layouts (1); version = 1 (sm_100); layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=256
__device__ __forceinline__ uint32_t umma_idesc_bf16_m128_n256() {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------ the kernel
template <int NT, bool CONV>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap wmap) {
  using Cfg = TcCfg<NT>;
  constexpr int NS = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  constexpr int kPipeOrStage =
      Cfg::kPipeBytes > Cfg::kStagingBytes ? Cfg::kPipeBytes : Cfg::kStagingBytes;
  const uint32_t bar_addr = sm_addr + kPipeOrStage;  // full[NS], empty[NS], accum, tmem slot
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(sm + kPipeOrStage + (2 * NS + 1) * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- tile coordinates
  int n0 = 0, row0, rows_valid, scene = 0;
  long long out_row0;
  if (CONV) {
    scene = blockIdx.y;
    const int nu = p.nuniq[scene];
    row0 = blockIdx.x * (NT * TC_BM);
    if (row0 >= nu) return;
    rows_valid = min(NT * TC_BM, nu - row0);
    out_row0 = (long long)scene * p.rcap + row0;
  } else {
    row0 = blockIdx.x * (NT * TC_BM);
    rows_valid = min(NT * TC_BM, p.M - row0);
    n0 = blockIdx.y * D;
    out_row0 = row0;
  }
  const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
  const int KC = p.K / TC_BK;

  auto full_bar = [&](int s) { return bar_addr + s * 8; };
  auto empty_bar = [&](int s) { return bar_addr + (NS + s) * 8; };
  const uint32_t accum_bar = bar_addr + 2 * NS * 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 128 + 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<Cfg::kTmemCols>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 4 && lane == 0) tma_prefetch_desc(&wmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ======================= A producers ==========================================
    const int tid = threadIdx.x;
    const int j = tid & 7;        // 16-byte chunk of the 128-byte row
    const int rb = tid >> 3;      // 0..15
    constexpr int RPT = NT * 8;   // rows per thread
    int info[RPT];                // CONV: (y << 16) | x, or -1 ; dense: 1 / -1
#pragma unroll
    for (int i = 0; i < RPT; ++i) {
      const int r = rb + 16 * i;
      if (r < rows_valid) {
        if (CONV) {
          const int pix = p.upix[(long long)scene * p.rcap + row0 + r];
          const int y = pix / p.W_;
          info[i] = (y << 16) | (pix - y * p.W_);
        } else {
          info[i] = 1;
        }
      } else {
        info[i] = -1;
      }
    }
    const __nv_bfloat16* Ab = reinterpret_cast<const __nv_bfloat16*>(p.A);
    const __nv_bfloat16* bev = reinterpret_cast<const __nv_bfloat16*>(p.bev);
    const int cchunks = CONV ? (p.C / TC_BK) : 1;

    auto signal = [&](int kc_done) {
      fence_proxy_async();
      mbar_arrive(full_bar(kc_done % NS));
    };

    for (int kc = 0; kc < KC; ++kc) {
      const int s = kc % NS;
      mbar_wait(empty_bar(s), ((kc / NS) & 1) ^ 1);
      const uint32_t a_stage = sm_addr + s * Cfg::kStageBytes;
      int dy = 0, dx = 0, c0 = 0;
      if (CONV) {
        const int tap = kc / cchunks;
        c0 = (kc - tap * cchunks) * TC_BK;
        dy = tap / 3 - 1;
        dx = tap % 3 - 1;
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rb + 16 * i;
        const int t = r >> 7, rt = r & 127;
        if (t < nt_active) {
          const void* src = CONV ? (const void*)bev : (const void*)Ab;
          uint32_t nbytes = 0;
          if (info[i] >= 0) {
            if (CONV) {
              const int yy = (info[i] >> 16) + dy, xx = (info[i] & 0xffff) + dx;
              if (yy >= 0 && yy < p.H && xx >= 0 && xx < p.W_) {
                src = bev + (((long long)scene * p.H + yy) * p.W_ + xx) * p.C + c0 + j * 8;
                nbytes = 16;
              }
            } else {
              src = Ab + (long long)(row0 + r) * p.lda + kc * TC_BK + j * 8;
              nbytes = 16;
            }
          }
          const uint32_t dst = a_stage + t * TC_A_TILE + rt * 128 + ((j ^ (rt & 7)) << 4);
          cp_async16(dst, src, nbytes);
        }
      }
      cp_async_commit();
      if (kc >= TC_LAG) {
        cp_async_wait<TC_LAG>();
        signal(kc - TC_LAG);
      }
    }
    // drain the last TC_LAG groups in order
    cp_async_wait<0>();
    for (int kc = (KC > TC_LAG ? KC - TC_LAG : 0); kc < KC; ++kc) signal(kc);

    // ======================= epilogue ==============================================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    float* Cs = reinterpret_cast<float*>(sm) + warp * 32 * TC_CS_LD;
    for (int t = 0; t < nt_active; ++t) {
#pragma unroll 1
      for (int c0 = 0; c0 < D; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + t * D + c0, v);
        tmem_ld_wait();
        float* dst = Cs + lane * TC_CS_LD + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<float4*>(dst + q * 4) =
              make_float4(__uint_as_float(v[q * 4 + 0]), __uint_as_float(v[q * 4 + 1]),
                          __uint_as_float(v[q * 4 + 2]), __uint_as_float(v[q * 4 + 3]));
      }
      __syncwarp();
      for (int rr = 0; rr < 32; ++rr) {
        const int r = t * TC_BM + warp * 32 + rr;
        if (r < rows_valid) {
          float v[8];
          load8(Cs + rr * TC_CS_LD, lane, v);
          row_epilogue(p.epi, v, out_row0 + r, n0, lane);
        }
      }
      __syncwarp();
    }
  } else if (warp == 4) {
    // ======================= TMA producer (weights) ================================
    if (lane == 0) {
      for (int kc = 0; kc < KC; ++kc) {
        const int s = kc % NS;
        mbar_wait(empty_bar(s), ((kc / NS) & 1) ^ 1);
        mbar_arrive_expect_tx(full_bar(s), TC_B_TILE);
        tma_load_2d(sm_addr + s * Cfg::kStageBytes + NT * TC_A_TILE, &wmap, full_bar(s),
                    kc * TC_BK, n0);
      }
    }
    __syncwarp();
  } else {
    // ======================= MMA issuer ============================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128_n256();
      for (int kc = 0; kc < KC; ++kc) {
        const int s = kc % NS;
        mbar_wait(full_bar(s), (kc / NS) & 1);
        tc_fence_after();
        const uint32_t a_stage = sm_addr + s * Cfg::kStageBytes;
        const uint32_t b_stage = a_stage + NT * TC_A_TILE;
        for (int t = 0; t < nt_active; ++t) {
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
            const uint64_t adesc = umma_desc_sw128(a_stage + t * TC_A_TILE + k4 * 32);
            const uint64_t bdesc = umma_desc_sw128(b_stage + k4 * 32);
            umma_bf16(tmem_base + t * D, adesc, bdesc, idesc, (kc | k4) ? 1u : 0u);
          }
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

int tc_engine_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(tc_gemm_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           TcCfg<1>::kSmemBytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_gemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           TcCfg<2>::kSmemBytes);
  return (int)e;
}

void launch_tc_gemm(const GemmParams& p, const CUtensorMap& wmap, int n_total, cudaStream_t st) {
  dim3 grid((p.M + TC_BM - 1) / TC_BM, n_total / D);
  tc_gemm_kernel<1, false><<<grid, TC_THREADS, TcCfg<1>::kSmemBytes, st>>>(p, wmap);
}

void launch_tc_conv(const GemmParams& p, const CUtensorMap& wmap, int B, cudaStream_t st) {
  dim3 grid((p.rcap + 2 * TC_BM - 1) / (2 * TC_BM), B);
  tc_gemm_kernel<2, true><<<grid, TC_THREADS, TcCfg<2>::kSmemBytes, st>>>(p, wmap);
}

}  // namespace ddh
