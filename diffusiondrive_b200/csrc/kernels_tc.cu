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
#include <type_traits>
#include "kernels.h"
#include "tc_ptx.cuh"

namespace ddh {

constexpr int TC_THREADS = 192;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                        // 64 bf16 = 128 B = one swizzle span
constexpr int TC_A_TILE = TC_BM * TC_BK * 2;     // 16 KiB
constexpr int TC_B_TILE = D * TC_BK * 2;         // 32 KiB
constexpr int TC_BAR_BYTES = 256;

// dense GEMM: persistent CTAs (one per SM) walking 128x256 tiles, 4 stages, two TMEM
// accumulator buffers so the epilogue of tile i overlaps the main loop of tile i+1;
// conv: 2 row tiles share each weight tile, 3 stages, 1 CTA per SM.
constexpr int G_NS = 4;
constexpr int G_STAGE = TC_A_TILE + TC_B_TILE;                 // 48 KiB
constexpr int G_THREADS = 320;                                 // 4 epilogue + 4 producer + TMA + MMA warps
constexpr int G_STG_BYTES = 4 * 32 * 36 * 4;                   // per-warp 32x32 transpose staging
// per-column epilogue vectors staged in smem (L1 is only ~12 KB next to 216 KB of smem, so
// "uniform" global loads of bias / LN / FiLM vectors would miss and serialise L2 round trips)
constexpr int G_PAR_BIAS = 0;        // up to 1024 floats (N_total <= 1024)
constexpr int G_PAR_LN1G = 1024, G_PAR_LN1B = 1280, G_PAR_LN2G = 1536, G_PAR_LN2B = 1792;
constexpr int G_PAR_FILM = 2048;     // 512
constexpr int G_PAR_DOTW = 2560;     // 256
constexpr int G_PAR_FLOATS = 2816;
constexpr int G_SMEM = G_NS * G_STAGE + G_STG_BYTES + G_PAR_FLOATS * 4 + TC_BAR_BYTES + 1024;
constexpr int C_NT = 2;
constexpr int C_NS = 3;
constexpr int C_STAGE = C_NT * TC_A_TILE + TC_B_TILE;          // 64 KiB
constexpr int C_PIPE = C_NS * C_STAGE;                         // 192 KiB
constexpr int C_VS_LD = 128 + 4;                               // fp32 staging row (half of N)
static_assert(C_NT * TC_BM * C_VS_LD * 4 <= C_PIPE, "combine staging must fit in the pipeline");

#ifdef DDH_TIMELINE
#define TL_STAMP(role, idx, which)                                                        \
  do {                                                                                    \
    if (p.dbg && blockIdx.x == 0 && (idx) < 40) p.dbg[((role)*40 + (idx)) * 2 + (which)] = clock64(); \
  } while (0)
#else
#define TL_STAMP(role, idx, which) do { } while (0)
#endif

// ------------------------------------------------------------------ shared roles
struct TcBars {
  uint32_t base;
  int ns;
  __device__ __forceinline__ uint32_t full(int s) const { return base + s * 8; }
  __device__ __forceinline__ uint32_t empty(int s) const { return base + (ns + s) * 8; }
  __device__ __forceinline__ uint32_t accum() const { return base + 2 * ns * 8; }
  __device__ __forceinline__ uint32_t passgo() const { return base + (2 * ns + 1) * 8; }
};

// One k-chunk of MMAs: NT_ACTIVE row tiles x (64/16) instructions, then release the stage.
template <int NT, bool TF32 = false>
__device__ __forceinline__ void mma_chunk(uint32_t a_stage, uint32_t b_stage, uint32_t tmem_base,
                                          int nt_active, bool first, uint32_t idesc,
                                          uint32_t empty_bar) {
  for (int t = 0; t < nt_active; ++t) {
#pragma unroll
    for (int k4 = 0; k4 < TC_BK / 16; ++k4) {   // 32 bytes of K per instruction: 16 bf16 or 8 tf32
      const uint64_t adesc = umma_desc_sw128(a_stage + t * TC_A_TILE + k4 * 32);
      const uint64_t bdesc = umma_desc_sw128(b_stage + k4 * 32);
      if (TF32) umma_tf32(tmem_base + t * D, adesc, bdesc, idesc, (first && k4 == 0) ? 0u : 1u);
      else umma_bf16(tmem_base + t * D, adesc, bdesc, idesc, (first && k4 == 0) ? 0u : 1u);
    }
  }
  umma_commit(empty_bar);
}

// ===================================================================================
// Dense rows GEMM + fused row epilogue (the Linear layers of the decoder chain).
// Epilogue: thread r of the 128 epilogue threads owns output row r (TMEM lane r) and walks
// its 256 columns in 8 blocks of 32 straight out of TMEM; LayerNorm statistics are
// thread-local sums, intermediate rows are written back to TMEM (tcgen05.st) between passes,
// nothing is staged in shared memory.
// ===================================================================================
__device__ __forceinline__ void ldg_row32(const float* p, float (&o)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
    o[4 * q + 0] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
  }
}

__device__ __forceinline__ void lds_row32(const float* p, float (&o)[32]) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const float4 t = *(reinterpret_cast<const float4*>(p) + q);
    o[4 * q + 0] = t.x; o[4 * q + 1] = t.y; o[4 * q + 2] = t.z; o[4 * q + 3] = t.w;
  }
}

__global__ void __launch_bounds__(G_THREADS, 1)
tc_gemm_kernel(const GemmParams p, const __grid_constant__ CUtensorMap wmap) {
  constexpr int NS = G_NS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  float* par = reinterpret_cast<float*>(sm + NS * G_STAGE + G_STG_BYTES);
  constexpr int kBarOff = NS * G_STAGE + G_STG_BYTES + G_PAR_FLOATS * 4;
  const uint32_t bar = sm_addr + kBarOff;
  auto full_bar = [&](int s) { return bar + s * 8; };
  auto empty_bar = [&](int s) { return bar + (NS + s) * 8; };
  auto tfull_bar = [&](int b) { return bar + (2 * NS + b) * 8; };
  auto tempty_bar = [&](int b) { return bar + (2 * NS + 2 + b) * 8; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + kBarOff + (2 * NS + 4) * 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_row_tiles = (p.M + TC_BM - 1) / TC_BM;
  const int n_tiles = n_row_tiles * p.n_blocks;
  const int KC = p.K / TC_BK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(full_bar(s), 128 + 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<2 * D>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 8 && lane == 0) tma_prefetch_desc(&wmap);
  {
    const RowEpi& e = p.epi;
    const int nb_tot = p.n_blocks * D;
    for (int i = threadIdx.x; i < nb_tot; i += G_THREADS) par[G_PAR_BIAS + i] = e.bias ? __ldg(e.bias + i) : 0.f;
    for (int i = threadIdx.x; i < D; i += G_THREADS) {
      if (e.ln1_g) { par[G_PAR_LN1G + i] = __ldg(e.ln1_g + i); par[G_PAR_LN1B + i] = __ldg(e.ln1_b + i); }
      if (e.ln2_g) { par[G_PAR_LN2G + i] = __ldg(e.ln2_g + i); par[G_PAR_LN2B + i] = __ldg(e.ln2_b + i); }
      if (e.film) { par[G_PAR_FILM + i] = __ldg(e.film + i); par[G_PAR_FILM + D + i] = __ldg(e.film + D + i); }
      if (e.dot_w) par[G_PAR_DOTW + i] = __ldg(e.dot_w + i);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ======================= epilogue warps (TMEM lane quarter = warp) =================
    // Arithmetic is thread-per-row (row = TMEM lane); global traffic goes through a 32x32
    // per-warp transpose in shared memory so that every global instruction covers 8 rows x
    // 64 contiguous bytes; residual / row-vector loads of block b+1 are in flight while
    // block b is processed.
    const RowEpi& e = p.epi;
    constexpr int SLD = 36;
    float* stg = reinterpret_cast<float*>(sm + NS * G_STAGE) + warp * (32 * SLD);
    uint32_t* stg16 = reinterpret_cast<uint32_t*>(stg);
    const int cr = lane & 7, cc = lane >> 3;
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int row0 = (tile / p.n_blocks) * TC_BM;
      const int n0 = (tile % p.n_blocks) * D;
      const int rows_valid = min(TC_BM, p.M - row0);
      const int buf = it & 1;
      const int r = warp * 32 + lane;
      const bool valid = r < rows_valid;
      const long long m = row0 + r;
      const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * D;

      auto issue_rows = [&](auto rowptr, int c0, float4 (&pre)[8]) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = cr + 8 * i;
          const bool ok = warp * 32 + rr < rows_valid;
          const float* g = rowptr((long long)row0 + warp * 32 + rr) + c0;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            pre[i * 2 + hh] = ok ? __ldg(reinterpret_cast<const float4*>(g) + cc + 4 * hh)
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      auto commit_rows = [&](const float4 (&pre)[8], float (&t)[32]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)
            *reinterpret_cast<float4*>(stg + (cr + 8 * i) * SLD + (cc + 4 * hh) * 4) = pre[i * 2 + hh];
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 v4 = *reinterpret_cast<const float4*>(stg + lane * SLD + 4 * q);
          t[4 * q + 0] = v4.x; t[4 * q + 1] = v4.y; t[4 * q + 2] = v4.z; t[4 * q + 3] = v4.w;
        }
        __syncwarp();
      };

      float dsum = 0.f;
      auto finalize = [&](int b, float (&v)[32]) {
        const int c0 = n0 + b * 32;
        if (e.film) {
          float sc[32], sh[32];
          lds_row32(par + G_PAR_FILM + c0, sc);
          lds_row32(par + G_PAR_FILM + D + c0, sh);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] * (1.0f + sc[i]) + sh[i];
        }
        if (e.dot_w) {
          float w[32];
          lds_row32(par + G_PAR_DOTW + c0, w);
#pragma unroll
          for (int i = 0; i < 32; ++i) dsum = fmaf(v[i], w[i], dsum);
        }
        if (e.out_f32) {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * SLD + 4 * q) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = cr + 8 * i;
            if (warp * 32 + rr < rows_valid) {
              float* g = e.out_f32 + ((long long)row0 + warp * 32 + rr) * e.ldo32 + c0;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int ch = cc + 4 * hh;
                reinterpret_cast<float4*>(g)[ch] =
                    *reinterpret_cast<const float4*>(stg + rr * SLD + ch * 4);
              }
            }
          }
          __syncwarp();
        }
        if (e.out_bf16) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * q + 0], v[8 * q + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * q + 2], v[8 * q + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * q + 4], v[8 * q + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * q + 6], v[8 * q + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(stg16 + lane * 20 + 4 * q) = u;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = cr + 8 * i;
            if (warp * 32 + rr < rows_valid) {
              __nv_bfloat16* g = e.out_bf16 + ((long long)row0 + warp * 32 + rr) * e.ldo16 + c0;
              reinterpret_cast<uint4*>(g)[cc] =
                  *reinterpret_cast<const uint4*>(stg16 + rr * 20 + cc * 4);
            }
          }
          __syncwarp();
        }
      };
      auto ld_block = [&](int b, float (&v)[32]) {
        uint32_t u[32];
        tmem_ld32(trow + b * 32, u);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
      };
      auto st_block = [&](int b, const float (&v)[32]) {
        uint32_t u[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(v[i]);
        tmem_st32(trow + b * 32, u);
      };
      const float* rbase = e.res;
      const long long ldr = e.ldres;
      auto res_ptr = [&](long long mm) { return rbase + mm * ldr; };
      const float* vbase = e.rowvec;
      const int rpg = e.rows_per_group;
      auto vec_ptr = [&](long long mm) { return vbase + (mm / rpg) * D; };

      if (threadIdx.x == 0) TL_STAMP(0, it, 0);
      mbar_wait(tfull_bar(buf), (it >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 0) TL_STAMP(1, it, 0);

      // pass 1: bias, ReLU, residual (+ statistics of LN1)
      float s1 = 0.f, q1 = 0.f, c1 = 0.f;
      float4 pre[8];
      if (e.res) issue_rows(res_ptr, n0, pre);
#pragma unroll 1
      for (int b = 0; b < 8; ++b) {
        float v[32];
        if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 8 + b * 4, 0);
        ld_block(b, v);
        if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 8 + b * 4, 1);
        if (e.bias) {
          float t[32];
          lds_row32(par + G_PAR_BIAS + n0 + b * 32, t);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
        }
        if (e.relu) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (e.res) {
          float t[32];
          if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 9 + b * 4, 0);
          commit_rows(pre, t);
          if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 9 + b * 4, 1);
          if (b < 7) issue_rows(res_ptr, n0 + (b + 1) * 32, pre);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += t[i];
        }
        if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 10 + b * 4, 0);
        if (e.ln1_g) {
          if (b == 0) c1 = v[0];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float d = v[i] - c1;
            s1 += d;
            q1 = fmaf(d, d, q1);
          }
          st_block(b, v);
        } else {
          finalize(b, v);
        }
        if (threadIdx.x == 0 && it == 1) TL_STAMP(3, 10 + b * 4, 1);
      }
      if (e.ln1_g) {
        tmem_st_wait();
        const float ms = s1 * (1.0f / D);
        const float mean = c1 + ms;
        const float rstd = 1.0f / sqrtf(fmaxf(q1 * (1.0f / D) - ms * ms, 0.f) + LN_EPS);
        float s2 = 0.f, q2 = 0.f, c2 = 0.f;
        if (e.rowvec) issue_rows(vec_ptr, 0, pre);
#pragma unroll 1
        for (int b = 0; b < 8; ++b) {
          float v[32], g[32], bb[32];
          ld_block(b, v);
          lds_row32(par + G_PAR_LN1G + b * 32, g);
          lds_row32(par + G_PAR_LN1B + b * 32, bb);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd * g[i] + bb[i];
          if (e.rowvec) {
            float t[32];
            commit_rows(pre, t);
            if (b < 7) issue_rows(vec_ptr, (b + 1) * 32, pre);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += t[i];
          }
          if (e.ln2_g) {
            if (b == 0) c2 = v[0];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float d = v[i] - c2;
              s2 += d;
              q2 = fmaf(d, d, q2);
            }
            st_block(b, v);
          } else {
            finalize(b, v);
          }
        }
        if (e.ln2_g) {
          tmem_st_wait();
          const float ms2 = s2 * (1.0f / D);
          const float mean2 = c2 + ms2;
          const float rstd2 = 1.0f / sqrtf(fmaxf(q2 * (1.0f / D) - ms2 * ms2, 0.f) + LN_EPS);
#pragma unroll 1
          for (int b = 0; b < 8; ++b) {
            float v[32], g[32], bb[32];
            ld_block(b, v);
            lds_row32(par + G_PAR_LN2G + b * 32, g);
            lds_row32(par + G_PAR_LN2B + b * 32, bb);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean2) * rstd2 * g[i] + bb[i];
            finalize(b, v);
          }
        }
      }
      if (e.dot_w && valid) e.dot_out[m] = dsum + e.dot_b[0];
      // this accumulator buffer may be overwritten by the MMA warp
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));
      if (threadIdx.x == 0) TL_STAMP(1, it, 1);
    }
  } else if (warp < 8) {
    // ======================= A producers ===============================================
    const int ptid = threadIdx.x - 128;
    const int j = ptid & 7, rb = ptid >> 3;
    const __nv_bfloat16* Ab = reinterpret_cast<const __nv_bfloat16*>(p.A);
    const uint32_t dst_base = rb * 128 + ((j ^ (rb & 7)) << 4);
    int g = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int row0 = (tile / p.n_blocks) * TC_BM;
      const int rows_valid = min(TC_BM, p.M - row0);
      const __nv_bfloat16* src[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = rb + 16 * i;
        src[i] = (r < rows_valid) ? (Ab + (long long)(row0 + r) * p.lda + j * 8) : nullptr;
      }
      for (int kc = 0; kc < KC; ++kc, ++g) {
        const int s = g % NS;
        mbar_wait(empty_bar(s), ((g / NS) & 1) ^ 1);
        const uint32_t a_dst = sm_addr + s * G_STAGE + dst_base;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const bool ok = src[i] != nullptr;
          cp_async16(a_dst + i * 2048, ok ? (const void*)(src[i] + kc * TC_BK) : (const void*)Ab,
                     ok ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(full_bar(s));
      }
    }
  } else if (warp == 8) {
    // ======================= TMA producer (weights) ====================================
    if (lane == 0) {
      int g = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int n0 = (tile % p.n_blocks) * D;
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(empty_bar(s), ((g / NS) & 1) ^ 1);
          mbar_arrive_expect_tx(full_bar(s), TC_B_TILE);
          tma_load_2d(sm_addr + s * G_STAGE + TC_A_TILE, &wmap, full_bar(s), kc * TC_BK, n0);
        }
      }
    }
    __syncwarp();
  } else {
    // ======================= MMA issuer ================================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128_n256();
      int g = 0, it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(tempty_bar(buf), ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        TL_STAMP(2, it, 0);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(full_bar(s), (g / NS) & 1);
          if (kc == 0) TL_STAMP(3, it, 0);
          tc_fence_after();
          const uint32_t a_stage = sm_addr + s * G_STAGE;
          mma_chunk<1>(a_stage, a_stage + TC_A_TILE, tmem_base + buf * D, 1, kc == 0, idesc,
                       empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
        TL_STAMP(2, it, 1);
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<2 * D>(tmem_base);
}

// ===================================================================================
// On-demand value_proj conv + bilinear/attention combine, one CTA per scene
// (GridSampleCrossBEVAttention, modules/blocks.py:110-126):
//   V[r, :] = ReLU(conv3x3(bev)[pixel r] + bias)   for the scene's unique sampled pixels r
//   S[a, :] = sum_{p,corner} w[a,p,corner] * V[slot[a,p,corner], :]
// Rows are processed 256 at a time ("pass": 2 row tiles sharing each weight tile).  A rows are
// 3x3xC patches gathered from the NHWC bf16 map by cp.async (zero padded); V never leaves the
// SM: the accumulators are drained from TMEM to a shared-memory staging area (aliasing the
// idle pipeline buffers) half of N at a time and combined there.
// ===================================================================================
struct EntPair { int slot; float w; };



// VOUT = true (tc_convv_kernel's role, denoise steps after the first with PlanReuse): the rows are a
// contiguous share of the cross-scene list p.vrows (pixel index in the batch, value row) and the epilogue
// writes V rows (bias + ReLU, bf16) to p.vout instead of combining them; no entry tables.
template <bool VOUT>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_conv_kernel(const GemmParams p, const __grid_constant__ CUtensorMap wmap) {
  constexpr int NS = C_NS, NT = C_NT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int scene = blockIdx.x;
  int nu, r0 = 0;
  if (VOUT) {
    const int count = __ldg(p.n_vrows);
    const int per = (((count + (int)gridDim.x - 1) / (int)gridDim.x) + 15) & ~15;
    r0 = (int)blockIdx.x * per;
    nu = min(per, count - r0);
    if (nu <= 0) return;
  } else {
    nu = p.nuniq[scene];
  }
  const int A = p.n_anchor, n_ent = p.n_anchor * p.ent_per_anchor;
  float* S32 = p.epi.out_f32;
  __nv_bfloat16* S16 = p.epi.out_bf16;
  const int HWp = p.H * p.W_;
  using off_type = typename std::conditional<VOUT, long long, int>::type;
  auto row_src = [&](int r) -> int {   // r already clamped to the CTA's rows
    return VOUT ? __ldg(&p.vrows[r0 + r].x) : __ldg(p.upix + (size_t)scene * p.rcap + r);
  };

  if (!VOUT && nu == 0) {   // every sample point fell outside the grid: grid_sample returns zeros
    for (int i = threadIdx.x; i < A * D; i += TC_THREADS) {
      if (S32) S32[(size_t)scene * A * D + i] = 0.f;
      if (S16) S16[(size_t)scene * A * D + i] = __float2bfloat16_rn(0.f);
    }
    return;
  }
  const int passes = (nu + NT * TC_BM - 1) / (NT * TC_BM);
  const int ent_bytes = ((n_ent * 8 + 15) / 16) * 16;
  EntPair* ent = reinterpret_cast<EntPair*>(sm + C_PIPE);
  float* bias_s = reinterpret_cast<float*>(sm + C_PIPE + ent_bytes);          // [256]
  const TcBars bars{sm_addr + C_PIPE + ent_bytes + D * 4, NS};
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(sm + C_PIPE + ent_bytes + D * 4 + (2 * NS + 2) * 8);
  constexpr int KC = 9 * (D / TC_BK);   // 36 k-chunks: (tap, 64-channel chunk)

  // Row coordinates of the first pass: 16 independent loads per producer thread, issued before
  // anything else so their latency hides under the barrier / TMEM setup (index clamped, the
  // validity is re-derived from nu).
  constexpr int RPT = NT * 8;
  int yx0[RPT];
  if (warp < 4) {
    const int rb0 = threadIdx.x >> 3;
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      yx0[i] = row_src(min(rb0 + 16 * i, nu - 1));
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars.full(s), 128 + 1);
      mbar_init(bars.empty(s), 1);
    }
    mbar_init(bars.accum(), 1);
    mbar_init(bars.passgo(), 128);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<NT * D>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 4 && lane == 0) tma_prefetch_desc(&wmap);
  if (threadIdx.x >= 128 && threadIdx.x < 192)
    reinterpret_cast<float4*>(bias_s)[threadIdx.x - 128] =
        __ldg(reinterpret_cast<const float4*>(p.epi.bias) + (threadIdx.x - 128));
  {
    // entry table: 4 (slot, weight) pairs in flight per thread and batch
    const int* es = p.ent_slot + (size_t)scene * n_ent;
    const float* ew = p.ent_w + (size_t)scene * n_ent;
    for (int base = threadIdx.x; base < n_ent; base += TC_THREADS * 4) {
      EntPair e[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * TC_THREADS;
        if (i < n_ent) { e[u].slot = __ldg(es + i); e[u].w = __ldg(ew + i); }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = base + u * TC_THREADS;
        if (i < n_ent) ent[i] = e[u];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    const int tid = threadIdx.x;
    const int j = tid & 7, rb = tid >> 3;
    const __nv_bfloat16* bev = reinterpret_cast<const __nv_bfloat16*>(p.bev) +
                               (VOUT ? (size_t)0 : (size_t)scene * p.H * p.W_ * D);
    float* Vs = reinterpret_cast<float*>(sm);
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    int g = 0;
    for (int pass = 0; pass < passes; ++pass) {
      const int row_base = pass * NT * TC_BM;
      const int rows_valid = min(NT * TC_BM, nu - row_base);
      const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
      // ---------------- A producers: gather 3x3xC patches
      // Loop-invariant per thread: the swizzled destination (row r = rb + 16 i always has
      // r & 7 == rb & 7, and consecutive i are 2048 B apart), the element offset of the row's
      // centre pixel and a 9-bit mask of the taps that fall inside the map.
      off_type rowoff[RPT];
      uint32_t vmask[RPT];
      if (pass > 0) {
#pragma unroll
        for (int i = 0; i < RPT; ++i)
          yx0[i] = row_src(min(row_base + rb + 16 * i, nu - 1));
      }
#pragma unroll
      for (int i = 0; i < RPT; ++i) {
        const int r = rb + 16 * i;
        rowoff[i] = 0;
        vmask[i] = 0;
        if (r < rows_valid) {
          int y, x;
          if (VOUT) {   // pixel index in the batch
            const int rem = yx0[i] % HWp;
            y = rem / p.W_; x = rem - y * p.W_;
            rowoff[i] = (off_type)yx0[i] * D + j * 8;
          } else {      // (y << 16) | x
            y = yx0[i] >> 16; x = yx0[i] & 0xffff;
            rowoff[i] = (y * p.W_ + x) * D + j * 8;
          }
          const uint32_t xm = (x > 0 ? 1u : 0u) | 2u | (x + 1 < p.W_ ? 4u : 0u);
          vmask[i] = (y > 0 ? xm : 0u) | (xm << 3) | (y + 1 < p.H ? (xm << 6) : 0u);
        }
      }
      const uint32_t dst_base = rb * 128 + ((j ^ (rb & 7)) << 4);
      for (int kc = 0; kc < KC; ++kc, ++g) {
        const int s = g % NS;
        mbar_wait(bars.empty(s), ((g / NS) & 1) ^ 1);
        if (tid == 0) TL_STAMP(0, g, 0);
        const uint32_t a_dst = sm_addr + s * C_STAGE + dst_base;
        const int tap = kc >> 2;
        const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
        const int tapoff = (dy * p.W_ + dx) * D + (kc & 3) * TC_BK;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (i < 8 * nt_active) {
            const bool ok = (vmask[i] >> tap) & 1u;
            const off_type off = ok ? rowoff[i] + tapoff : 0;
            cp_async16(a_dst + i * 2048, bev + off, ok ? 16u : 0u);
          }
        }
        if (kc == 0) {
          // Accumulators start at the conv bias (every MMA accumulates), so the drain is only
          // ReLU + store.  Done here, after the first chunk's copies are in flight and before
          // this thread's arrival lets the first MMA start; bias comes from shared memory.
#pragma unroll 1
          for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t u[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint4 t4 = *reinterpret_cast<const uint4*>(bias_s + cb * 32 + 4 * q);
              u[4 * q + 0] = t4.x; u[4 * q + 1] = t4.y; u[4 * q + 2] = t4.z; u[4 * q + 3] = t4.w;
            }
            for (int t = 0; t < nt_active; ++t) tmem_st32(trow + t * D + cb * 32, u);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        cp_async_mbar_arrive_noinc(bars.full(s));
        if (tid == 0) TL_STAMP(0, g, 1);
      }
      // ---------------- epilogue: drain TMEM -> smem, combine, half of N at a time
      if (tid == 0) TL_STAMP(3, pass * 4 + 0, 0);
      mbar_wait(bars.accum(), pass & 1);
      if (tid == 0) TL_STAMP(3, pass * 4 + 0, 1);
      tc_fence_after();
      if (VOUT) {
        // ---------------- value rows out: thread = row (TMEM lane), 512 contiguous bytes
        for (int t = 0; t < nt_active; ++t) {
          const int r = t * TC_BM + warp * 32 + lane;
          const bool rok = r < rows_valid;
          const int vrow = rok ? __ldg(&p.vrows[r0 + row_base + r].y) : 0;
          uint4* dst = reinterpret_cast<uint4*>(p.vout + (size_t)vrow * D);
#pragma unroll 1
          for (int cb = 0; cb < D / 32; ++cb) {
            uint32_t u0[32];
            tmem_ld32(trow + t * D + cb * 32, u0);
            tmem_ld_wait();
            if (rok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 w;
                uint32_t* wp = &w.x;
#pragma unroll
                for (int h2 = 0; h2 < 4; ++h2) {
                  const __nv_bfloat162 hh = __floats2bfloat162_rn(fmaxf(__uint_as_float(u0[8 * q + 2 * h2]), 0.f),
                                                                   fmaxf(__uint_as_float(u0[8 * q + 2 * h2 + 1]), 0.f));
                  wp[h2] = *reinterpret_cast<const uint32_t*>(&hh);
                }
                dst[cb * 4 + q] = w;
              }
            }
          }
        }
      }
      for (int half = 0; half < (VOUT ? 0 : 2); ++half) {
        for (int t = 0; t < nt_active; ++t) {
          float* vrow = Vs + (size_t)(t * TC_BM + warp * 32 + lane) * C_VS_LD;
#pragma unroll 1
          for (int cb = 0; cb < 4; cb += 2) {
            uint32_t u0[32], u1[32];
            const int col = half * 128 + cb * 32;
            tmem_ld32(trow + t * D + col, u0);
            tmem_ld32(trow + t * D + col + 32, u1);
            tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o;
              o.x = fmaxf(__uint_as_float(u0[4 * q + 0]), 0.f);
              o.y = fmaxf(__uint_as_float(u0[4 * q + 1]), 0.f);
              o.z = fmaxf(__uint_as_float(u0[4 * q + 2]), 0.f);
              o.w = fmaxf(__uint_as_float(u0[4 * q + 3]), 0.f);
              *reinterpret_cast<float4*>(vrow + cb * 32 + 4 * q) = o;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 o;
              o.x = fmaxf(__uint_as_float(u1[4 * q + 0]), 0.f);
              o.y = fmaxf(__uint_as_float(u1[4 * q + 1]), 0.f);
              o.z = fmaxf(__uint_as_float(u1[4 * q + 2]), 0.f);
              o.w = fmaxf(__uint_as_float(u1[4 * q + 3]), 0.f);
              *reinterpret_cast<float4*>(vrow + cb * 32 + 32 + 4 * q) = o;
            }
          }
        }
        named_bar_sync(1, 128);
        for (int a = warp; a < A; a += 4) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          const EntPair* ea = ent + a * p.ent_per_anchor;
#pragma unroll 8
          for (int k = 0; k < p.ent_per_anchor; ++k) {
            const EntPair e = ea[k];
            const int rr = e.slot - row_base;
            if (rr >= 0 && rr < rows_valid) {
              const float4 v = *reinterpret_cast<const float4*>(Vs + (size_t)rr * C_VS_LD + lane * 4);
              acc.x = fmaf(e.w, v.x, acc.x);
              acc.y = fmaf(e.w, v.y, acc.y);
              acc.z = fmaf(e.w, v.z, acc.z);
              acc.w = fmaf(e.w, v.w, acc.w);
            }
          }
          const size_t o = ((size_t)scene * A + a) * D + half * 128 + lane * 4;
          if (pass > 0) {   // same thread wrote it in the previous pass: deterministic RMW
            const float4 old = *reinterpret_cast<const float4*>(S32 + o);
            acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
          }
          if (pass + 1 < passes) {          // fp32 partial sums only between passes
            *reinterpret_cast<float4*>(S32 + o) = acc;
          } else if (S16) {                 // final value: the bf16 A operand of output_proj
            __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(acc.z, acc.w);
            uint2 u2;
            u2.x = *reinterpret_cast<uint32_t*>(&h0);
            u2.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(S16 + o) = u2;
          }
        }
        named_bar_sync(1, 128);
      }
      if (tid == 0) TL_STAMP(3, pass * 4 + 1, 0);
      // staging (generic proxy) is done: the weight TMA of the next pass may overwrite it
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bars.passgo());
    }
  } else if (warp == 4) {
    if (lane == 0) {
      int g = 0;
      for (int pass = 0; pass < passes; ++pass) {
        if (pass > 0) mbar_wait(bars.passgo(), (pass - 1) & 1);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(bars.empty(s), ((g / NS) & 1) ^ 1);
          TL_STAMP(1, g, 0);
          mbar_arrive_expect_tx(bars.full(s), TC_B_TILE);
          tma_load_2d(sm_addr + s * C_STAGE + NT * TC_A_TILE, &wmap, bars.full(s), kc * TC_BK, 0);
          TL_STAMP(1, g, 1);
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128_n256();
      int g = 0;
      for (int pass = 0; pass < passes; ++pass) {
        const int rows_valid = min(NT * TC_BM, nu - pass * NT * TC_BM);
        const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(bars.full(s), (g / NS) & 1);
          TL_STAMP(2, g, 0);
          tc_fence_after();
          const uint32_t a_stage = sm_addr + s * C_STAGE;
          mma_chunk<NT>(a_stage, a_stage + NT * TC_A_TILE, tmem_base, nt_active, false, idesc,
                        bars.empty(s));
          TL_STAMP(2, g, 1);
        }
        umma_commit(bars.accum());
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<NT * D>(tmem_base);
}

// ===================================================================================
// Persistent variant of the conv (default): one CTA per SM walks scenes blockIdx.x, + gridDim.x, ...
// so that barrier / TMEM set-up is paid once, the next scene's pixel list is fetched under the
// current scene's main loop, and EIGHT dedicated epilogue warps (two per TMEM lane quarter, one
// 128-row tile each) drain and combine, instead of the four producer warps switching roles.
//   warps 0-3   A producers (cp.async gather)        warp 4  TMA (weights)     warp 5  MMA issuer
//   warps 6-13  epilogue: bias + ReLU drain of tile (w - 6) / 4 into the staging area (aliasing the
//               idle pipeline buffers), bilinear * attention combine with the anchors dealt over
//               the eight warps; the scene's entry table is fetched while the main loop runs
// Same arithmetic as tc_conv_kernel except that the bias is added at the drain (acc + bias) instead
// of seeding the accumulators.
// ===================================================================================
// timeline (p.dbg != nullptr): CTA 0 stamps clock64 during the first pass of its second scene
#define C2_STAMP(cond, idx) do { if (p.dbg && blockIdx.x == 0 && (cond)) p.dbg[idx] = clock64(); } while (0)
constexpr int C2_THREADS = 448;
constexpr int C2_EPI = 256;

template <bool TF32>
__global__ void __launch_bounds__(C2_THREADS, 1)
tc_conv2_kernel(const GemmParams p, const __grid_constant__ CUtensorMap wmap, int B) {
  constexpr int NS = C_NS, NT = C_NT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int A = p.n_anchor, n_ent = p.n_anchor * p.ent_per_anchor;
  float* S32 = p.epi.out_f32;
  __nv_bfloat16* S16 = p.epi.out_bf16;
  const int ent_bytes = ((n_ent * 8 + 15) / 16) * 16;
  EntPair* ent = reinterpret_cast<EntPair*>(sm + C_PIPE);
  float* bias_s = reinterpret_cast<float*>(sm + C_PIPE + ent_bytes);          // [256]
  const TcBars bars{sm_addr + C_PIPE + ent_bytes + D * 4, NS};
  volatile uint32_t* tmem_slot =
      reinterpret_cast<volatile uint32_t*>(sm + C_PIPE + ent_bytes + D * 4 + (2 * NS + 2) * 8);
  // bf16: 36 k-chunks (tap, 64-channel chunk).  TF32 (fp32 engine, 3xTF32): a k-chunk is 32 fp32 channels
  // (the same 128 bytes per row), 72 per operand-plane combination, and three combinations run into
  // the same accumulators, small terms first: (a_hi, w_lo), (a_lo, w_hi), (a_hi, w_hi), where x_hi keeps
  // the 10 mantissa bits the tensor core reads and x_lo = x - x_hi (exact): the dropped a_lo . w_lo is
  // ~2^-22 of the product, i.e. fp32-level accuracy
  constexpr int ES = TF32 ? 4 : 2;                 // bytes per element
  constexpr int CPT = D * ES / 128;                // k-chunks per tap
  constexpr int KC1 = 9 * CPT;                     // k-chunks of one plane combination
  constexpr int KC = TF32 ? 3 * KC1 : KC1;
  constexpr int RPT = NT * 8;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(bars.full(s), 128 + 1);
      mbar_init(bars.empty(s), 1);
    }
    mbar_init(bars.accum(), 1);
    mbar_init(bars.passgo(), C2_EPI);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<NT * D>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 4 && lane == 0) tma_prefetch_desc(&wmap);
  if (threadIdx.x < 64)
    reinterpret_cast<float4*>(bias_s)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(p.epi.bias) + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ======================= A producers: gather 3x3xC patches =========================
    const int tid = threadIdx.x;
    const int j = tid & 7, rb = tid >> 3;
    const uint32_t dst_base = rb * 128 + ((j ^ (rb & 7)) << 4);
    int g = 0, pi = 0, sidx = 0;
    int scene = blockIdx.x;
    int nu = scene < B ? __ldg(p.nuniq + scene) : 0;
    int yx0[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      yx0[i] = (scene < B && nu > 0) ? __ldg(p.upix + (size_t)scene * p.rcap + min(rb + 16 * i, nu - 1)) : 0;
    while (scene < B) {
      // the next scene's pixel list (first pass) is fetched now and consumed after this scene
      const int scene_n = scene + gridDim.x;
      const int nu_n = scene_n < B ? __ldg(p.nuniq + scene_n) : 0;
      int yxn[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i)
        yxn[i] = (scene_n < B && nu_n > 0) ? __ldg(p.upix + (size_t)scene_n * p.rcap + min(rb + 16 * i, nu_n - 1)) : 0;
      const uint8_t* bev = reinterpret_cast<const uint8_t*>(p.bev) + (size_t)scene * p.H * p.W_ * D * ES;
      const uint8_t* bev_lo = TF32 ? reinterpret_cast<const uint8_t*>(p.bev_lo) + (size_t)scene * p.H * p.W_ * D * ES : bev;
      const int passes = (nu + NT * TC_BM - 1) / (NT * TC_BM);
      for (int pass = 0; pass < passes; ++pass, ++pi) {
        const int row_base = pass * NT * TC_BM;
        const int rows_valid = min(NT * TC_BM, nu - row_base);
        const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
        if (pass > 0) {
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            yx0[i] = __ldg(p.upix + (size_t)scene * p.rcap + min(row_base + rb + 16 * i, nu - 1));
        }
        int rowoff[RPT];
        uint32_t vmask[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = rb + 16 * i;
          rowoff[i] = 0;
          vmask[i] = 0;
          if (r < rows_valid) {
            const int yx = yx0[i];   // (y << 16) | x
            const int y = yx >> 16, x = yx & 0xffff;
            rowoff[i] = (y * p.W_ + x) * D * ES + j * 16;   // bytes
            const uint32_t xm = (x > 0 ? 1u : 0u) | 2u | (x + 1 < p.W_ ? 4u : 0u);
            vmask[i] = (y > 0 ? xm : 0u) | (xm << 3) | (y + 1 < p.H ? (xm << 6) : 0u);
          }
        }
        // the staging area of the previous pass aliases the pipeline buffers
        C2_STAMP(sidx == 1 && pass == 0 && tid == 0, 0);
        if (pi > 0) mbar_wait(bars.passgo(), (uint32_t)(pi - 1) & 1u);
        C2_STAMP(sidx == 1 && pass == 0 && tid == 0, 1);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(bars.empty(s), ((g / NS) & 1) ^ 1);
          C2_STAMP(sidx == 1 && pass == 0 && tid == 0 && (kc & 3) == 0, 2 + (kc >> 2));
          const uint32_t a_dst = sm_addr + s * C_STAGE + dst_base;
          const int combo = kc / KC1, kc1 = kc - combo * KC1;
          const int tap = kc1 / CPT;
          const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          const int tapoff = (dy * p.W_ + dx) * D * ES + (kc1 - tap * CPT) * 128;   // bytes
          const uint8_t* plane = (TF32 && combo == 1) ? bev_lo : bev;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            if (i < 8 * nt_active) {
              const bool ok = (vmask[i] >> tap) & 1u;
              const int off = ok ? rowoff[i] + tapoff : 0;
              cp_async16(a_dst + i * 2048, plane + off, ok ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive_noinc(bars.full(s));
        }
      }
      scene = scene_n;
      nu = nu_n;
      ++sidx;
#pragma unroll
      for (int i = 0; i < RPT; ++i) yx0[i] = yxn[i];
    }
  } else if (warp == 4) {
    // ======================= TMA producer (weights) ====================================
    if (lane == 0) {
      int g = 0, pi = 0;
      for (int scene = blockIdx.x; scene < B; scene += gridDim.x) {
        const int nu = __ldg(p.nuniq + scene);
        const int passes = (nu + NT * TC_BM - 1) / (NT * TC_BM);
        for (int pass = 0; pass < passes; ++pass, ++pi) {
          if (pi > 0) mbar_wait(bars.passgo(), (uint32_t)(pi - 1) & 1u);
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int s = g % NS;
            mbar_wait(bars.empty(s), ((g / NS) & 1) ^ 1);
            mbar_arrive_expect_tx(bars.full(s), TC_B_TILE);
            // element coordinate along K; TF32: the packed matrix is [256][w_hi (K) | w_lo (K)]
            const int combo = kc / KC1, kc1 = kc - combo * KC1;
            const int k0 = TF32 ? ((combo == 0 ? p.K : 0) + kc1 * 32) : kc * TC_BK;
            tma_load_2d(sm_addr + s * C_STAGE + NT * TC_A_TILE, &wmap, bars.full(s), k0, 0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ======================= MMA issuer ================================================
    if (lane == 0) {
      const uint32_t idesc = TF32 ? umma_idesc_tf32_m128_n256() : umma_idesc_bf16_m128_n256();
      int g = 0, sidx = 0;
      for (int scene = blockIdx.x; scene < B; scene += gridDim.x, ++sidx) {
        const int nu = __ldg(p.nuniq + scene);
        const int passes = (nu + NT * TC_BM - 1) / (NT * TC_BM);
        for (int pass = 0; pass < passes; ++pass) {
          const int rows_valid = min(NT * TC_BM, nu - pass * NT * TC_BM);
          const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
          // (the accumulators of the previous pass are drained: the operands of this pass were only
          // produced after the producers saw its epilogue finish)
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int s = g % NS;
            mbar_wait(bars.full(s), (g / NS) & 1);
            tc_fence_after();
            C2_STAMP(sidx == 1 && pass == 0 && (kc & 3) == 0, 16 + (kc >> 2));
            const uint32_t a_stage = sm_addr + s * C_STAGE;
            mma_chunk<NT, TF32>(a_stage, a_stage + NT * TC_A_TILE, tmem_base, nt_active, kc == 0, idesc,
                                bars.empty(s));
          }
          umma_commit(bars.accum());
          C2_STAMP(sidx == 1 && pass == 0, 25);
        }
      }
    }
    __syncwarp();
  } else {
    // ======================= epilogue warps ============================================
    const int ew = warp - 6;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    // warp w may only touch TMEM lanes 32 * (w % 4) ..; warps 6..9 cover the four quarters of tile 0,
    // warps 10..13 those of tile 1
    const int tile_of_warp = (ew < 4) ? 0 : 1;
    const int etid = threadIdx.x - 6 * 32;   // 0..255
    float* Vs = reinterpret_cast<float*>(sm);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + tile_of_warp * D;
    int pi = 0, sidx = 0;
    for (int scene = blockIdx.x; scene < B; scene += gridDim.x, ++sidx) {
      const int nu = __ldg(p.nuniq + scene);
      if (nu == 0) {   // every sample point fell outside the grid: grid_sample returns zeros
        for (int i = etid; i < A * D; i += C2_EPI) {
          if (S32) S32[(size_t)scene * A * D + i] = 0.f;
          if (S16) S16[(size_t)scene * A * D + i] = __float2bfloat16_rn(0.f);
        }
        continue;
      }
      {  // entry table of this scene (the previous scene's combine is behind a barrier of these warps)
        const int* es = p.ent_slot + (size_t)scene * n_ent;
        const float* ew_ = p.ent_w + (size_t)scene * n_ent;
        for (int base = etid; base < n_ent; base += C2_EPI * 4) {
          EntPair e[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = base + u * C2_EPI;
            if (i < n_ent) { e[u].slot = __ldg(es + i); e[u].w = __ldg(ew_ + i); }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = base + u * C2_EPI;
            if (i < n_ent) ent[i] = e[u];
          }
        }
      }
      named_bar_sync(1, C2_EPI);
      const int passes = (nu + NT * TC_BM - 1) / (NT * TC_BM);
      for (int pass = 0; pass < passes; ++pass, ++pi) {
        const int row_base = pass * NT * TC_BM;
        const int rows_valid = min(NT * TC_BM, nu - row_base);
        const int nt_active = (rows_valid + TC_BM - 1) / TC_BM;
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 32);
        mbar_wait(bars.accum(), (uint32_t)pi & 1u);
        tc_fence_after();
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 33);
        for (int half = 0; half < 2; ++half) {
          if (tile_of_warp < nt_active) {
            float* vrow = Vs + (size_t)(tile_of_warp * TC_BM + q * 32 + lane) * C_VS_LD;
#pragma unroll 1
            for (int cb = 0; cb < 4; cb += 2) {
              uint32_t u0[32], u1[32];
              const int col = half * 128 + cb * 32;
              tmem_ld32(trow + col, u0);
              tmem_ld32(trow + col + 32, u1);
              tmem_ld_wait();
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col + 4 * qq);
                float4 o;
                o.x = fmaxf(__uint_as_float(u0[4 * qq + 0]) + b4.x, 0.f);
                o.y = fmaxf(__uint_as_float(u0[4 * qq + 1]) + b4.y, 0.f);
                o.z = fmaxf(__uint_as_float(u0[4 * qq + 2]) + b4.z, 0.f);
                o.w = fmaxf(__uint_as_float(u0[4 * qq + 3]) + b4.w, 0.f);
                *reinterpret_cast<float4*>(vrow + cb * 32 + 4 * qq) = o;
              }
#pragma unroll
              for (int qq = 0; qq < 8; ++qq) {
                const float4 b4 = *reinterpret_cast<const float4*>(bias_s + col + 32 + 4 * qq);
                float4 o;
                o.x = fmaxf(__uint_as_float(u1[4 * qq + 0]) + b4.x, 0.f);
                o.y = fmaxf(__uint_as_float(u1[4 * qq + 1]) + b4.y, 0.f);
                o.z = fmaxf(__uint_as_float(u1[4 * qq + 2]) + b4.z, 0.f);
                o.w = fmaxf(__uint_as_float(u1[4 * qq + 3]) + b4.w, 0.f);
                *reinterpret_cast<float4*>(vrow + cb * 32 + 32 + 4 * qq) = o;
              }
            }
          }
          C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 34 + 3 * half);
          named_bar_sync(1, C2_EPI);
          C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 35 + 3 * half);
          for (int a = ew; a < A; a += 8) {
            // lane k fetches entry k of the anchor once; (row, weight) pairs are then broadcast by
            // shuffles, rows outside this pass contribute weight 0 on a clamped row: branch-free,
            // eight independent staging reads in flight (same summation order as the entry order)
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const EntPair* ea = ent + a * p.ent_per_anchor;
            for (int k0 = 0; k0 < p.ent_per_anchor; k0 += 32) {
              EntPair mine;
              mine.slot = -1; mine.w = 0.f;
              if (k0 + lane < p.ent_per_anchor) mine = ea[k0 + lane];
              const int rr = mine.slot - row_base;
              const bool ok = rr >= 0 && rr < rows_valid;
              const int rrc = ok ? rr * C_VS_LD : 0;
              const float wv = ok ? mine.w : 0.f;
#pragma unroll 8
              for (int k = 0; k < 32; ++k) {
                const int off = __shfl_sync(0xffffffffu, rrc, k);
                const float wk = __shfl_sync(0xffffffffu, wv, k);
                const float4 v = *reinterpret_cast<const float4*>(Vs + off + lane * 4);
                acc.x = fmaf(wk, v.x, acc.x);
                acc.y = fmaf(wk, v.y, acc.y);
                acc.z = fmaf(wk, v.z, acc.z);
                acc.w = fmaf(wk, v.w, acc.w);
              }
            }
            const size_t o = ((size_t)scene * A + a) * D + half * 128 + lane * 4;
            if (pass > 0) {   // same thread wrote it in the previous pass: deterministic RMW
              const float4 old = *reinterpret_cast<const float4*>(S32 + o);
              acc.x += old.x; acc.y += old.y; acc.z += old.z; acc.w += old.w;
            }
            if (pass + 1 < passes) {          // fp32 partial sums only between passes
              *reinterpret_cast<float4*>(S32 + o) = acc;
            } else if (S16) {                 // final value: the bf16 A operand of output_proj
              __nv_bfloat162 h0 = __floats2bfloat162_rn(acc.x, acc.y);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(acc.z, acc.w);
              uint2 u2;
              u2.x = *reinterpret_cast<uint32_t*>(&h0);
              u2.y = *reinterpret_cast<uint32_t*>(&h1);
              *reinterpret_cast<uint2*>(S16 + o) = u2;
            } else {                          // fp32 engine: S feeds the fp32 output_proj
              *reinterpret_cast<float4*>(S32 + o) = acc;
            }
          }
          C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 36 + 3 * half);
          named_bar_sync(1, C2_EPI);
        }
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 40);
        // staging (generic proxy) is done: the weight TMA of the next pass may overwrite it, and
        // the accumulators may be overwritten by the next pass's first MMA
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(bars.passgo());
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<NT * D>(tmem_base);
}

// ===================================================================================
// tc_conv3_kernel: persistent conv with the bilinear x attention combine ON THE TENSOR CORE.
//
// The main loop swaps the operand roles of tc_conv2_kernel: the weight tile [256 channels x 64 k]
// is the A operand (two M = 128 channel tiles) and the gathered patches [<= 256 pixel rows x 64 k]
// are the B operand (N = 256, or 128 when the pass has <= 128 rows), so the accumulators hold
// V^T: TMEM lane = output channel, column = pixel row.  A thread of the epilogue warps then owns one
// channel: bias + ReLU + bf16 and its row of V^T goes straight into the K-major, 128-byte-swizzled
// layout of a B operand [256 channels][K = pixel rows] (16-byte vector stores, the same pattern as
// the chain kernel's operand writes).  The combine
//     S[a, c] = sum_r Wc[a, r] * V[r, c]          (Wc: merged weights per anchor: the epilogue warps
//                                                  build them from the entry tables while the main
//                                                  loop runs -- duplicates of a pixel summed in entry
//                                                  order by MATCH -- and scatter them after it)
// is 8-16 more tcgen05.mma (A = Wc [128 anchor rows (only n_anchor valid) x K], B = V^T), whose
// accumulator rows (lane = anchor) are written out as bf16 S.  The CUDA-core combine of
// tc_conv2_kernel read 655 KB of staged fp32 per scene and was shared-memory-bandwidth bound
// (~10 k cycles); here the epilogue is one drain pass plus 16 MMAs.
// Numerics: V and the combine weights are rounded to bf16 before the weighted sum (fp32 accumulate).
// Shared memory: V^T aliases pipeline stages 0-1 (128 KiB), Wc stage 2 (4 x 16 KiB K-chunks, only
// the first n_anchor rows of each are written; the MMA's other rows read stale bytes whose
// products land in accumulator rows nobody reads).
// ===================================================================================
constexpr int C3_VT_CHUNK = 256 * 128;     // V^T operand: 64 pixel rows (K) x 256 channels
constexpr int C3_WC_OFF = 2 * C_STAGE;     // Wc operand region (pipeline stage 2)
constexpr int C3_THREADS = C2_THREADS + 32;   // + warp 14: second MMA issuer (channel tile 1)

// A scene's unique pixel rows are split into ceil(nu / 256) passes of EQUAL size, rounded up to the MMA's
// N granularity (16): 276 rows run as 2 x 144 instead of 256 + 128-padded, and a single-pass scene uses
// N = roundup16(nu) instead of 256 -- the main loop is tensor-bound, so its time is proportional to N.
struct C3Split { int passes, rpp; };
__device__ __forceinline__ C3Split c3_split(int nu) {
  C3Split s;
  s.passes = (nu + 255) / 256;
  s.rpp = s.passes ? (((nu + s.passes - 1) / s.passes + 15) & ~15) : 0;
  return s;
}

__global__ void __launch_bounds__(C3_THREADS, 1)
tc_conv3_kernel(const GemmParams p, const __grid_constant__ CUtensorMap wmap, int B) {
  constexpr int NS = C_NS, NT = C_NT;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int A = p.n_anchor;
  float* S32 = p.epi.out_f32;
  __nv_bfloat16* S16 = p.epi.out_bf16;
  float* bias_s = reinterpret_cast<float*>(sm + C_PIPE);          // [256]
  const uint32_t bar0 = sm_addr + C_PIPE + D * 4;
  auto full = [&](int s) { return bar0 + s * 8; };
  auto empty = [&](int s) { return bar0 + (NS + s) * 8; };
  const uint32_t accum = bar0 + 2 * NS * 8, vt_ready = accum + 8, comb_done = accum + 16, passgo = accum + 24;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + C_PIPE + D * 4 + (2 * NS + 4) * 8);
  // Scene queue.  The k-th scene of this CTA is blockIdx.x for k = 0 and comes from a global counter
  // after that (p.sched; static round-robin when it is null): scenes differ in unique pixels and
  // passes, and with ~28 scenes per CTA a static deal left the SMs idle for ~12 % of the launch
  // waiting for the slowest one.  MMA issuer 0 requests scene k + 2 when it starts scene k and
  // publishes it (slot (k + 1) & 3, one mbarrier phase per use) after the scene's last MMA, so the
  // atomic's round trip hides under the main loop; every role reads the queue in order.  A slot is
  // reused four scenes later: by then every role has read it (no role is more than one scene away
  // from the issuer: producers wait for the epilogue's passgo, the epilogue for the issuer's commits).
  static_assert((2 * NS + 4) * 8 + 4 <= 96 && 144 <= TC_BAR_BYTES, "barrier area layout");
  const uint32_t qf0 = bar0 + 96;
  volatile int* sq = reinterpret_cast<volatile int*>(sm + C_PIPE + D * 4 + 128);
  auto scene_at = [&](int kk) -> int {
    if (kk == 0) return (int)blockIdx.x;
    mbar_wait(qf0 + ((kk - 1) & 3) * 8, (uint32_t)((kk - 1) >> 2) & 1u);
    return sq[(kk - 1) & 3];
  };
  constexpr int KC = 9 * (D / TC_BK);
  constexpr int RPT = NT * 8;

  if (threadIdx.x == 0) {
    // two MMA issuers (a tcgen05.mma costs its issuing thread ~130-160 cycles whatever its shape;
    // threads of different warps issue concurrently): each releases a stage / commits the pass
    for (int s = 0; s < NS; ++s) { mbar_init(full(s), 128 + 1); mbar_init(empty(s), 2); }
    mbar_init(accum, 2);
    mbar_init(vt_ready, C2_EPI);
    mbar_init(comb_done, 1);
    mbar_init(passgo, C2_EPI);
    for (int i = 0; i < 4; ++i) mbar_init(qf0 + i * 8, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<NT * D>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 4 && lane == 0) tma_prefetch_desc(&wmap);
  if (threadIdx.x < 64)
    reinterpret_cast<float4*>(bias_s)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(p.epi.bias) + threadIdx.x);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ======================= producers: gather 3x3xC patches (B operand) ===============
    const int tid = threadIdx.x;
    const int j = tid & 7, rb = tid >> 3;
    const uint32_t dst_base = rb * 128 + ((j ^ (rb & 7)) << 4);
    int g = 0, pi = 0, sidx = 0;
    int scene = scene_at(0);
    int nu = scene < B ? __ldg(p.nuniq + scene) : 0;
    int yx0[RPT];
#pragma unroll
    for (int i = 0; i < RPT; ++i)
      yx0[i] = (scene < B && nu > 0) ? __ldg(p.upix + (size_t)scene * p.rcap + min(rb + 16 * i, nu - 1)) : 0;
    while (scene < B) {
      const int scene_n = scene_at(sidx + 1);
      const int nu_n = scene_n < B ? __ldg(p.nuniq + scene_n) : 0;
      int yxn[RPT];
#pragma unroll
      for (int i = 0; i < RPT; ++i)
        yxn[i] = (scene_n < B && nu_n > 0) ? __ldg(p.upix + (size_t)scene_n * p.rcap + min(rb + 16 * i, nu_n - 1)) : 0;
      const __nv_bfloat16* bev = reinterpret_cast<const __nv_bfloat16*>(p.bev) + (size_t)scene * p.H * p.W_ * D;
      const C3Split sp = c3_split(nu);
      const int passes = sp.passes, ngrp = sp.rpp >> 4;      // ngrp: 16-row groups per pass (N / 16)
      for (int pass = 0; pass < passes; ++pass, ++pi) {
        const int row_base = pass * sp.rpp;
        const int rows_valid = min(sp.rpp, nu - row_base);
        if (pass > 0) {
#pragma unroll
          for (int i = 0; i < RPT; ++i)
            yx0[i] = __ldg(p.upix + (size_t)scene * p.rcap + min(row_base + rb + 16 * i, nu - 1));
        }
        int rowoff[RPT];
        uint32_t vmask[RPT];
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const int r = rb + 16 * i;
          rowoff[i] = 0;
          vmask[i] = 0;
          if (r < rows_valid) {
            const int yx = yx0[i];
            const int y = yx >> 16, x = yx & 0xffff;
            DDH_ASSERT(y >= 0 && y < p.H && x >= 0 && x < p.W_);
            rowoff[i] = (y * p.W_ + x) * D + j * 8;
            const uint32_t xm = (x > 0 ? 1u : 0u) | 2u | (x + 1 < p.W_ ? 4u : 0u);
            vmask[i] = (y > 0 ? xm : 0u) | (xm << 3) | (y + 1 < p.H ? (xm << 6) : 0u);
          }
        }
        C2_STAMP(sidx == 1 && pass == 0 && tid == 0, 0);
        if (pi > 0) mbar_wait(passgo, (uint32_t)(pi - 1) & 1u);   // V^T / Wc of the previous pass alias the pipeline
        C2_STAMP(sidx == 1 && pass == 0 && tid == 0, 1);
        for (int kc = 0; kc < KC; ++kc, ++g) {
          const int s = g % NS;
          mbar_wait(empty(s), ((g / NS) & 1) ^ 1);
          C2_STAMP(sidx == 1 && pass == 0 && tid == 0 && (kc & 3) == 0, 2 + (kc >> 2));
          const uint32_t a_dst = sm_addr + s * C_STAGE + dst_base;
          const int tap = kc >> 2;
          const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          const int tapoff = (dy * p.W_ + dx) * D + (kc & 3) * TC_BK;
#pragma unroll
          for (int i = 0; i < RPT; ++i) {
            if (i < ngrp) {
              const bool ok = (vmask[i] >> tap) & 1u;
              const int off = ok ? rowoff[i] + tapoff : 0;
              cp_async16(a_dst + i * 2048, bev + off, ok ? 16u : 0u);
            }
          }
          cp_async_mbar_arrive_noinc(full(s));
        }
      }
      scene = scene_n;
      nu = nu_n;
      ++sidx;
#pragma unroll
      for (int i = 0; i < RPT; ++i) yx0[i] = yxn[i];
    }
  } else if (warp == 4) {
    // ======================= TMA producer (weights: A operand) =========================
    if (lane == 0) {
      int g = 0, pi = 0, sk = 0;
      for (int scene = scene_at(0); scene < B; scene = scene_at(++sk)) {
        const int nu = __ldg(p.nuniq + scene);
        const int passes = c3_split(nu).passes;
        for (int pass = 0; pass < passes; ++pass, ++pi) {
          if (pi > 0) mbar_wait(passgo, (uint32_t)(pi - 1) & 1u);
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int s = g % NS;
            mbar_wait(empty(s), ((g / NS) & 1) ^ 1);
            mbar_arrive_expect_tx(full(s), TC_B_TILE);
            tma_load_2d(sm_addr + s * C_STAGE + NT * TC_A_TILE, &wmap, full(s), kc * TC_BK, 0);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5 || warp == 14) {
    // ======================= MMA issuers: warp 5 channel tile 0 (+ the combine), warp 14 tile 1 ==
    if (lane == 0) {
      const int mt = warp == 5 ? 0 : 1;
      // combine: M = 64 (<= 64 anchor rows; an M = 64 instruction holds the tensor pipe half as long as
      // M = 128), N = 256 channels; accumulator row r lives in TMEM lane 32 (r / 16) + r % 16
      const uint32_t idesc256 = (1u << 4) | (1u << 7) | (1u << 10) | ((256u >> 3) << 17) | ((64u >> 4) << 24);
      int g = 0, sidx = 0;
      uint32_t pi = 0;
      // issuer 0 runs the scene queue: `req` = the scene requested at the start of the current one
      auto request = [&](int kk) -> int {   // kk-th scene of this CTA, kk >= 1
        const int sc = p.sched ? (int)gridDim.x + atomicAdd(p.sched, 1) : (int)blockIdx.x + kk * (int)gridDim.x;
        return sc < B ? sc : B;
      };
      auto publish = [&](int kk, int sc) {
        sq[(kk - 1) & 3] = sc;
        mbar_arrive(qf0 + ((kk - 1) & 3) * 8);
      };
      if (mt == 0) publish(1, request(1));
      for (int scene = scene_at(0); scene < B; scene = scene_at(++sidx)) {
        const int req = mt == 0 ? request(sidx + 2) : 0;
        const int nu = __ldg(p.nuniq + scene);
        const C3Split sp = c3_split(nu);
        const int passes = sp.passes;
        if (mt == 0 && passes == 0) publish(sidx + 2, req);
        // D = f32, A = B = bf16, K-major, M = 128, N = rows of a pass
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)sp.rpp >> 3) << 17) | ((128u >> 4) << 24);
        for (int pass = 0; pass < passes; ++pass, ++pi) {
          for (int kc = 0; kc < KC; ++kc, ++g) {
            const int s = g % NS;
            mbar_wait(full(s), (g / NS) & 1);
            tc_fence_after();
            C2_STAMP(mt == 0 && sidx == 1 && pass == 0 && (kc & 3) == 0, 16 + (kc >> 2));
            const uint32_t p_stage = sm_addr + s * C_STAGE;             // patches: B operand
            const uint32_t w_stage = p_stage + NT * TC_A_TILE;          // weights: A operand
#pragma unroll
            for (int k4 = 0; k4 < TC_BK / 16; ++k4)
              umma_bf16(tmem_base + mt * D, umma_desc_sw128(w_stage + mt * TC_A_TILE + k4 * 32),
                        umma_desc_sw128(p_stage + k4 * 32), idesc, (kc == 0 && k4 == 0) ? 0u : 1u);
            umma_commit(empty(s));
          }
          umma_commit(accum);
          C2_STAMP(mt == 0 && sidx == 1 && pass == 0, 25);
          if (mt == 0 && pass == 0) publish(sidx + 2, req);
          if (mt == 0) {
            // ---- combine: S^(pass)[a, c] = Wc[a, r] . V^T[c, r], K = the pass's pixel rows
            mbar_wait(vt_ready, pi & 1u);
            tc_fence_after();
            const int ksteps = sp.rpp >> 4;      // K = the pass's pixel rows, 16 per instruction
            for (int ks = 0; ks < ksteps; ++ks) {
              const int kc = ks >> 2, k4 = ks & 3;
              umma_bf16(tmem_base, umma_desc_sw128(sm_addr + C3_WC_OFF + kc * TC_A_TILE + k4 * 32),
                        umma_desc_sw128(sm_addr + kc * C3_VT_CHUNK + k4 * 32), idesc256, ks ? 1u : 0u);
            }
            umma_commit(comb_done);
            C2_STAMP(sidx == 1 && pass == 0, 26);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp < 14) {
    // ======================= epilogue warps ============================================
    const int ew = warp - 6;                 // 0..7
    const int q = warp & 3;                  // TMEM lane quarter
    const int mt = ew >> 2;                  // channel tile (drain) / column half (final store)
    const int etid = threadIdx.x - 6 * 32;   // 0..255
    const int c = mt * TC_BM + q * 32 + lane;   // the channel this thread drains
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const float bias_c = bias_s[c];
    uint32_t pi = 0;
    int sidx = 0;
    for (int scene = scene_at(0); scene < B; scene = scene_at(++sidx)) {
      const int nu = __ldg(p.nuniq + scene);
      if (nu == 0) {   // every sample point fell outside the grid: grid_sample returns zeros
        for (int i = etid; i < A * D; i += C2_EPI) {
          if (S32) S32[(size_t)scene * A * D + i] = 0.f;
          if (S16) S16[(size_t)scene * A * D + i] = __float2bfloat16_rn(0.f);
        }
        continue;
      }
      // ---- merged combine weights of this warp's anchors (a = ew + 8 i), built while the main loop
      // runs: lane k = entry k (pose, corner); entries of one anchor that hit the same unique pixel are
      // summed in entry order by the first of them (MATCH groups the lanes), so sums are deterministic
      int mslot[8];
      float msum[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int a = ew + 8 * i;
        mslot[i] = -1;
        msum[i] = 0.f;
        if (a < A) {
          const int slot = __ldg(p.ent_slot + ((size_t)scene * A + a) * 32 + lane);
          const float wv = __ldg(p.ent_w + ((size_t)scene * A + a) * 32 + lane);
          const unsigned grp = __match_any_sync(0xffffffffu, slot);
          const int most = __reduce_max_sync(0xffffffffu, __popc(grp));
          unsigned m = grp;
          float sum = 0.f;
          for (int t = 0; t < most; ++t) {
            const int src = m ? (__ffs((int)m) - 1) : lane;
            const float wk = __shfl_sync(0xffffffffu, wv, src);
            if (m) sum += wk;
            m &= m - 1;
          }
          if (slot >= 0 && (__ffs((int)grp) - 1) == lane) { mslot[i] = slot; msum[i] = sum; }
        }
      }
      const C3Split sp = c3_split(nu);
      const int passes = sp.passes;
      for (int pass = 0; pass < passes; ++pass, ++pi) {
        const int row_base = pass * sp.rpp;
        const int upa = sp.rpp >> 3;             // 16-byte units (8 pixel rows) per anchor row of Wc
        const int n_units = A * upa;
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 32);
        mbar_wait(accum, pi & 1u);
        tc_fence_after();
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 33);
        for (int u = etid; u < n_units; u += C2_EPI) {
          const int a = u / upa, k8 = u - a * upa;
          *reinterpret_cast<uint4*>(sm + C3_WC_OFF + (k8 >> 3) * TC_A_TILE + a * 128 + (((k8 & 7) ^ (a & 7)) << 4)) =
              make_uint4(0u, 0u, 0u, 0u);
        }
        named_bar_sync(1, C2_EPI);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int a = ew + 8 * i;
          const int rr = mslot[i] - row_base;
          DDH_ASSERT(mslot[i] < nu);
          if (mslot[i] >= 0 && rr >= 0 && rr < sp.rpp)
            *reinterpret_cast<__nv_bfloat16*>(sm + C3_WC_OFF + (rr >> 6) * TC_A_TILE + a * 128 +
                                              ((((rr & 63) >> 3) ^ (a & 7)) << 4) + (rr & 7) * 2) = __float2bfloat16_rn(msum[i]);
        }
        // ---- kept value rows (PlanReuse, p.vout): the pass's rows of V go out transposed back from the V^T
        // operand, one 8-pixel group per warp: lane = channel octet, eight 16-byte reads (channel (i + lane) & 7
        // of the octet at step i: the eight lanes of a quarter warp hit eight different swizzle units), a register
        // rotation by lane, an 8 x 8 transpose of bf16 (PRMT), and per pixel ONE 512-byte row written by the warp.
        // The SM pushes ~27 B/clk of stores (2-byte stores: 17 B/clk), i.e. ~4.6 k cycles per pass, of which the
        // combine wait below hides 2.4 k.  (Storing each 32-pixel block as soon as it is drained was slower: the
        // stores stall the issuing warps, which then hold up the drain barriers.)
        const int vn = min(sp.rpp, nu - row_base);
        uint4* vo = p.vout ? reinterpret_cast<uint4*>(p.vout + ((size_t)scene * p.vcap + row_base) * D) + lane : nullptr;
        auto store_group = [&](int k8) {
          const uint8_t* src = sm + (k8 >> 3) * C3_VT_CHUNK + lane * 1024;
          uint4 v[8], w[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = (i + lane) & 7;
            v[i] = *reinterpret_cast<const uint4*>(src + r * 128 + (((k8 & 7) ^ r) << 4));
          }
          // w[j] = channel j of the octet = v[(j - lane) & 7]
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = (lane & 1) ? v[(j + 7) & 7] : v[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (lane & 2) ? w[(j + 6) & 7] : w[j];
#pragma unroll
          for (int j = 0; j < 8; ++j) w[j] = (lane & 4) ? v[(j + 4) & 7] : v[j];
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            uint4 o;
            uint32_t* op = &o.x;
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const uint32_t* a0 = &w[2 * m].x;
              const uint32_t* a1 = &w[2 * m + 1].x;
              op[m] = __byte_perm(a0[r >> 1], a1[r >> 1], (r & 1) ? 0x7632 : 0x5410);
            }
            if (k8 * 8 + r < vn) vo[(size_t)(k8 * 8 + r) * (D / 8)] = o;
          }
        };
        // ---- drain: this thread's channel row of V^T, 32 pixel rows at a time
        const int nblk = (sp.rpp + 31) >> 5;     // (a last half block reads 16 stale columns: never used as K)
#pragma unroll 1
        for (int b = 0; b < nblk; ++b) {
          uint32_t u32[32];
          tmem_ld32(tlane + mt * D + b * 32, u32);
          tmem_ld_wait();
          uint8_t* dst = sm + (b >> 1) * C3_VT_CHUNK + c * 128;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            uint32_t* wp = &w.x;
#pragma unroll
            for (int h2 = 0; h2 < 4; ++h2) {
              const float lo = fmaxf(__uint_as_float(u32[8 * i + 2 * h2]) + bias_c, 0.f);
              const float hi = fmaxf(__uint_as_float(u32[8 * i + 2 * h2 + 1]) + bias_c, 0.f);
              const __nv_bfloat162 hh = __floats2bfloat162_rn(lo, hi);
              wp[h2] = *reinterpret_cast<const uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(dst + ((((b & 1) * 4 + i) ^ (c & 7)) << 4)) = w;
          }
        }
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 34);
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(vt_ready);
        if (vo) {   // all eight warps have drained: warp = groups ew, ew + 8, ...
          named_bar_sync(1, C2_EPI);
#pragma unroll 1
          for (int k8 = ew; k8 * 8 < vn; k8 += 8) store_group(k8);
        }
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 36);
        // ---- S rows of this pass: anchor a = 16 q + lane (lanes 0-15 of quarter q), columns 128 mt .. + 127
        mbar_wait(comb_done, pi & 1u);
        tc_fence_after();
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 35);
        const int a = lane < 16 ? q * 16 + lane : A;   // (M = 64 accumulator layout)
        if (q * 16 < A) {
#pragma unroll 1
          for (int b = 0; b < 4; ++b) {
            uint32_t u32[32];
            tmem_ld32(tlane + mt * 128 + b * 32, u32);
            tmem_ld_wait();
            if (a < A) {
              const size_t o = ((size_t)scene * A + a) * D + mt * 128 + b * 32;
              DDH_ASSERT(scene < B && o + 32 <= (size_t)B * A * D);
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u32[i]);
              if (pass > 0) {   // same thread wrote it in the previous pass: deterministic RMW
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 old = *reinterpret_cast<const float4*>(S32 + o + 4 * i);
                  v[4 * i] += old.x; v[4 * i + 1] += old.y; v[4 * i + 2] += old.z; v[4 * i + 3] += old.w;
                }
              }
              if (pass + 1 < passes) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  *reinterpret_cast<float4*>(S32 + o + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              } else if (S16) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  uint4 w;
                  uint32_t* wp = &w.x;
#pragma unroll
                  for (int h2 = 0; h2 < 4; ++h2) {
                    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[8 * i + 2 * h2], v[8 * i + 2 * h2 + 1]);
                    wp[h2] = *reinterpret_cast<const uint32_t*>(&hh);
                  }
                  *reinterpret_cast<uint4*>(S16 + o + 8 * i) = w;
                }
              }
            }
          }
        }
        C2_STAMP(sidx == 1 && pass == 0 && etid == 0, 40);
        // V^T / Wc (generic-proxy writes, read by the MMA) and the accumulators may be overwritten
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(passgo);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<NT * D>(tmem_base);
}

int tc_engine_init() {
  cudaError_t e;
  e = cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_conv2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_conv2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           227 * 1024);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(tc_conv3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           C_PIPE + D * 4 + TC_BAR_BYTES + 1024);
  return (int)e;
}

void launch_tc_gemm(const GemmParams& p0, const CUtensorMap& wmap, int n_total, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  GemmParams p = p0;
  p.n_blocks = n_total / D;
  const int n_tiles = ((p.M + TC_BM - 1) / TC_BM) * p.n_blocks;
  tc_gemm_kernel<<<n_tiles < num_sms ? n_tiles : num_sms, G_THREADS, G_SMEM, st>>>(p, wmap);
}

int tc_conv_smem_bytes(int A, int ent_per_anchor) {
  return C_PIPE + ((A * ent_per_anchor * 8 + 15) / 16) * 16 + D * 4 + TC_BAR_BYTES + 1024;
}

void launch_tc_conv(const GemmParams& p0, const CUtensorMap& wmap, int B, cudaStream_t st, int mode) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  GemmParams p = p0;
  p.M = num_sms;
  if (mode == 2 && p.n_anchor <= 64 && p.ent_per_anchor == 32)
    tc_conv3_kernel<<<B < num_sms ? B : num_sms, C3_THREADS, C_PIPE + D * 4 + TC_BAR_BYTES + 1024, st>>>(p, wmap, B);
  else if (mode >= 1)
    tc_conv2_kernel<false><<<B < num_sms ? B : num_sms, C2_THREADS, tc_conv_smem_bytes(p.n_anchor, p.ent_per_anchor), st>>>(p, wmap, B);
  else
    tc_conv_kernel<false><<<B, TC_THREADS, tc_conv_smem_bytes(p.n_anchor, p.ent_per_anchor), st>>>(p, wmap);
}

// value rows of the cross-scene list p.vrows[0 .. *p.n_vrows) -> p.vout (PlanReuse): the device-side count
// is split evenly over one CTA per SM; CTAs without rows leave at once
void launch_tc_convv(const GemmParams& p0, const CUtensorMap& wmap, cudaStream_t st) {
  int dev = 0, num_sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  if (num_sms <= 0) num_sms = 148;
  GemmParams p = p0;
  p.n_anchor = 0;
  p.ent_per_anchor = 0;
  tc_conv_kernel<true><<<num_sms, TC_THREADS, tc_conv_smem_bytes(0, 0), st>>>(p, wmap);
}

// fp32 engine: the same persistent conv with fp32 operands on the tensor core as 3xTF32 (p.bev /
// p.bev_lo: high / low plane of the NHWC fp32 map, wmap: fp32 [256][w_hi | w_lo], box {32, 256});
// fp32 combine on the CUDA cores, S written as fp32
void launch_tc_conv_tf32(const GemmParams& p0, const CUtensorMap& wmap, int B, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  GemmParams p = p0;
  p.M = num_sms;
  tc_conv2_kernel<true><<<B < num_sms ? B : num_sms, C2_THREADS, tc_conv_smem_bytes(p.n_anchor, p.ent_per_anchor), st>>>(p, wmap, B);
}

}  // namespace ddh
