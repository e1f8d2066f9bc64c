// Anchor-resident engine of the ddh planning head (sm_100a): the whole
// TrajectoryHead.forward_test (transfuser_model_v2.py:578-641) of one scene runs in ONE kernel on
// ONE 16-CTA thread-block cluster, and -- unlike the first resident engine (kernels_res.cu), which
// split every Linear over its output features and exchanged activations through L2 after every
// stage -- the decoder chain of an anchor never leaves the SM that owns it.
//
// Work split.  CTA `rank` owns anchors {rank, rank + 16} (A <= 32).  Every Linear of the chain is
//     out^T[f, n] = sum_k W[f, k] * x[n, k]       (tcgen05.mma, M = 128 weight rows, N = 16)
// with the FULL weight matrix streamed through a shared-memory ring by a dedicated TMA thread that
// free-runs ahead of the math (weights do not depend on data) and the owner's <= 2 activation
// rows as the B operand, rebuilt in shared memory by the epilogue of the previous stage: one
// thread per output feature reads its accumulator from TMEM, applies bias / ReLU / residual /
// LayerNorm (block reductions) / FiLM and writes the bf16 operand of the next stage.  No global
// memory, no cluster barrier and no exchange inside the chain.  What does cross CTAs:
//   * the sampling plan needs every anchor's points and attention weights: owners push them into
//     all 16 CTAs' shared memory (st.shared::cluster), every CTA then builds the identical plan;
//   * the on-demand value_proj conv (modules/blocks.py:68-76,114) runs as (128-row tile) x
//     (32-column group) tcgen05 tiles, one per CTA, gathered from the NHWC bf16 map with
//     cp.async; each CTA pushes its [A x 32] slice of the sampled features to the anchor owners;
//   * the step-invariant agent K|V and ego projections are computed once, feature-split, and
//     exchanged through L2; each CTA stages K|V of a layer in shared memory under the conv;
//   * NCHW callers: the BEV rows a conv call reads are converted on demand, dealt over the CTAs.
// Cluster-wide synchronisation is an mbarrier per CTA that one thread of every CTA arrives on
// remotely (release.cluster / acquire.cluster), ~10 times per forward.
//
// Numerics are those of the bf16 tensor engine (kernels_tc.cu): bf16 operands, fp32 accumulate,
// fp32 LayerNorm / softmax / embeddings / residuals / regression tail.
#include <stdlib.h>

#include "geom.cuh"
#include "kernels_res2.h"
#include "tc_ptx.cuh"

namespace ddh {
namespace {

constexpr int NT = 352;                      // 8 compute warps + ring TMA + MMA + conv-weight TMA
constexpr int NCT = 256;                     // compute threads
constexpr int NAL = 2;                       // anchors per CTA
constexpr int SLOT = 16 * 1024;              // weight ring slot: 128 rows x 64 bf16
constexpr int NSLOT = 5;
constexpr int RING = NSLOT * SLOT;
constexpr int NCC = D / (RES_CL / 2);        // conv output columns per CTA (32)
constexpr int A_TILE = 128 * 128;
constexpr int CNS = 4;                       // conv pipeline stages
constexpr int CSTAGE = A_TILE + NCC * 128;
constexpr int PIPE = CNS * CSTAGE;
constexpr int KC_CONV = 9 * (D / 64);        // 36 k-chunks: (tap, 64-channel chunk)
constexpr int BCH = 1024;                    // chain B operand: 8 rows x 128 B per k-chunk; rows 8..15 of
                                             // the N = 16 operand alias rows 0..7 (descriptor SBO = 0)
constexpr int BCH32 = 4096;                  // hoisted stage: 32 rows x 128 B
constexpr uint32_t ACC_CONV = 0, ACC_LIN = 32;
constexpr int TMEM_COLS = 256;
constexpr int VS_LD = NCC + 4;
constexpr int KS_LD = D + 4;                 // padded K rows: conflict-free 128-bit reads, lane = agent

// fixed region behind RING + PIPE
constexpr int F_CONSTS = 0;                  // R2Consts
constexpr int F_BOP = 4096;                  // main B operand (16 KiB) | also the conv drain staging
constexpr int F_BOP2 = F_BOP + 16384;        // cls-branch B operand (4 KiB)
constexpr int F_Q0 = F_BOP2 + 4096;          // float [NAL][256]
constexpr int F_X1 = F_Q0 + 2048;
constexpr int F_T0 = F_X1 + 2048;            // scratch rows (q, r2)
constexpr int F_SP = F_T0 + 2048;            // float [2 tiles][NAL][256] sampled-feature partials
constexpr int F_EGO = F_SP + 4096;           // float [L][256]
constexpr int F_ENT = F_EGO + 4096;          // EntPair [A*P*4] (<= 1024)
constexpr int F_UPIX = F_ENT + 8192;         // int [rcap] (<= 1024)
constexpr int F_BM = F_UPIX + 4096;          // uint [HW/32] pixel bitmap | int [HW/32] prefix
constexpr int F_AW = F_BM + 1024;            // float [L][A*P]
constexpr int F_PTS = F_AW + 4096;           // float [A*P*2] every anchor's current points
constexpr int F_OWN = F_PTS + 2048;          // own img [NAL][16] | own pts [NAL][16] | regraw [NAL][24] | modes [NAL][24]
constexpr int F_RED = F_OWN + 1024;          // block-reduction scratch
constexpr int F_MISC = F_RED + 512;          // conv bias slice [32] | logits [64] | ints [16]
constexpr int F_FIN = F_MISC + 512;          // rank 0: scores [32] | modes [A*3P] (<= 768)
constexpr int F_BAR = F_FIN + 3328;
constexpr int F_END = F_BAR + 256;
constexpr int SMEM_BYTES = RING + PIPE + F_END + 1024;
static_assert(sizeof(R2Consts) <= F_BOP - F_CONSTS, "R2Consts must fit its shared-memory slot");
static_assert(sizeof(R2Consts) % 16 == 0, "R2Consts is copied as uint4");
static_assert(128 * VS_LD * 4 <= 16384 + 4096, "drain staging must fit the B operand buffers");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct EntPair { int slot; float w; };

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_hw() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `cta`
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t raddr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(raddr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_release(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// D = f32, A = B = bf16, K-major, M = 128, N = n
__device__ __forceinline__ constexpr uint32_t idesc_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// K-major 128-byte-swizzled operand whose 8-row groups are `sbo` bytes apart
__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t saddr, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(sbo >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t a, b;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  v[0] = __uint_as_float(a);
  v[1] = __uint_as_float(b);
}

// byte offset of element (row n, column k) of a K-major, 128-byte-swizzled bf16 operand whose
// 64-wide k-chunks are `chunk_bytes` apart
__device__ __forceinline__ uint32_t sw_off(int n, int k, int chunk_bytes) {
  return (uint32_t)((k >> 6) * chunk_bytes + n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
}
__device__ __forceinline__ void bop_store(uint8_t* bop, int n, int k, float v) {
  *reinterpret_cast<__nv_bfloat16*>(bop + sw_off(n, k, BCH)) = __float2bfloat16_rn(v);
}

template <typename TI>
__device__ __forceinline__ float4 ld4_bev(const TI* p);
template <>
__device__ __forceinline__ float4 ld4_bev<float>(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
template <>
__device__ __forceinline__ float4 ld4_bev<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

// NCHW -> NHWC bf16 for one (row y, 32-pixel block): 256 channels x 32 pixels through a padded
// shared-memory tile (same scheme as bev_rows_to_nhwc_kernel); 256 compute threads
template <typename TI>
__device__ __forceinline__ void layout_item(const TI* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                            int HW, int px0, uint32_t* tile_u32, int tid) {
  constexpr int LDW = 129;
  constexpr int LDE = LDW * 2;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(tile_u32);
  const int px4 = tid & 7, cl = tid >> 3, lane = tid & 31, warp = tid >> 5;
  const TI* s = src + px0 + px4 * 4;
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ld4_bev<TI>(s + (size_t)(i * 32 + cl) * HW);
  named_bar_sync(1, NCT);   // previous tile fully written out
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + cl;
    tile[(px4 * 4 + 0) * LDE + c] = __float2bfloat16_rn(v[i].x);
    tile[(px4 * 4 + 1) * LDE + c] = __float2bfloat16_rn(v[i].y);
    tile[(px4 * 4 + 2) * LDE + c] = __float2bfloat16_rn(v[i].z);
    tile[(px4 * 4 + 3) * LDE + c] = __float2bfloat16_rn(v[i].w);
  }
  named_bar_sync(1, NCT);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)px0 * D);
  for (int px = warp; px < 32; px += 8) {
#pragma unroll
    for (int w = lane; w < 128; w += 32) d[(size_t)px * 128 + w] = tile_u32[px * LDW + w];
  }
}

__global__ void __launch_bounds__(NT, 1)
res2_forward_kernel(const R2Consts* __restrict__ gconsts, const ResCall call) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = (int)cluster_ctarank();
  const int scene = (int)cluster_id_x();
  uint8_t* pipe = sm + RING;
  const uint32_t pipe_addr = sm_addr + RING;
  uint8_t* fix = sm + RING + PIPE;
  const uint32_t fix_addr = sm_addr + RING + PIPE;
  const R2Consts& C = *reinterpret_cast<const R2Consts*>(fix + F_CONSTS);
  uint8_t* bop = fix + F_BOP;
  uint8_t* bop2 = fix + F_BOP2;
  float* q0_s = reinterpret_cast<float*>(fix + F_Q0);
  float* x1_s = reinterpret_cast<float*>(fix + F_X1);
  float* t0_s = reinterpret_cast<float*>(fix + F_T0);
  float* sp_s = reinterpret_cast<float*>(fix + F_SP);
  float* ego_s = reinterpret_cast<float*>(fix + F_EGO);
  EntPair* ent = reinterpret_cast<EntPair*>(fix + F_ENT);
  int* upix_s = reinterpret_cast<int*>(fix + F_UPIX);
  unsigned int* bm_s = reinterpret_cast<unsigned int*>(fix + F_BM);
  int* pre_s = reinterpret_cast<int*>(fix + F_BM + 512);
  float* aw_s = reinterpret_cast<float*>(fix + F_AW);
  float* pts_s = reinterpret_cast<float*>(fix + F_PTS);
  float* img_o = reinterpret_cast<float*>(fix + F_OWN);          // [NAL][16]
  float* pts_o = img_o + NAL * 16;                               // [NAL][16]
  float* raw_o = pts_o + NAL * 16;                               // [NAL][24]
  float* red_s = reinterpret_cast<float*>(fix + F_RED);          // 2 x [8][4]
  float* cbias_s = reinterpret_cast<float*>(fix + F_MISC);       // [32]
  float* logit_s = cbias_s + 32;                                 // [64]
  int* ints_s = reinterpret_cast<int*>(logit_s + 64);            // [16]
  unsigned long long* need_s = reinterpret_cast<unsigned long long*>(ints_s + 8);
  float* fin_scores = reinterpret_cast<float*>(fix + F_FIN);     // [32]
  float* fin_modes = fin_scores + 32;                            // [A*3P]
  const uint32_t bar = fix_addr + F_BAR;
  auto ring_full = [&](int s) { return bar + s * 8; };
  auto ring_empty = [&](int s) { return bar + (NSLOT + s) * 8; };
  auto conv_full = [&](int s) { return bar + (2 * NSLOT + s) * 8; };
  auto conv_empty = [&](int s) { return bar + (2 * NSLOT + CNS + s) * 8; };
  const uint32_t conv_acc = bar + (2 * NSLOT + 2 * CNS) * 8;
  const uint32_t acc_full = conv_acc + 8;
  const uint32_t b_ready = conv_acc + 16;
  const uint32_t conv_go = conv_acc + 24;
  const uint32_t cl_bar = conv_acc + 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(fix + F_BAR + (2 * NSLOT + 2 * CNS + 5) * 8);

  {  // constants -> shared memory
    const uint4* s = reinterpret_cast<const uint4*>(gconsts);
    uint4* d = reinterpret_cast<uint4*>(fix + F_CONSTS);
    for (int i = tid; i < (int)(sizeof(R2Consts) / 16); i += NT) d[i] = __ldg(s + i);
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), 1); }
    for (int s = 0; s < CNS; ++s) { mbar_init(conv_full(s), NCT + 1); mbar_init(conv_empty(s), 1); }
    mbar_init(conv_acc, 1);
    mbar_init(acc_full, 1);
    mbar_init(b_ready, NCT);
    mbar_init(conv_go, 1);
    mbar_init(cl_bar, RES_CL);
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  cluster_sync_hw();   // every CTA's barriers exist before anyone arrives remotely

  const int A = C.A, P = C.P, Na = C.Na, L = C.L, S = C.S, H = C.H, W = C.W;
  const int AP = A * P, HW = H * W;
  const int tile = rank / (RES_CL / 2), cgp = rank % (RES_CL / 2);

  // =============================================================== weight-ring TMA thread
  if (warp == 8) {
    if (lane == 0) {
      int seq = 0;
      for (int si = 0; si < C.n_stages; ++si) {
        const R2Stage stg = C.stages[si];
        if (stg.flags & R2F_CONV) continue;
        const int row0 = (stg.flags & R2F_RANKROWS) ? rank * (int)stg.rows : 0;
        for (int mt = 0; mt < (int)stg.mtiles; ++mt)
          for (int kc = 0; kc < (int)stg.kchunks; ++kc) {
            const int slot = seq % NSLOT, use = seq / NSLOT;
            if (use > 0) mbar_wait(ring_empty(slot), (uint32_t)((use - 1) & 1));
            mbar_arrive_expect_tx(ring_full(slot), (uint32_t)stg.rows * 128u);
            tma_load_2d(sm_addr + slot * SLOT, stg.map, ring_full(slot), kc * 64, row0 + mt * 128);
            ++seq;
          }
      }
    }
    __syncwarp();
  }
  // =============================================================== MMA thread
  else if (warp == 9) {
    if (lane == 0) {
      int seq = 0, cg = 0;
      uint32_t bpar = 0, gopar = 0;
      for (int si = 0; si < C.n_stages; ++si) {
        const R2Stage stg = C.stages[si];
        if (stg.flags & R2F_CONV) {
          mbar_wait(conv_go, gopar);
          gopar ^= 1u;
          const int nu = *reinterpret_cast<volatile int*>(ints_s + 4);
          const int passes = (nu + 255) / 256;
          constexpr uint32_t idesc = idesc_m128(NCC);
          for (int pass = 0; pass < passes; ++pass) {
            if (pass * 256 + tile * 128 >= nu) continue;
            for (int kc = 0; kc < KC_CONV; ++kc) {
              const int g = cg + kc, s = g % CNS;
              mbar_wait(conv_full(s), (uint32_t)((g / CNS) & 1));
              tc_fence_after();
              const uint32_t a_stage = pipe_addr + s * CSTAGE;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)
                umma_bf16(tmem + ACC_CONV, umma_desc_sw128(a_stage + k4 * 32),
                          umma_desc_sw128(a_stage + A_TILE + k4 * 32), idesc, 1u);
              umma_commit(conv_empty(s));
            }
            umma_commit(conv_acc);
            cg += KC_CONV;
          }
          continue;
        }
        if (stg.flags & R2F_WAITB) {
          mbar_wait(b_ready, bpar);
          bpar ^= 1u;
          tc_fence_after();
        }
        const bool n32 = (stg.flags & R2F_N32) != 0;
        const uint32_t ncol = n32 ? 32u : 16u;
        const uint32_t idesc = idesc_m128(ncol);
        const uint32_t b_addr = fix_addr + (stg.bsel ? F_BOP2 : F_BOP);
        const uint32_t bch = n32 ? BCH32 : BCH, sbo = n32 ? 1024u : 0u;
        for (int mt = 0; mt < (int)stg.mtiles; ++mt)
          for (int kc = 0; kc < (int)stg.kchunks; ++kc) {
            const int slot = seq % NSLOT;
            mbar_wait(ring_full(slot), (uint32_t)((seq / NSLOT) & 1));
            tc_fence_after();
            const uint32_t a_base = sm_addr + slot * SLOT;
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4)
              umma_bf16(tmem + stg.acc_col + mt * ncol, umma_desc_sw128(a_base + k4 * 32),
                        umma_desc_sw128_sbo(b_addr + kc * bch + k4 * 32, sbo), idesc, (kc | k4) ? 1u : 0u);
            umma_commit(ring_empty(slot));
            ++seq;
          }
        if (stg.flags & R2F_COMMIT) umma_commit(acc_full);
      }
    }
    __syncwarp();
  }
  // =============================================================== conv-weight TMA thread
  else if (warp == 10) {
    if (lane == 0) {
      int cg = 0;
      uint32_t gopar = 0;
      for (int si = 0; si < C.n_stages; ++si) {
        const R2Stage stg = C.stages[si];
        if (!(stg.flags & R2F_CONV)) continue;
        mbar_wait(conv_go, gopar);
        gopar ^= 1u;
        const int nu = *reinterpret_cast<volatile int*>(ints_s + 4);
        const int passes = (nu + 255) / 256;
        for (int pass = 0; pass < passes; ++pass) {
          if (pass * 256 + tile * 128 >= nu) continue;
          for (int kc = 0; kc < KC_CONV; ++kc) {
            const int g = cg + kc, s = g % CNS;
            mbar_wait(conv_empty(s), (uint32_t)(((g / CNS) & 1) ^ 1));
            mbar_arrive_expect_tx(conv_full(s), NCC * 128);
            tma_load_2d(pipe_addr + s * CSTAGE + A_TILE, stg.map, conv_full(s), kc * 64, cgp * NCC);
          }
          cg += KC_CONV;
        }
      }
    }
    __syncwarp();
  }
  // =============================================================== compute warps
  else {
    const int quad = warp & 3, half = warp >> 2;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int own_a[NAL] = {rank, rank + RES_CL};
    const bool own_v[NAL] = {rank < A, rank + RES_CL < A};
    uint32_t acc_par = 0, cl_par = 0, conv_par = 0;
    int cg = 0, rsel = 0, dbg_i = 0;
    unsigned long long done_rows = 0ull;
    const __nv_bfloat16* bevn =
        call.bev_nhwc_bf16 ? reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)scene * HW * D
                           : C.bev_nhwc + (size_t)scene * HW * D;
    float* kvg = C.kv + (size_t)scene * L * Na * 2 * D;
    float* egog = C.egov + (size_t)scene * L * D;

    auto mark = [&](int label) {
      if (call.dbg && scene == 0 && rank == 0 && tid == 0 && dbg_i < 1000)
        call.dbg[dbg_i++] = ((long long)label << 48) | (clock64() & 0xFFFFFFFFFFFFll);
    };
    auto bsync = [&]() { named_bar_sync(1, NCT); };
    // all compute threads of all CTAs; release/acquire at cluster scope (covers the remote
    // shared-memory pushes and the global memory exchanged through L2)
    auto csync = [&]() {
      mark(104);
      bsync();
      if (tid < RES_CL) mbar_arrive_remote_release(mapa(cl_bar, (uint32_t)tid));
      mbar_wait_acq_cluster(cl_bar, cl_par);
      cl_par ^= 1u;
      mark(105);
    };
    // B operand(s) written by this thread -> visible to the tensor core -> count me in
    auto b_done = [&]() {
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(b_ready);
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_full, acc_par);
      acc_par ^= 1u;
      tc_fence_after();
    };
    // sums over the 256 features (= compute threads) of two values per thread
    auto block_sum2 = [&](float& a, float& b) {
      a = warp_sum(a);
      b = warp_sum(b);
      float* r = red_s + rsel * 16;
      rsel ^= 1;
      if (lane == 0) { r[warp * 2] = a; r[warp * 2 + 1] = b; }
      bsync();
      const float4 r0 = *reinterpret_cast<const float4*>(r), r1 = *reinterpret_cast<const float4*>(r + 4),
                   r2 = *reinterpret_cast<const float4*>(r + 8), r3 = *reinterpret_cast<const float4*>(r + 12);
      a = ((r0.x + r0.z) + (r1.x + r1.z)) + ((r2.x + r2.z) + (r3.x + r3.z));
      b = ((r0.y + r0.w) + (r1.y + r1.w)) + ((r2.y + r2.w) + (r3.y + r3.w));
    };
    // LayerNorm over the features of both anchors; this thread holds feature `tid`
    auto ln_feat = [&](float (&v)[NAL], float g, float bt) {
      float s0 = v[0], s1 = v[1];
      block_sum2(s0, s1);
      const float m0 = s0 * (1.0f / D), m1 = s1 * (1.0f / D);
      const float d0 = v[0] - m0, d1 = v[1] - m1;
      float q0 = d0 * d0, q1 = d1 * d1;
      block_sum2(q0, q1);
      v[0] = d0 * (1.0f / sqrtf(q0 * (1.0f / D) + LN_EPS)) * g + bt;
      v[1] = d1 * (1.0f / sqrtf(q1 * (1.0f / D) + LN_EPS)) * g + bt;
    };
    auto acc2 = [&](int mt, int group, float (&v)[NAL]) {   // accumulator of feature tile `mt`
      tmem_ld2(tlane + ACC_LIN + (uint32_t)(group * 32 + mt * 16), v);
    };

    // ---- img = sqrt(ac) * norm_odo(anchors) + sqrt(1-ac) * noise   (:591-597); every CTA derives
    // the step-0 points of every anchor itself, owners keep the diffused sample of their anchors
    for (int i = tid; i < AP * 2; i += NCT) {
      const float a = __ldg(C.anchors + i);
      const float nv = (i & 1) ? norm_y(a) : norm_x(a);
      const float im = __fadd_rn(__fmul_rn(C.sa_tr, nv), __fmul_rn(C.sb_tr, __ldg(call.noise + (size_t)scene * AP * 2 + i)));
      const float v = fminf(fmaxf(im, -1.0f), 1.0f);
      const float pt = (i & 1) ? denorm_y(v) : denorm_x(v);
      pts_s[i] = pt;
      const int a_i = i / (2 * P), r = i - a_i * 2 * P;
      if ((a_i & (RES_CL - 1)) == rank) {
        img_o[(a_i >> 4) * 16 + r] = im;
        pts_o[(a_i >> 4) * 16 + r] = pt;
      }
    }
    mark(1);

    // ================= hoisted agent K|V and ego vectors (step-invariant, :316-327,355-364),
    // feature-split over the cluster, exchanged through L2
    {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int a = warp + 8 * r;
        if (a < Na + 1) {
          const float* src = a < Na ? call.agents + ((size_t)scene * Na + a) * D : call.ego + (size_t)scene * D;
          const float4 u0 = __ldg(reinterpret_cast<const float4*>(src + lane * 4));
          const float4 u1 = __ldg(reinterpret_cast<const float4*>(src + 128 + lane * 4));
          const float vv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            const int k = h2 * 128 + lane * 4;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(vv[4 * h2 + 0], vv[4 * h2 + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(vv[4 * h2 + 2], vv[4 * h2 + 3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(bop + sw_off(a, k, BCH32)) = u;
          }
        }
      }
      b_done();
      wait_acc();
      const int rows = 3 * D / RES_CL;   // 48 features of [K | V | ego] per CTA
      if (half == 0 && quad * 32 < rows) {
        const int f = quad * 32 + lane;
        const int gfeat = rank * rows + f;
        for (int l = 0; l < L; ++l) {
          uint32_t u[32];
          tmem_ld32(tlane + ACC_LIN + 32 * l, u);
          tmem_ld_wait();
          if (f < rows) {
            const float bias = __ldg(C.layer[l].b_kvego + gfeat);
            float* kvl = kvg + (size_t)l * Na * 2 * D;
#pragma unroll
            for (int a = 0; a < 32; ++a) {
              const float v = __uint_as_float(u[a]) + bias;
              if (gfeat < 2 * D) { if (a < Na) kvl[(size_t)a * 2 * D + gfeat] = v; }
              else if (a == Na) egog[(size_t)l * D + gfeat - 2 * D] = v;
            }
          }
        }
      }
      tc_fence_before();
      csync();
      for (int i = tid; i < L * D; i += NCT) ego_s[i] = __ldcg(egog + i);
    }
    mark(2);

    for (int si = 0; si < S; ++si) {
      const bool last_step = (si == S - 1);
      // ============ sine embedding of the owner's anchors (blocks.py:22-40) -> B operand (K = 64 P)
      {
        const float two_pi = 6.283185307179586f;
        for (int i = tid; i < NAL * P * 64; i += NCT) {
          const int n = i / (P * 64), r = i - n * P * 64;
          const int p = r >> 6, j = r & 63, hf = j >> 5, ii = j & 31;
          float val = 0.f;
          if (own_v[n]) {
            const float v = hf ? pts_o[n * 16 + p * 2 + 0] : pts_o[n * 16 + p * 2 + 1];   // (pos_y | pos_x)
            const float arg = __fdiv_rn(__fmul_rn(v, two_pi), __ldg(C.dim_t + ii));
            val = (ii & 1) ? cosf(arg) : sinf(arg);
          }
          bop_store(bop, n, r, val);
        }
        b_done();
      }
      mark(10);
      // ============ plan_anchor_encoder (:459-462): Linear(512->256)+ReLU+LN, Linear(256->256)
      {
        const float bias = __ldg(C.b_enc0 + tid), g = __ldg(C.enc_ln_g + tid), bt = __ldg(C.enc_ln_b + tid);
        wait_acc();
        float v[NAL];
        acc2(half, 0, v);
        v[0] = fmaxf(v[0] + bias, 0.f);
        v[1] = fmaxf(v[1] + bias, 0.f);
        ln_feat(v, g, bt);
        bop_store(bop, 0, tid, v[0]);
        bop_store(bop, 1, tid, v[1]);
        b_done();
      }
      mark(11);
      {
        const float bias = __ldg(C.b_enc3 + tid);
        wait_acc();
        float v[NAL];
        acc2(half, 0, v);
        q0_s[tid] = v[0] + bias;
        q0_s[D + tid] = v[1] + bias;
        if (call.dbg) {
#pragma unroll
          for (int n = 0; n < NAL; ++n)
            if (own_v[n]) C.tap_q0[((size_t)scene * A + own_a[n]) * D + tid] = v[n] + bias;
        }
        bsync();
        // attention weights of every layer (blocks.py:98-100): softmax_p(q0 . Wa^T + ba)
        const int ndot = NAL * L * P;            // (n, l, p)
        for (int d = warp; d < ndot; d += 8) {
          const int n = d / (L * P), r = d - n * L * P, l = r / P, p = r - l * P;
          const float* wr = C.layer[l].attw_w + (size_t)p * D;
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(q0_s[n * D + lane + 32 * i], __ldg(wr + lane + 32 * i), s);
          s = warp_sum(s);
          if (lane == 0) logit_s[d] = s + __ldg(C.layer[l].attw_b + p);
        }
        bsync();
        if (tid < NAL * L * P) {
          const int n = tid / (L * P), r = tid - n * L * P, l = r / P, p = r - l * P;
          if (own_v[n]) {
            const float* lg = logit_s + (n * L + l) * P;
            float mx = lg[0];
            for (int o = 1; o < P; ++o) mx = fmaxf(mx, lg[o]);
            float den = 0.f;
            for (int o = 0; o < P; ++o) den += expf(lg[o] - mx);
            const float w = expf(lg[p] - mx) / den;
            const uint32_t dst = fix_addr + F_AW + (uint32_t)((l * AP + own_a[n] * P + p) * 4);
            for (int c = 0; c < RES_CL; ++c) st_cluster_f32(mapa(dst, (uint32_t)c), w);
          }
        }
        csync();
      }
      mark(12);

      for (int l = 0; l < L; ++l) {
        const ResLayerC& LC = C.layer[l];
        const bool last_layer = (l == L - 1);
        const bool want_cls = last_layer && last_step;
        const bool do_ddim = last_layer && !last_step;
        // ============ sampling plan (blocks.py:98-125), identically in every CTA
        int nu;
        unsigned long long todo;
        {
          const int nwords = HW / 32;
          if (tid < nwords) bm_s[tid] = 0u;
          if (tid < NCC) cbias_s[tid] = __ldg(LC.b_conv + cgp * NCC + tid);
          if (tid == 0) *need_s = 0ull;
          Corners c;
          float a_w = 0.f;
          if (tid < AP) {
            c = corners_of(pts_s[tid * 2 + 0], pts_s[tid * 2 + 1], H, W, C.oc);
            a_w = aw_s[l * AP + tid];
          }
          bsync();
          if (tid < AP) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (c.pix[k] >= 0) atomicOr(bm_s + (c.pix[k] >> 5), 1u << (c.pix[k] & 31));
          }
          bsync();
          // ordered compaction (pixel order == memory order of the NHWC map)
          if (tid < 128) {
            const unsigned int bits = tid < nwords ? bm_s[tid] : 0u;
            const int cnt = __popc(bits);
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const int t = __shfl_up_sync(0xffffffffu, incl, o);
              if (lane >= o) incl += t;
            }
            if (lane == 31) ints_s[warp] = incl;
            named_bar_sync(3, 128);
            int base = incl - cnt;
            for (int w = 0; w < warp; ++w) base += ints_s[w];
            if (tid == 127) ints_s[4] = base + cnt;
            if (tid < nwords) pre_s[tid] = base;
            if (bits) {
              const int y = (tid * 32) / W, x0 = tid * 32 - y * W;
              unsigned int rest = bits;
              while (rest) {
                const int b = __ffs((int)rest) - 1;
                rest &= rest - 1;
                upix_s[base++] = (y << 16) | (x0 + b);
              }
              atomicOr(need_s, (y > 0) ? (7ull << (y - 1)) : 3ull);
            }
          }
          bsync();
          nu = ints_s[4];
          const unsigned long long need_all = (H >= 64) ? *need_s : (*need_s & ((1ull << H) - 1ull));
          todo = call.bev_nhwc_bf16 ? 0ull : (need_all & ~done_rows);
          done_rows |= need_all;
          if (tid < AP) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              EntPair ep;
              ep.slot = -1;
              ep.w = 0.f;
              if (c.pix[k] >= 0) {
                const int wd = c.pix[k] >> 5;
                ep.slot = pre_s[wd] + __popc(bm_s[wd] & ((1u << (c.pix[k] & 31)) - 1u));
                ep.w = c.w[k] * a_w;
              }
              ent[tid * 4 + k] = ep;
            }
          }
        }
        mark(20);
        // ============ on-demand BEV layout: the rows this conv call reads, not converted yet
        if (todo) {
          uint32_t* tile_u32 = reinterpret_cast<uint32_t*>(pipe);
          __nv_bfloat16* dst = C.bev_nhwc + (size_t)scene * HW * D;
          const int tpr = W / 32;
          unsigned long long rest = todo;
          int idx = 0;
          while (rest) {
            const int y = __ffsll((long long)rest) - 1;
            rest &= rest - 1;
            for (int xt = 0; xt < tpr; ++xt, ++idx) {
              if (idx % RES_CL != rank) continue;
              const int px0 = y * W + xt * 32;
              if (call.bev_dtype == 0)
                layout_item<float>(reinterpret_cast<const float*>(call.bev) + (size_t)scene * D * HW, dst, HW,
                                   px0, tile_u32, tid);
              else
                layout_item<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)scene * D * HW,
                                           dst, HW, px0, tile_u32, tid);
            }
          }
          csync();
          mark(21);
        } else {
          bsync();
        }
        // ============ value_proj conv at the unique pixels + bilinear/attention combine
        {
          if (tid == 0) mbar_arrive(conv_go);
          const int passes = (nu + 255) / 256;
          const int a_c = tid >> 3, cqd = tid & 7;          // combine: (anchor, 4-column group)
          float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
          float* Vs = reinterpret_cast<float*>(bop);
          for (int pass = 0; pass < passes; ++pass) {
            const int row_base = pass * 256 + tile * 128;
            if (row_base >= nu) continue;
            const int rows_valid = min(128, nu - row_base);
            const int j = tid & 7, rb = tid >> 3;
            int rowoff[4];
            uint32_t vmask[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = rb + 32 * i;
              rowoff[i] = 0;
              vmask[i] = 0;
              if (r < rows_valid) {
                const int yx = upix_s[row_base + r];
                const int y = yx >> 16, x = yx & 0xffff;
                rowoff[i] = (y * W + x) * D + j * 8;
                const uint32_t xm = (x > 0 ? 1u : 0u) | 2u | (x + 1 < W ? 4u : 0u);
                vmask[i] = (y > 0 ? xm : 0u) | (xm << 3) | (y + 1 < H ? (xm << 6) : 0u);
              }
            }
            const uint32_t dst_base = rb * 128 + ((j ^ (rb & 7)) << 4);
            for (int kc = 0; kc < KC_CONV; ++kc) {
              const int g = cg + kc, s = g % CNS;
              mbar_wait(conv_empty(s), (uint32_t)(((g / CNS) & 1) ^ 1));
              const uint32_t a_dst = pipe_addr + s * CSTAGE + dst_base;
              const int tap = kc >> 2;
              const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
              const int tapoff = (dy * W + dx) * D + (kc & 3) * 64;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const bool ok = (vmask[i] >> tap) & 1u;
                const int off = ok ? rowoff[i] + tapoff : 0;
                cp_async16(a_dst + i * 4096, bevn + off, ok ? 16u : 0u);
              }
              if (kc == 0 && warp < 4) {   // accumulators start at the conv bias
                uint32_t u[32];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  const uint4 t4 = *reinterpret_cast<const uint4*>(cbias_s + 4 * q);
                  u[4 * q + 0] = t4.x; u[4 * q + 1] = t4.y; u[4 * q + 2] = t4.z; u[4 * q + 3] = t4.w;
                }
                tmem_st32(tlane + ACC_CONV, u);
                tmem_st_wait();
                tc_fence_before();
              }
              cp_async_mbar_arrive_noinc(conv_full(s));
            }
            cg += KC_CONV;
            mark(125);
            mbar_wait(conv_acc, conv_par);
            conv_par ^= 1u;
            tc_fence_after();
            mark(126);
            if (warp < 4) {   // drain (ReLU) into the staging area
              float* vrow = Vs + (size_t)(warp * 32 + lane) * VS_LD;
              uint32_t u0[32];
              tmem_ld32(tlane + ACC_CONV, u0);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 8; ++q)
                *reinterpret_cast<float4*>(vrow + 4 * q) = make_float4(
                    fmaxf(__uint_as_float(u0[4 * q]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 1]), 0.f),
                    fmaxf(__uint_as_float(u0[4 * q + 2]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 3]), 0.f));
              tc_fence_before();
            }
            bsync();
            if (a_c < A) {
              const EntPair* ea = ent + a_c * P * 4;
#pragma unroll 8
              for (int k = 0; k < P * 4; ++k) {
                const EntPair e = ea[k];
                const int rr = e.slot - row_base;
                if (rr >= 0 && rr < rows_valid) {
                  const float4 v = *reinterpret_cast<const float4*>(Vs + (size_t)rr * VS_LD + cqd * 4);
                  sacc.x = fmaf(e.w, v.x, sacc.x); sacc.y = fmaf(e.w, v.y, sacc.y);
                  sacc.z = fmaf(e.w, v.z, sacc.z); sacc.w = fmaf(e.w, v.w, sacc.w);
                }
              }
            }
            bsync();
          }
          // this CTA's [A x 32] slice of the sampled features -> the anchor owners
          if (a_c < A) {
            const uint32_t dst = fix_addr + F_SP +
                                 (uint32_t)(((tile * NAL + (a_c >> 4)) * D + cgp * NCC + cqd * 4) * 4);
            st_cluster_v4(mapa(dst, (uint32_t)(a_c & (RES_CL - 1))), sacc);
          }
          // stage this layer's agent K|V (fp32) in the idle conv pipeline buffers
          {
            const float* kvl = kvg + (size_t)l * Na * 2 * D;
            const int n16 = Na * 128;
            for (int i = tid; i < n16; i += NCT) {
              const int jrow = i >> 7, u = i & 127;
              const uint32_t dst = (u < 64) ? pipe_addr + (uint32_t)((jrow * KS_LD + u * 4) * 4)
                                            : pipe_addr + (uint32_t)(32 * KS_LD * 4 + (jrow * D + (u - 64) * 4) * 4);
              cp_async16(dst, kvl + (size_t)jrow * 2 * D + u * 4, 16u);
            }
            cp_async_commit();
          }
          csync();
        }
        mark(22);
        // ============ output_proj + residual (blocks.py:127-129): x1 = S.Wo + b + q0
        {
          bop_store(bop, 0, tid, sp_s[tid] + sp_s[NAL * D + tid]);
          bop_store(bop, 1, tid, sp_s[D + tid] + sp_s[NAL * D + D + tid]);
          b_done();
          const float bias = __ldg(LC.b_bev_out + tid);
          wait_acc();
          float v[NAL];
          acc2(half, 0, v);
          v[0] += bias + q0_s[tid];
          v[1] += bias + q0_s[D + tid];
          x1_s[tid] = v[0];
          x1_s[D + tid] = v[1];
          if (call.dbg) {
#pragma unroll
            for (int n = 0; n < NAL; ++n)
              if (own_v[n]) C.tap_x1[((size_t)scene * A + own_a[n]) * D + tid] = v[n];
          }
          bop_store(bop, 0, tid, v[0]);
          bop_store(bop, 1, tid, v[1]);
          b_done();
        }
        mark(23);
        // ============ cross_agent_attention (:316-321,355-357): q projection, softmax(qK^T)V
        {
          const float bias = __ldg(LC.b_q + tid);
          wait_acc();
          float v[NAL];
          acc2(half, 0, v);
          const float scale = 0.17677669529663687f;   // 1/sqrt(32)
          t0_s[tid] = (v[0] + bias) * scale;
          t0_s[D + tid] = (v[1] + bias) * scale;
          cp_async_wait_all();
          bsync();
          const float* Ks = reinterpret_cast<const float*>(pipe);
          const float* Vv = Ks + 32 * KS_LD;
          const int hc = warp * 32;                   // one head per warp
#pragma unroll
          for (int n = 0; n < NAL; ++n) {
            float s = -INFINITY;
            if (lane < Na) {
              s = 0.f;
              const float* kr = Ks + lane * KS_LD + hc;
              const float* qr = t0_s + n * D + hc;
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 kk = *reinterpret_cast<const float4*>(kr + c4 * 4);
                const float4 qq = *reinterpret_cast<const float4*>(qr + c4 * 4);
                s = fmaf(qq.x, kk.x, s); s = fmaf(qq.y, kk.y, s); s = fmaf(qq.z, kk.z, s); s = fmaf(qq.w, kk.w, s);
              }
            }
            float mx = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float e = (lane < Na) ? expf(s - mx) : 0.f;
            const float pj = e / warp_sum(e);
            float acc = 0.f;
            for (int jj = 0; jj < Na; ++jj) acc = fmaf(__shfl_sync(0xffffffffu, pj, jj), Vv[jj * D + hc + lane], acc);
            bop_store(bop, n, hc + lane, acc);
          }
          b_done();
        }
        mark(24);
        // ============ attention out_proj + residual, norm1, + ego, norm2   (:355-364)
        {
          const float bias = __ldg(LC.b_attn_out + tid);
          const float g1 = __ldg(LC.norm1_g + tid), b1 = __ldg(LC.norm1_b + tid);
          const float g2 = __ldg(LC.norm2_g + tid), b2 = __ldg(LC.norm2_b + tid);
          const float eg = ego_s[l * D + tid];
          wait_acc();
          float v[NAL];
          acc2(half, 0, v);
          v[0] += bias + x1_s[tid];
          v[1] += bias + x1_s[D + tid];
          ln_feat(v, g1, b1);
          v[0] += eg;
          v[1] += eg;
          ln_feat(v, g2, b2);
          bop_store(bop, 0, tid, v[0]);
          bop_store(bop, 1, tid, v[1]);
          b_done();
        }
        mark(25);
        // ============ FFN up: h = relu(x2.W1 + b)   (:366-368)
        {
          const int ntile = C.F / 256;   // feature tiles per thread (tiles half, half+2, ...)
          float bias[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) bias[i] = (i < ntile) ? __ldg(LC.b_ffn0 + (half + 2 * i) * 128 + quad * 32 + lane) : 0.f;
          wait_acc();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (i < ntile) {
              const int mt = half + 2 * i;
              float v[NAL];
              acc2(mt, 0, v);
              const int f = mt * 128 + quad * 32 + lane;
              bop_store(bop, 0, f, fmaxf(v[0] + bias[i], 0.f));
              bop_store(bop, 1, f, fmaxf(v[1] + bias[i], 0.f));
            }
          }
          b_done();
        }
        mark(26);
        // ============ FFN down, norm3, time FiLM   (:368-373)
        {
          const float bias = __ldg(LC.b_ffn2 + tid);
          const float g3 = __ldg(LC.norm3_g + tid), b3 = __ldg(LC.norm3_b + tid);
          const float* film = C.film + ((size_t)si * L + l) * 2 * D;
          const float sc = __ldg(film + tid), sh = __ldg(film + D + tid);
          wait_acc();
          float v[NAL];
          acc2(half, 0, v);
          v[0] += bias;
          v[1] += bias;
          ln_feat(v, g3, b3);
          v[0] = v[0] * (1.0f + sc) + sh;
          v[1] = v[1] * (1.0f + sc) + sh;
          bop_store(bop, 0, tid, v[0]);
          bop_store(bop, 1, tid, v[1]);
          b_done();
        }
        mark(27);
        // ============ reg / cls hidden 1   (:221-231)
        {
          const float bias_r = __ldg(LC.b_reg0 + tid);
          float bias_c = 0.f, gc = 0.f, bc = 0.f;
          if (want_cls) { bias_c = __ldg(LC.b_cls0 + tid); gc = __ldg(LC.cls_ln2_g + tid); bc = __ldg(LC.cls_ln2_b + tid); }
          wait_acc();
          float v[NAL], c[NAL];
          acc2(half, 0, v);
          if (want_cls) acc2(half, 1, c);
          bop_store(bop, 0, tid, fmaxf(v[0] + bias_r, 0.f));
          bop_store(bop, 1, tid, fmaxf(v[1] + bias_r, 0.f));
          if (want_cls) {
            c[0] = fmaxf(c[0] + bias_c, 0.f);
            c[1] = fmaxf(c[1] + bias_c, 0.f);
            ln_feat(c, gc, bc);
            bop_store(bop2, 0, tid, c[0]);
            bop_store(bop2, 1, tid, c[1]);
          }
          b_done();
        }
        mark(28);
        // ============ reg / cls hidden 2, regression head (256 -> 3P, fp32), cls score
        {
          const float bias_r = __ldg(LC.b_reg2 + tid);
          float bias_c = 0.f, gc = 0.f, bc = 0.f, w6 = 0.f;
          if (want_cls) {
            bias_c = __ldg(LC.b_cls3 + tid); gc = __ldg(LC.cls_ln5_g + tid); bc = __ldg(LC.cls_ln5_b + tid);
            w6 = __ldg(LC.cls6_w + tid);
          }
          // regression-head rows of this warp: outputs c = warp, warp + 8, warp + 16 (3P = 24)
          float w4[3][8];
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            const int c = warp + 8 * ci;
            if (c < 3 * P) {
              const float4 w0 = __ldg(reinterpret_cast<const float4*>(LC.reg4_w + (size_t)c * D + lane * 4));
              const float4 w1 = __ldg(reinterpret_cast<const float4*>(LC.reg4_w + (size_t)c * D + 128 + lane * 4));
              w4[ci][0] = w0.x; w4[ci][1] = w0.y; w4[ci][2] = w0.z; w4[ci][3] = w0.w;
              w4[ci][4] = w1.x; w4[ci][5] = w1.y; w4[ci][6] = w1.z; w4[ci][7] = w1.w;
            }
          }
          wait_acc();
          float v[NAL], c[NAL];
          acc2(half, 0, v);
          if (want_cls) acc2(half, 1, c);
          t0_s[tid] = fmaxf(v[0] + bias_r, 0.f);
          t0_s[D + tid] = fmaxf(v[1] + bias_r, 0.f);
          float score[NAL] = {0.f, 0.f};
          if (want_cls) {   // scores = LN(c2).w6 + b6   (:221-224)
            c[0] = fmaxf(c[0] + bias_c, 0.f);
            c[1] = fmaxf(c[1] + bias_c, 0.f);
            ln_feat(c, gc, bc);
            score[0] = c[0] * w6;
            score[1] = c[1] * w6;
            block_sum2(score[0], score[1]);
            const float b6 = __ldg(LC.cls6_b);
            score[0] += b6;
            score[1] += b6;
          } else {
            bsync();
          }
#pragma unroll
          for (int ci = 0; ci < 3; ++ci) {
            const int cc = warp + 8 * ci;
            if (cc < 3 * P) {
#pragma unroll
              for (int n = 0; n < NAL; ++n) {
                float x[8];
                load8(t0_s + n * D, lane, x);
                float s = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(x[i], w4[ci][i], s);
                s = warp_sum(s);
                if (lane == 0) raw_o[n * 24 + cc] = s + __ldg(LC.reg4_b + cc);
              }
            }
          }
          bsync();
          mark(29);
          // reg[..., :2] += points; heading = tanh(.)*pi; next points; DDIM update (:378-380,424,632-636)
          if (tid < NAL * 3 * P) {
            const int n = tid / (3 * P), cidx = tid - n * 3 * P;
            const int p = cidx / 3, comp = cidx - p * 3;
            if (own_v[n]) {
              const int a = own_a[n];
              const float mine = raw_o[n * 24 + cidx];
              if (call.dbg) C.tap_regraw[((size_t)scene * A + a) * 3 * P + cidx] = mine;
              float out, nxt = 0.f;
              if (comp < 2) {
                const int pi = n * 16 + p * 2 + comp;
                out = __fadd_rn(mine, pts_o[pi]);
                nxt = out;
                if (do_ddim) {
                  const DdimCoef dc = C.dc[si];
                  const float x0 = comp ? norm_y(out) : norm_x(out);
                  const float sample = img_o[pi];
                  const float eps = __fdiv_rn(__fsub_rn(sample, __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
                  const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
                  const float im = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
                  img_o[pi] = im;
                  const float vc = fminf(fmaxf(im, -1.0f), 1.0f);
                  nxt = comp ? denorm_y(vc) : denorm_x(vc);
                }
                pts_o[pi] = nxt;
                if (!want_cls) {   // every CTA's next plan needs them
                  const uint32_t dst = fix_addr + F_PTS + (uint32_t)(((a * P + p) * 2 + comp) * 4);
                  for (int cta = 0; cta < RES_CL; ++cta) st_cluster_f32(mapa(dst, (uint32_t)cta), nxt);
                }
              } else {
                out = __fmul_rn(tanhf(mine), 3.14159265358979323846f);
              }
              if (want_cls) {
                if (call.out_modes) call.out_modes[((size_t)scene * A + a) * 3 * P + cidx] = out;
                st_cluster_f32(mapa(fix_addr + F_FIN + (uint32_t)((32 + a * 3 * P + cidx) * 4), 0u), out);
                if (cidx == 0) {
                  if (call.out_scores) call.out_scores[(size_t)scene * A + a] = score[n];
                  st_cluster_f32(mapa(fix_addr + F_FIN + (uint32_t)(a * 4), 0u), score[n]);
                }
              }
            }
          }
          if (!last_layer || want_cls) csync(); else bsync();
        }
        mark(31);
      }
    }
    // ================= mode = argmax(cls) (first maximum wins), trajectory = reg[mode]  (:637-640)
    if (rank == 0) {
      int best = 0;
      float bv = fin_scores[0];
      for (int a = 1; a < A; ++a) {
        const float v = fin_scores[a];
        if (v > bv) { bv = v; best = a; }
      }
      if (tid == 0 && call.out_mode_idx) call.out_mode_idx[scene] = best;
      if (call.out_traj)
        for (int i = tid; i < 3 * P; i += NCT) call.out_traj[(size_t)scene * 3 * P + i] = fin_modes[best * 3 * P + i];
    }
    mark(99);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_hw();   // nobody leaves while a peer may still push into its shared memory
  if (warp == 8) tmem_dealloc<TMEM_COLS>(tmem);
}

}  // namespace

int res2_smem_bytes() { return SMEM_BYTES; }

static bool g_res2_ready = false;

static void res2_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int B, cudaStream_t st) {
  cfg = {};
  cfg.gridDim = dim3(RES_CL * B);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RES_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}

int res2_engine_init() {
  if (g_res2_ready) return 0;
  cudaError_t e = cudaFuncSetAttribute(res2_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }
  e = cudaFuncSetAttribute(res2_forward_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { cudaGetLastError(); return 2; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  res2_cfg(cfg, attr, 1, nullptr);
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, res2_forward_kernel, &cfg);
  if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 3; }
  g_res2_ready = true;
  return 0;
}

int launch_res2_forward(const R2Consts* consts_dev, const ResCall& call, int B, cudaStream_t st) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  res2_cfg(cfg, attr, B, st);
  cudaError_t e = cudaLaunchKernelEx(&cfg, res2_forward_kernel, consts_dev, call);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ddh
