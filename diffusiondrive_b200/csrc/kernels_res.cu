// Resident engine of the ddh planning head (sm_100a): the whole TrajectoryHead.forward_test
// (transfuser_model_v2.py:578-641) of one scene runs in ONE kernel on ONE 16-CTA thread-block
// cluster, so that a batch-1 forward costs one launch instead of ~54 dependent ones.
//
// Work split.  Every Linear of the decoder chain is split over its OUTPUT FEATURES: CTA `rank`
// owns features [rank*rows, (rank+1)*rows).  A stage is
//     out^T[f, a] = sum_k W[f, k] * x[a, k]        (tcgen05.mma, M = 128 lanes of which `rows`
//                                                   are real features, N = 32 anchor columns)
// with the weight slice as the A operand (streamed by TMA through a 3-slot ring, always 2-3
// stages ahead of use: weights do not depend on data) and the <= 32 activation rows as the B
// operand, which every CTA rebuilds in its own shared memory from the full activation rows:
// stages exchange their [A x slice] outputs through global memory (L2) and a hardware cluster
// barrier (barrier.cluster, release/acquire).  LayerNorm / FiLM / the ego add are applied by the
// consumer while it converts the rows to the swizzled bf16 operand layout.  The on-demand
// value_proj conv (modules/blocks.py:68-76,114) runs as (128-row tile) x (32-column group)
// tcgen05 tiles, one per CTA, gathered from the NHWC bf16 map with cp.async; the sampling plan
// (blocks.py:98-125), DDIM arithmetic and point bookkeeping are computed redundantly by every
// CTA so that they need no exchange at all.
//
// Numerics are those of the bf16 tensor engine (kernels_tc.cu): bf16 operands, fp32 accumulate,
// fp32 LayerNorm / softmax / embeddings / residuals / regression tail.
#include <stdlib.h>

#include "geom.cuh"
#include "kernels_res.h"
#include "tc_ptx.cuh"

namespace ddh {
namespace {

constexpr int RT = 256;                      // threads per CTA
constexpr int NCC = D / (RES_CL / 2);        // conv output columns per CTA (32)
constexpr int SLOT = 32 * 1024;              // weight ring slot
constexpr int NSLOT = 3;
constexpr int RING = NSLOT * SLOT;
constexpr int A_TILE = 128 * 128;            // 128 rows x 64 bf16
constexpr int CNS = 4;                       // conv pipeline stages
constexpr int CSTAGE = A_TILE + NCC * 128;
constexpr int PIPE = CNS * CSTAGE;           // also: B operands, staging areas (see offsets)
constexpr int BT = 32 * 128;                 // B operand: 32 rows x 128 B per 64-wide k-chunk
constexpr int VS_LD = NCC + 4;
constexpr int KC_CONV = 9 * (D / 64);        // 36 k-chunks: (tap, 64-channel chunk)
constexpr uint32_t ACC_CONV = 0, ACC0 = 64, ACC1 = 96;   // TMEM columns
constexpr int TMEM_COLS = 128;

// scratch inside the PIPE region (only while no conv pipeline is live)
constexpr int P_B0 = 0, P_B1 = 16384;        // B operand tiles
constexpr int P_ATT = 32768;                 // attention scratch (qs | ks | vs)
constexpr int P_MODES = 49152;               // final modes [A*P*3]

// fixed region behind RING + PIPE
constexpr int F_CONSTS = 0;
constexpr int F_TABLE = 6144;                // ushort [H*W] (<= 4096 pixels)
constexpr int F_ENT = F_TABLE + 8192;        // EntPair [A*P*4] (<= 1024)
constexpr int F_UPIX = F_ENT + 8192;         // int [rcap] (<= 1024)
constexpr int F_AW = F_UPIX + 4096;          // float [A*P]
constexpr int F_PTS = F_AW + 1024;           // float [A*P*2]
constexpr int F_IMG = F_PTS + 2048;          // float [A*P*2]
constexpr int F_MISC = F_IMG + 2048;         // conv bias slice [64] | scores [32] | ints [32]
constexpr int F_BAR = F_MISC + 1024;
constexpr int F_END = F_BAR + 256;
constexpr int SMEM_BYTES = RING + PIPE + F_END + 1024;
static_assert(sizeof(ResConsts) <= F_TABLE - F_CONSTS, "ResConsts must fit its shared-memory slot");
static_assert(sizeof(ResConsts) % 16 == 0, "ResConsts is copied as uint4");
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
// all threads of all CTAs of the cluster; release/acquire at cluster scope (covers the global
// memory the stages exchange through)
__device__ __forceinline__ void cluster_sync_all() {
  // generic-proxy accesses to shared memory before the barrier are ordered before async-proxy
  // (TMA / tcgen05) accesses after it: the staging areas are reused across proxies
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

// D = f32, A = B = bf16, K-major, M = 128, N = n
__device__ __forceinline__ constexpr uint32_t idesc_m128(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void ldcg8(const float* p, int lane, float (&o)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p + lane * 4));
  const float4 b = __ldcg(reinterpret_cast<const float4*>(p + 128 + lane * 4));
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// byte offset of element (row a, column k) of a K-major, 128-byte-swizzled bf16 operand whose
// 64-wide k-chunks are `chunk_bytes` apart
__device__ __forceinline__ uint32_t sw_off(int a, int k, int chunk_bytes) {
  return (uint32_t)((k >> 6) * chunk_bytes + a * 128 + ((((k & 63) >> 3) ^ (a & 7)) << 4) + (k & 7) * 2);
}

// lane's 8 values of row a (columns lane*4+{0..3} and 128+lane*4+{0..3}) -> bf16 B operand
__device__ __forceinline__ void bt_store8(uint8_t* bt, int a, int lane, const float (&v)[8]) {
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int k = h2 * 128 + lane * 4;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[4 * h2 + 0], v[4 * h2 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[4 * h2 + 2], v[4 * h2 + 3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(bt + sw_off(a, k, BT)) = u;
  }
}

// fp32 rows (one warp per row, rows warp, warp+8, ...) -> prologue -> bf16 B operand; all row
// loads of a warp are in flight before any is consumed
template <typename RowPtr, typename Pro>
__device__ __forceinline__ void bt_rows_f32(uint8_t* bt, int nrows, int warp, int lane, RowPtr rowptr,
                                            Pro pro) {
  float v[4][8];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int a = warp + 8 * r;
    if (a < nrows) ldcg8(rowptr(a), lane, v[r]);
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int a = warp + 8 * r;
    if (a < nrows) {
      pro(a, v[r]);
      bt_store8(bt, a, lane, v[r]);
    }
  }
}

// bf16 rows [nrows][K] -> B operand (16-byte units, up to 12 in flight per thread)
__device__ __forceinline__ void bt_rows_bf16(uint8_t* bt, const __nv_bfloat16* src, int nrows, int K,
                                             int tid) {
  const int upr = K >> 3, n = nrows * upr;
  for (int base = tid; base < n; base += RT * 12) {
    uint4 t[12];
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const int i = base + u * RT;
      if (i < n) t[u] = __ldcg(reinterpret_cast<const uint4*>(src) + i);
    }
#pragma unroll
    for (int u = 0; u < 12; ++u) {
      const int i = base + u * RT;
      if (i < n) {
        const int a = i / upr, k = (i - a * upr) << 3;
        *reinterpret_cast<uint4*>(bt + sw_off(a, k, BT)) = t[u];
      }
    }
  }
}

// accumulator tile [128 feature lanes x 32 anchor columns] -> fn(feature f, anchor a, value)
template <typename Fn>
__device__ __forceinline__ void epi_tile(uint32_t tmem_acc, int rows, int nA, int warp, int lane,
                                         Fn fn) {
  if (warp < 4 && warp * 32 < rows) {
    uint32_t u[32];
    tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16), u);
    tmem_ld_wait();
    const int f = warp * 32 + lane;
    if (f < rows) {
#pragma unroll
      for (int a = 0; a < 32; ++a)
        if (a < nA) fn(f, a, __uint_as_float(u[a]));
    }
  }
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
// shared-memory tile (same scheme as bev_rows_to_nhwc_kernel)
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
  __syncthreads();   // previous tile fully written out
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = i * 32 + cl;
    tile[(px4 * 4 + 0) * LDE + c] = __float2bfloat16_rn(v[i].x);
    tile[(px4 * 4 + 1) * LDE + c] = __float2bfloat16_rn(v[i].y);
    tile[(px4 * 4 + 2) * LDE + c] = __float2bfloat16_rn(v[i].z);
    tile[(px4 * 4 + 3) * LDE + c] = __float2bfloat16_rn(v[i].w);
  }
  __syncthreads();
  uint32_t* d = reinterpret_cast<uint32_t*>(dst + (size_t)px0 * D);
  for (int px = warp; px < 32; px += 8) {
#pragma unroll
    for (int w = lane; w < 128; w += 32) d[(size_t)px * 128 + w] = tile_u32[px * LDW + w];
  }
}

__global__ void __launch_bounds__(RT, 1)
res_forward_kernel(const ResConsts* __restrict__ gconsts, const ResCall call) {
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
  const ResConsts& C = *reinterpret_cast<const ResConsts*>(fix + F_CONSTS);
  unsigned short* table = reinterpret_cast<unsigned short*>(fix + F_TABLE);
  EntPair* ent = reinterpret_cast<EntPair*>(fix + F_ENT);
  int* upix_s = reinterpret_cast<int*>(fix + F_UPIX);
  float* aw_s = reinterpret_cast<float*>(fix + F_AW);
  float* pts_s = reinterpret_cast<float*>(fix + F_PTS);
  float* img_s = reinterpret_cast<float*>(fix + F_IMG);
  float* cbias_s = reinterpret_cast<float*>(fix + F_MISC);          // [64]
  float* scores_s = cbias_s + 64;                                   // [32]
  int* ints_s = reinterpret_cast<int*>(scores_s + 32);              // [32]
  unsigned long long* need_s = reinterpret_cast<unsigned long long*>(ints_s + 16);
  const uint32_t bar = sm_addr + RING + PIPE + F_BAR;
  auto ring_full = [&](int s) { return bar + s * 8; };
  auto ring_empty = [&](int s) { return bar + (NSLOT + s) * 8; };
  auto conv_full = [&](int s) { return bar + (2 * NSLOT + s) * 8; };
  auto conv_empty = [&](int s) { return bar + (2 * NSLOT + CNS + s) * 8; };
  const uint32_t conv_acc = bar + (2 * NSLOT + 2 * CNS) * 8;
  const uint32_t lin_acc = conv_acc + 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(fix + F_BAR + (2 * NSLOT + 2 * CNS + 2) * 8);

  {  // constants -> shared memory
    const uint4* s = reinterpret_cast<const uint4*>(gconsts);
    uint4* d = reinterpret_cast<uint4*>(fix + F_CONSTS);
    for (int i = tid; i < (int)(sizeof(ResConsts) / 16); i += RT) d[i] = __ldg(s + i);
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), 1); }
    for (int s = 0; s < CNS; ++s) { mbar_init(conv_full(s), 128 + 1); mbar_init(conv_empty(s), 1); }
    mbar_init(conv_acc, 1);
    mbar_init(lin_acc, 1);
    fence_barrier_init();
  }
  if (warp == 5) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int A = C.A, P = C.P, Na = C.Na, F = C.F, L = C.L, S = C.S, H = C.H, W = C.W;
  const int AP = A * P, HW = H * W;
  const bool is_ctrl = (warp == 4 && lane == 0);
  // per-scene views of the exchange buffers: the shared-memory copy of the constants is
  // re-based to this cluster's scene once, so that every use is a plain shared-memory load
  if (tid == 0) {
    ResConsts& Cw = *reinterpret_cast<ResConsts*>(fix + F_CONSTS);
    const size_t sAD = (size_t)scene * A * D;
    Cw.emb16 += (size_t)scene * A * 64 * P; Cw.o16 += sAD; Cw.h16 += (size_t)scene * A * F;
    Cw.r1_16 += sAD; Cw.e1 += sAD; Cw.q0 += sAD; Cw.spart += sAD * Cw.tiles_max; Cw.x1 += sAD;
    Cw.y2 += sAD; Cw.y3 += sAD; Cw.c1 += sAD; Cw.r2 += sAD; Cw.c2 += sAD;
    Cw.regraw += (size_t)scene * A * 3 * P; Cw.kv += (size_t)scene * L * Na * 2 * D;
    Cw.egov += (size_t)scene * L * D; Cw.bev_nhwc += (size_t)scene * HW * D;
  }
  __syncthreads();
#define emb16 C.emb16
#define o16 C.o16
#define h16 C.h16
#define r1_16 C.r1_16
#define e1 C.e1
#define q0 C.q0
#define spart C.spart
#define x1 C.x1
#define y2 C.y2
#define y3 C.y3
#define c1 C.c1
#define r2 C.r2
#define c2 C.c2
#define regraw C.regraw
#define kvbuf C.kv
#define egov C.egov
  const __nv_bfloat16* bevn =
      call.bev_nhwc_bf16 ? reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)scene * HW * D
                         : C.bev_nhwc;

  // ---- running state
  int item_idx = 0;                 // all threads: next entry of C.items (stage order)
  int pf_item = 0, pf_seq = 0;      // control thread: weight prefetcher
  int cons_seq = 0;                 // control thread: weight consumer
  uint32_t lin_par = 0;             // parity of lin_acc
  int cg = 0;                       // conv k-chunks issued so far (pipeline phase)
  uint32_t conv_par = 0;
  unsigned long long done_rows = 0ull;
  int dbg_i = 0;
  // timeline (debug builds of the caller pass call.dbg): (label << 48) | clock of thread 0 of
  // cluster 0 / rank 0
  auto mark = [&](int label) {
    if (call.dbg && scene == 0 && rank == 0 && tid == 0 && dbg_i < 1000)
      call.dbg[dbg_i++] = ((long long)label << 48) | (clock64() & 0xFFFFFFFFFFFFll);
  };
  auto csync = [&]() {
    mark(104);
    cluster_sync_all();
  };

  auto prefetch = [&](int upto_seq) {   // control thread only
    while (pf_seq <= upto_seq && pf_item < C.n_items) {
      const ResItem it = C.items[pf_item];
      if (rank * (int)it.rows < it.n_total) {
        const int slot = pf_seq % NSLOT, use = pf_seq / NSLOT;
        if (use > 0) mbar_wait(ring_empty(slot), (uint32_t)((use - 1) & 1));
        mbar_arrive_expect_tx(ring_full(slot), (uint32_t)it.rows * it.kchunks * 128u);
        for (int kc = 0; kc < (int)it.kchunks; ++kc)
          tma_load_2d(sm_addr + slot * SLOT + kc * it.rows * 128, it.map, ring_full(slot), kc * 64,
                      rank * (int)it.rows);
        ++pf_seq;
      }
      ++pf_item;
    }
  };
  // control thread: all MMAs of one weight item against the B operand at b_addr
  auto mma_item = [&](const ResItem& it, uint32_t b_addr, uint32_t acc_col) {
    const int slot = cons_seq % NSLOT;
    mbar_wait(ring_full(slot), (uint32_t)((cons_seq / NSLOT) & 1));
    tc_fence_after();
    const uint32_t a_base = sm_addr + slot * SLOT;
    constexpr uint32_t idesc = idesc_m128(32);
    for (int kc = 0; kc < (int)it.kchunks; ++kc) {
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4)
        umma_bf16(tmem + acc_col, umma_desc_sw128(a_base + kc * it.rows * 128 + k4 * 32),
                  umma_desc_sw128(b_addr + kc * BT + k4 * 32), idesc, (kc | k4) ? 1u : 0u);
    }
    umma_commit(ring_empty(slot));
    ++cons_seq;
  };
  // B operand(s) written -> MMAs issued -> accumulators complete (all threads return together)
  auto run_mma1 = [&](const ResItem& it, uint32_t b_addr) {
    mark(101);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    mark(102);
    if (is_ctrl) {
      mma_item(it, b_addr, ACC0);
      umma_commit(lin_acc);
      prefetch(cons_seq + NSLOT - 1);
    }
  };
  auto wait_acc = [&]() {
    mbar_wait(lin_acc, lin_par);
    lin_par ^= 1u;
    tc_fence_after();
    mark(103);
  };

  // ---- img = sqrt(ac) * norm_odo(anchors) + sqrt(1-ac) * noise   (:591-597)
  for (int i = tid; i < AP * 2; i += RT) {
    const float a = __ldg(C.anchors + i);
    const float nv = (i & 1) ? norm_y(a) : norm_x(a);
    img_s[i] = __fadd_rn(__fmul_rn(C.sa_tr, nv), __fmul_rn(C.sb_tr, __ldg(call.noise + (size_t)scene * AP * 2 + i)));
  }
  if (is_ctrl) prefetch(NSLOT - 1);
  mark(1);

  // ================= hoisted agent K|V and ego vectors (step-invariant, :316-327,355-364)
  {
    uint8_t* bt = pipe + P_B0;
    bt_rows_f32(bt, Na + 1, warp, lane,
                [&](int a) {
                  return a < Na ? call.agents + ((size_t)scene * Na + a) * D : call.ego + (size_t)scene * D;
                },
                [&](int, float (&)[8]) {});
    for (int l = 0; l < L; ++l) {
      const ResItem it = C.items[item_idx++];
      run_mma1(it, pipe_addr + P_B0);
      const int f = warp * 32 + lane;
      const int gfeat = rank * (int)it.rows + f;
      const float bias = (f < (int)it.rows) ? __ldg(C.layer[l].b_kvego + gfeat) : 0.f;
      wait_acc();
      float* kvl = kvbuf + (size_t)l * Na * 2 * D;
      float* egl = egov + (size_t)l * D;
      epi_tile(tmem + ACC0, it.rows, Na + 1, warp, lane, [&](int, int a, float v) {
        if (gfeat < 2 * D) { if (a < Na) kvl[(size_t)a * 2 * D + gfeat] = v + bias; }
        else if (a == Na) egl[gfeat - 2 * D] = v + bias;
      });
      tc_fence_before();
      __syncthreads();
    }
  }
  mark(2);

  for (int si = 0; si < S; ++si) {
    const bool last_step = (si == S - 1);
    // ============ clamp + denorm_odo (:601-602) and sine embedding (blocks.py:22-40)
    for (int i = tid; i < AP * 2; i += RT) {
      const float v = fminf(fmaxf(img_s[i], -1.0f), 1.0f);
      pts_s[i] = (i & 1) ? denorm_y(v) : denorm_x(v);
    }
    __syncthreads();
    {
      const int total = AP * 64, per = (total + RES_CL - 1) / RES_CL;
      const int i0 = rank * per, i1 = min(total, i0 + per);
      const float two_pi = 6.283185307179586f;
      for (int i = i0 + tid; i < i1; i += RT) {
        const int j = i & 63, ap = i >> 6;            // feature j of pose ap (= a*P + p)
        const int half = j >> 5, ii = j & 31;
        const float v = half ? pts_s[ap * 2 + 0] : pts_s[ap * 2 + 1];   // (pos_y | pos_x)
        const float arg = __fdiv_rn(__fmul_rn(v, two_pi), __ldg(C.dim_t + ii));
        emb16[i] = __float2bfloat16_rn((ii & 1) ? cosf(arg) : sinf(arg));
      }
    }
    csync();
    mark(10);
    // ============ plan_anchor_encoder (:459-462): Linear(512->256)+ReLU+LN, Linear(256->256)
    {
      const ResItem it = C.items[item_idx++];
      bt_rows_bf16(pipe + P_B0, emb16, A, 64 * P, tid);
      run_mma1(it, pipe_addr + P_B0);
      const int n = rank * (int)it.rows + warp * 32 + lane;
      const float bias = (warp * 32 + lane < (int)it.rows) ? __ldg(C.b_enc0 + n) : 0.f;
      wait_acc();
      epi_tile(tmem + ACC0, it.rows, A, warp, lane,
               [&](int, int a, float v) { e1[(size_t)a * D + n] = fmaxf(v + bias, 0.f); });
      tc_fence_before();
      csync();
    }
    mark(11);
    {
      const ResItem it = C.items[item_idx++];
      bt_rows_f32(pipe + P_B0, A, warp, lane, [&](int a) { return e1 + (size_t)a * D; },
                  [&](int, float (&v)[8]) { layer_norm_row(v, C.enc_ln_g, C.enc_ln_b, lane); });
      run_mma1(it, pipe_addr + P_B0);
      const int n = rank * (int)it.rows + warp * 32 + lane;
      const float bias = (warp * 32 + lane < (int)it.rows) ? __ldg(C.b_enc3 + n) : 0.f;
      wait_acc();
      epi_tile(tmem + ACC0, it.rows, A, warp, lane,
               [&](int, int a, float v) { q0[(size_t)a * D + n] = v + bias; });
      tc_fence_before();
      csync();
    }
    mark(12);

    for (int l = 0; l < L; ++l) {
      const ResLayerC& LC = C.layer[l];
      const bool last_layer = (l == L - 1);
      const bool want_cls = last_layer && last_step;
      const bool do_ddim = last_layer && !last_step;
      // ============ sampling plan (blocks.py:98-125), redundantly in every CTA
      int nu;
      unsigned long long todo;
      {
        float* attw_s = reinterpret_cast<float*>(pipe);   // [P][256] staged weights
        for (int i = tid; i < HW / 2; i += RT) reinterpret_cast<uint32_t*>(table)[i] = 0u;
        for (int i = tid; i < P * D / 4; i += RT)
          reinterpret_cast<float4*>(attw_s)[i] = __ldg(reinterpret_cast<const float4*>(LC.attw_w) + i);
        if (tid < NCC) cbias_s[tid] = __ldg(LC.b_conv + (rank % (RES_CL / 2)) * NCC + tid);
        if (tid == 0) *need_s = 0ull;
        float qv[4][8];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int a = warp + 8 * r;
          if (a < A) {
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[r][i] = __ldcg(q0 + (size_t)a * D + lane + 32 * i);
          }
        }
        __syncthreads();
        mark(110);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int a = warp + 8 * r;
          if (a < A) {
            float logit[8];
#pragma unroll
            for (int o = 0; o < 8; ++o) {
              float s = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) s = fmaf(qv[r][i], attw_s[o * D + lane + 32 * i], s);
              logit[o] = warp_sum(s) + __ldg(LC.attw_b + o);
            }
            float mx = logit[0];
#pragma unroll
            for (int o = 1; o < 8; ++o) mx = fmaxf(mx, logit[o]);
            float e[8], den = 0.f;
#pragma unroll
            for (int o = 0; o < 8; ++o) { e[o] = expf(logit[o] - mx); den += e[o]; }
            if (lane < 8) {
              float mine = e[0];
#pragma unroll
              for (int o = 1; o < 8; ++o) if (lane == o) mine = e[o];
              aw_s[a * P + lane] = mine / den;
            }
          }
        }
        // mark the in-bounds bilinear corners
        for (int e = tid; e < AP; e += RT) {
          const Corners c = corners_of(pts_s[e * 2 + 0], pts_s[e * 2 + 1], H, W, C.oc);
#pragma unroll
          for (int k = 0; k < 4; ++k) if (c.pix[k] >= 0) table[c.pix[k]] = 1;
        }
        __syncthreads();
        mark(111);
        // ordered compaction (pixel order == memory order of the NHWC map)
        const int ipt = (HW + RT - 1) / RT;
        const int beg = tid * ipt, end = min(HW, beg + ipt);
        int cnt = 0;
        for (int i = beg; i < end; ++i) cnt += table[i] ? 1 : 0;
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += t;
        }
        if (lane == 31) ints_s[warp] = incl;
        __syncthreads();
        int base = incl - cnt;
        for (int w = 0; w < warp; ++w) base += ints_s[w];
        if (tid == RT - 1) ints_s[8] = base + cnt;
        unsigned long long need = 0ull;
        for (int i = beg; i < end; ++i) {
          if (table[i]) {
            const int yy = i / W;
            upix_s[base] = (yy << 16) | (i - yy * W);
            table[i] = (unsigned short)(base + 1);
            ++base;
            need |= (yy > 0) ? (7ull << (yy - 1)) : 3ull;
          }
        }
        if (need) atomicOr(need_s, need);
        __syncthreads();
        mark(112);
        nu = ints_s[8];
        const unsigned long long need_all = (H >= 64) ? *need_s : (*need_s & ((1ull << H) - 1ull));
        todo = call.bev_nhwc_bf16 ? 0ull : (need_all & ~done_rows);
        done_rows |= need_all;
        for (int e = tid; e < AP; e += RT) {
          const Corners c = corners_of(pts_s[e * 2 + 0], pts_s[e * 2 + 1], H, W, C.oc);
          const float a_w = aw_s[e];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            EntPair ep;
            ep.slot = c.pix[k] >= 0 ? (int)table[c.pix[k]] - 1 : -1;
            ep.w = c.pix[k] >= 0 ? c.w[k] * a_w : 0.f;
            ent[e * 4 + k] = ep;
          }
        }
        __syncthreads();
      }
      mark(20);
      // ============ on-demand BEV layout: the rows this conv call reads, not converted yet
      if (todo) {
        uint32_t* tile_u32 = reinterpret_cast<uint32_t*>(pipe);
        __nv_bfloat16* dst = C.bev_nhwc;
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
      }
      // ============ value_proj conv at the unique pixels + bilinear/attention combine
      {
        const int tile = rank / (RES_CL / 2), cgp = rank % (RES_CL / 2);
        const int passes = (nu + 255) / 256;
        for (int pass = 0; pass < passes; ++pass) {
          const int row_base = pass * 256 + tile * 128;
          if (row_base < nu) {
            const int rows_valid = min(128, nu - row_base);
            if (warp < 4) {
              const int j = tid & 7, rb = tid >> 3;
              const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16) + ACC_CONV;
              int rowoff[8];
              uint32_t vmask[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = rb + 16 * i;
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
                if ((kc & 7) == 0) mark(120 + (kc >> 3));
                const uint32_t a_dst = pipe_addr + s * CSTAGE + dst_base;
                const int tap = kc >> 2;
                const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                const int tapoff = (dy * W + dx) * D + (kc & 3) * 64;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const bool ok = (vmask[i] >> tap) & 1u;
                  const int off = ok ? rowoff[i] + tapoff : 0;
                  cp_async16(a_dst + i * 2048, bevn + off, ok ? 16u : 0u);
                }
                if (kc == 0) {   // accumulators start at the conv bias
                  uint32_t u[32];
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const uint4 t4 = *reinterpret_cast<const uint4*>(cbias_s + 4 * q);
                    u[4 * q + 0] = t4.x; u[4 * q + 1] = t4.y; u[4 * q + 2] = t4.z; u[4 * q + 3] = t4.w;
                  }
                  tmem_st32(trow, u);
                  tmem_st_wait();
                  tc_fence_before();
                }
                cp_async_mbar_arrive_noinc(conv_full(s));
              }
              // drain (ReLU) into the staging area, combine this tile's rows
              mark(125);
              mbar_wait(conv_acc, conv_par);
              tc_fence_after();
              mark(126);
              float* Vs = reinterpret_cast<float*>(pipe);
              {
                float* vrow = Vs + (size_t)(warp * 32 + lane) * VS_LD;
                uint32_t u0[32];
                tmem_ld32(trow, u0);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  *reinterpret_cast<float4*>(vrow + 4 * q) = make_float4(
                      fmaxf(__uint_as_float(u0[4 * q]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 1]), 0.f),
                      fmaxf(__uint_as_float(u0[4 * q + 2]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 3]), 0.f));
              }
              tc_fence_before();
              named_bar_sync(1, 128);
              mark(127);
              const int cqd = tid & 7, ag = tid >> 3;
              float* sp = spart + (size_t)(pass * 2 + tile) * A * D;
              for (int a = ag; a < A; a += 16) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                const EntPair* ea = ent + a * P * 4;
#pragma unroll 8
                for (int k = 0; k < P * 4; ++k) {
                  const EntPair e = ea[k];
                  const int rr = e.slot - row_base;
                  if (rr >= 0 && rr < rows_valid) {
                    const float4 v = *reinterpret_cast<const float4*>(Vs + (size_t)rr * VS_LD + cqd * 4);
                    acc.x = fmaf(e.w, v.x, acc.x); acc.y = fmaf(e.w, v.y, acc.y);
                    acc.z = fmaf(e.w, v.z, acc.z); acc.w = fmaf(e.w, v.w, acc.w);
                  }
                }
                *reinterpret_cast<float4*>(sp + (size_t)a * D + cgp * NCC + cqd * 4) = acc;
              }
            } else if (warp == 4) {
              if (lane == 0) {
                for (int kc = 0; kc < KC_CONV; ++kc) {
                  const int g = cg + kc, s = g % CNS;
                  mbar_wait(conv_empty(s), (uint32_t)(((g / CNS) & 1) ^ 1));
                  mbar_arrive_expect_tx(conv_full(s), NCC * 128);
                  tma_load_2d(pipe_addr + s * CSTAGE + A_TILE, LC.conv_map, conv_full(s), kc * 64, cgp * NCC);
                }
              }
              __syncwarp();
            } else if (warp == 5) {
              if (lane == 0) {
                constexpr uint32_t idesc = idesc_m128(NCC);
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
              }
              __syncwarp();
            }
            cg += KC_CONV;
            conv_par ^= 1u;
          }
          fence_proxy_async();
          __syncthreads();
        }
      }
      csync();
      mark(22);
      // ============ output_proj + residual (blocks.py:127-129): x1 = S.Wo + b + q0
      {
        const ResItem it = C.items[item_idx++];
        const int parts = (nu + 127) / 128;
        {
          float v[4][8];
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
          for (int t = 0; t < parts; ++t) {
            float u[4][8];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int a = warp + 8 * r;
              if (a < A) ldcg8(spart + ((size_t)t * A + a) * D, lane, u[r]);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int a = warp + 8 * r;
              if (a < A)
#pragma unroll
                for (int i = 0; i < 8; ++i) v[r][i] += u[r][i];
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int a = warp + 8 * r;
            if (a < A) bt_store8(pipe + P_B0, a, lane, v[r]);
          }
        }
        run_mma1(it, pipe_addr + P_B0);
        const int f = warp * 32 + lane;
        const int n = rank * (int)it.rows + f;
        const bool fv = f < (int)it.rows;
        const float bias = fv ? __ldg(LC.b_bev_out + n) : 0.f;
        float res[32];
        {
          const float* rp = q0 + n;
#pragma unroll
          for (int a = 0; a < 32; ++a) res[a] = (fv && a < A) ? __ldcg(rp + (size_t)a * D) : 0.f;
        }
        wait_acc();
        {
          float* op = x1 + n;
          epi_tile(tmem + ACC0, it.rows, A, warp, lane,
                   [&](int, int a, float v) { op[(size_t)a * D] = v + bias + res[a]; });
        }
        tc_fence_before();
        csync();
      }
      mark(23);
      // ============ cross_agent_attention (:316-321,355-357): q projection + softmax(qK^T)V, one
      // head per CTA (ranks < heads)
      {
        const ResItem it = C.items[item_idx++];
        const bool part = rank * (int)it.rows < it.n_total;
        if (part) {
          float* qs = reinterpret_cast<float*>(pipe + P_ATT);   // [32][32]
          float* ks = qs + 32 * 32;                             // [32][33]
          float* vs = ks + 32 * 33;                             // [32][32]
          const float* kvl = kvbuf + (size_t)l * Na * 2 * D;
          const int hc = rank * 32;
          float kt[4], vt[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = tid + u * RT, j = i >> 5, c = i & 31;
            kt[u] = vt[u] = 0.f;
            if (j < Na) {
              kt[u] = __ldcg(kvl + (size_t)j * 2 * D + hc + c);
              vt[u] = __ldcg(kvl + (size_t)j * 2 * D + D + hc + c);
            }
          }
          bt_rows_f32(pipe + P_B0, A, warp, lane, [&](int a) { return x1 + (size_t)a * D; },
                      [&](int, float (&)[8]) {});
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = tid + u * RT, j = i >> 5, c = i & 31;
            if (j < Na) { ks[j * 33 + c] = kt[u]; vs[j * 32 + c] = vt[u]; }
          }
          run_mma1(it, pipe_addr + P_B0);
          const float bias = (warp == 0) ? __ldg(LC.b_q + hc + lane) : 0.f;
          wait_acc();
          const float scale = 0.17677669529663687f;   // 1/sqrt(32)
          epi_tile(tmem + ACC0, 32, A, warp, lane,
                   [&](int f, int a, float v) { qs[a * 32 + f] = (v + bias) * scale; });
          tc_fence_before();
          __syncthreads();
          for (int a = warp; a < A; a += 8) {
            float s = -INFINITY;
            if (lane < Na) {
              s = 0.f;
#pragma unroll
              for (int c = 0; c < 32; ++c) s = fmaf(qs[a * 32 + c], ks[lane * 33 + c], s);
            }
            float mx = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float e = (lane < Na) ? expf(s - mx) : 0.f;
            const float pj = e / warp_sum(e);
            float acc = 0.f;
            for (int j = 0; j < Na; ++j) acc = fmaf(__shfl_sync(0xffffffffu, pj, j), vs[j * 32 + lane], acc);
            o16[(size_t)a * D + hc + lane] = __float2bfloat16_rn(acc);
          }
        }
        csync();
      }
      mark(24);
      // ============ attention out_proj + residual: y2 = o.Wo + b + x1   (pre-norm1)
      {
        const ResItem it = C.items[item_idx++];
        bt_rows_bf16(pipe + P_B0, o16, A, D, tid);
        run_mma1(it, pipe_addr + P_B0);
        const int n = rank * (int)it.rows + warp * 32 + lane;
        const bool fv = warp * 32 + lane < (int)it.rows;
        const float bias = fv ? __ldg(LC.b_attn_out + n) : 0.f;
        float res[32];
        {
          const float* rp = x1 + n;
#pragma unroll
          for (int a = 0; a < 32; ++a) res[a] = (fv && a < A) ? __ldcg(rp + (size_t)a * D) : 0.f;
        }
        wait_acc();
        {
          float* op = y2 + n;
          epi_tile(tmem + ACC0, it.rows, A, warp, lane,
                   [&](int, int a, float v) { op[(size_t)a * D] = v + bias + res[a]; });
        }
        tc_fence_before();
        csync();
      }
      mark(25);
      // ============ FFN up: h = relu(x2.W1 + b), x2 = LN2(LN1(y2) + ego)   (:358-368)
      {
        const ResItem it = C.items[item_idx++];
        const float* egl = egov + (size_t)l * D;
        float eg[8];
        ldcg8(egl, lane, eg);
        bt_rows_f32(pipe + P_B0, A, warp, lane, [&](int a) { return y2 + (size_t)a * D; },
                    [&](int, float (&v)[8]) {
                      layer_norm_row(v, LC.norm1_g, LC.norm1_b, lane);
#pragma unroll
                      for (int i = 0; i < 8; ++i) v[i] += eg[i];
                      layer_norm_row(v, LC.norm2_g, LC.norm2_b, lane);
                    });
        run_mma1(it, pipe_addr + P_B0);
        const int n = rank * (int)it.rows + warp * 32 + lane;
        const float bias = (warp * 32 + lane < (int)it.rows) ? __ldg(LC.b_ffn0 + n) : 0.f;
        wait_acc();
        epi_tile(tmem + ACC0, it.rows, A, warp, lane, [&](int, int a, float v) {
          h16[(size_t)a * F + n] = __float2bfloat16_rn(fmaxf(v + bias, 0.f));
        });
        tc_fence_before();
        csync();
      }
      mark(26);
      // ============ FFN down: y3 = h.W2 + b   (no residual, pre-norm3)
      {
        const ResItem it = C.items[item_idx++];
        bt_rows_bf16(pipe + P_B0, h16, A, F, tid);
        run_mma1(it, pipe_addr + P_B0);
        const int n = rank * (int)it.rows + warp * 32 + lane;
        const float bias = (warp * 32 + lane < (int)it.rows) ? __ldg(LC.b_ffn2 + n) : 0.f;
        wait_acc();
        epi_tile(tmem + ACC0, it.rows, A, warp, lane,
                 [&](int, int a, float v) { y3[(size_t)a * D + n] = v + bias; });
        tc_fence_before();
        csync();
      }
      mark(27);
      // ============ reg / cls hidden 1 on x3 = FiLM(LN3(y3))   (:221-231,370-373)
      {
        const ResItem it = C.items[item_idx++];
        ResItem itc = it;
        if (want_cls) itc = C.items[item_idx++];
        const float* film = C.film + ((size_t)si * L + l) * 2 * D;
        float sc[8], sh[8];
        load8(film, lane, sc);
        load8(film + D, lane, sh);
        bt_rows_f32(pipe + P_B0, A, warp, lane, [&](int a) { return y3 + (size_t)a * D; },
                    [&](int, float (&v)[8]) {
                      layer_norm_row(v, LC.norm3_g, LC.norm3_b, lane);
#pragma unroll
                      for (int i = 0; i < 8; ++i) v[i] = v[i] * (1.0f + sc[i]) + sh[i];
                    });
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (is_ctrl) {
          mma_item(it, pipe_addr + P_B0, ACC0);
          if (want_cls) mma_item(itc, pipe_addr + P_B0, ACC1);
          umma_commit(lin_acc);
          prefetch(cons_seq + NSLOT - 1);
        }
        const int f = warp * 32 + lane;
        const int n = rank * (int)it.rows + f;
        const float bias_r = (f < (int)it.rows) ? __ldg(LC.b_reg0 + n) : 0.f;
        const float bias_c = (want_cls && f < (int)it.rows) ? __ldg(LC.b_cls0 + n) : 0.f;
        wait_acc();
        epi_tile(tmem + ACC0, it.rows, A, warp, lane, [&](int, int a, float v) {
          r1_16[(size_t)a * D + n] = __float2bfloat16_rn(fmaxf(v + bias_r, 0.f));
        });
        if (want_cls)
          epi_tile(tmem + ACC1, it.rows, A, warp, lane,
                   [&](int, int a, float v) { c1[(size_t)a * D + n] = fmaxf(v + bias_c, 0.f); });
        tc_fence_before();
        csync();
      }
      mark(28);
      // ============ reg / cls hidden 2
      {
        const ResItem it = C.items[item_idx++];
        ResItem itc = it;
        if (want_cls) itc = C.items[item_idx++];
        bt_rows_bf16(pipe + P_B0, r1_16, A, D, tid);
        if (want_cls)
          bt_rows_f32(pipe + P_B1, A, warp, lane, [&](int a) { return c1 + (size_t)a * D; },
                      [&](int, float (&v)[8]) { layer_norm_row(v, LC.cls_ln2_g, LC.cls_ln2_b, lane); });
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (is_ctrl) {
          mma_item(it, pipe_addr + P_B0, ACC0);
          if (want_cls) mma_item(itc, pipe_addr + P_B1, ACC1);
          umma_commit(lin_acc);
          prefetch(cons_seq + NSLOT - 1);
        }
        const int f = warp * 32 + lane;
        const int n = rank * (int)it.rows + f;
        const float bias_r = (f < (int)it.rows) ? __ldg(LC.b_reg2 + n) : 0.f;
        const float bias_c = (want_cls && f < (int)it.rows) ? __ldg(LC.b_cls3 + n) : 0.f;
        wait_acc();
        epi_tile(tmem + ACC0, it.rows, A, warp, lane,
                 [&](int, int a, float v) { r2[(size_t)a * D + n] = fmaxf(v + bias_r, 0.f); });
        if (want_cls)
          epi_tile(tmem + ACC1, it.rows, A, warp, lane,
                   [&](int, int a, float v) { c2[(size_t)a * D + n] = fmaxf(v + bias_c, 0.f); });
        tc_fence_before();
        csync();
      }
      mark(29);
      // ============ regression head (256 -> 3P, fp32 weights), outputs dealt over the CTAs;
      // rank 0 also finishes the cls branch: scores = LN(c2).w6 + b6   (:221-224)
      {
        float* r2_s = reinterpret_cast<float*>(pipe);   // [A][256]
        {
          const int n4 = A * D / 4;
          for (int base = tid; base < n4; base += RT * 8) {
            float4 t[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = base + u * RT;
              if (i < n4) t[u] = __ldcg(reinterpret_cast<const float4*>(r2) + i);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int i = base + u * RT;
              if (i < n4) reinterpret_cast<float4*>(r2_s)[i] = t[u];
            }
          }
        }
        __syncthreads();
        const int total = A * 3 * P, per = (total + RES_CL - 1) / RES_CL;
        const int g0 = rank * per, g1 = min(total, g0 + per);
        for (int g = g0 + warp; g < g1; g += 8) {
          const int a = g / (3 * P), c = g - a * (3 * P);
          float w[8], x[8];
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(LC.reg4_w + (size_t)c * D + lane * 4));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(LC.reg4_w + (size_t)c * D + 128 + lane * 4));
          w[0] = w0.x; w[1] = w0.y; w[2] = w0.z; w[3] = w0.w; w[4] = w1.x; w[5] = w1.y; w[6] = w1.z; w[7] = w1.w;
          load8(r2_s + (size_t)a * D, lane, x);
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) s = fmaf(x[i], w[i], s);
          s = warp_sum(s);
          if (lane == 0) regraw[g] = s + __ldg(LC.reg4_b + c);
        }
        if (want_cls && rank == 0) {
          for (int a = warp; a < A; a += 8) {
            float v[8], w[8];
            ldcg8(c2 + (size_t)a * D, lane, v);
            layer_norm_row(v, LC.cls_ln5_g, LC.cls_ln5_b, lane);
            load8(LC.cls6_w, lane, w);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(v[i], w[i], s);
            s = warp_sum(s);
            if (lane == 0) {
              const float sc = s + __ldg(LC.cls6_b);
              scores_s[a] = sc;
              if (call.out_scores) call.out_scores[(size_t)scene * A + a] = sc;
            }
          }
        }
        csync();
      }
      mark(30);
      // ============ reg[..., :2] += points; heading = tanh(.)*pi; next points; DDIM update
      // (:378-380,424,632-636), redundantly in every CTA
      {
        const DdimCoef dc = C.dc[si];
        float* modes_s = reinterpret_cast<float*>(pipe + P_MODES);
        for (int g = tid; g < A * 3 * P; g += RT) {
          const int ap = g / 3, comp = g - ap * 3;     // ap = a*P + p
          const float mine = __ldcg(regraw + g);
          float out;
          if (comp < 2) {
            const int pi = ap * 2 + comp;
            out = __fadd_rn(mine, pts_s[pi]);
            pts_s[pi] = out;
            if (do_ddim) {
              const float x0 = comp ? norm_y(out) : norm_x(out);
              const float sample = img_s[pi];
              const float eps = __fdiv_rn(__fsub_rn(sample, __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
              const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
              img_s[pi] = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
            }
          } else {
            out = __fmul_rn(tanhf(mine), 3.14159265358979323846f);
          }
          if (rank == 0) {
            modes_s[g] = out;
            if (want_cls && call.out_modes) call.out_modes[(size_t)scene * A * 3 * P + g] = out;
          }
        }
        __syncthreads();
      }
      mark(31);
    }
  }
  // ================= mode = argmax(cls) (first maximum wins), trajectory = reg[mode]  (:637-640)
  if (rank == 0) {
    const float* modes_s = reinterpret_cast<const float*>(pipe + P_MODES);
    int best = 0;
    float bv = scores_s[0];
    for (int a = 1; a < A; ++a) {
      const float v = scores_s[a];
      if (v > bv) { bv = v; best = a; }
    }
    if (tid == 0 && call.out_mode_idx) call.out_mode_idx[scene] = best;
    if (call.out_traj)
      for (int i = tid; i < 3 * P; i += RT) call.out_traj[(size_t)scene * 3 * P + i] = modes_s[best * 3 * P + i];
  }
  mark(99);
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<TMEM_COLS>(tmem);
}

#undef emb16
#undef o16
#undef h16
#undef r1_16
#undef e1
#undef q0
#undef spart
#undef x1
#undef y2
#undef y3
#undef c1
#undef r2
#undef c2
#undef regraw
#undef kvbuf
#undef egov

}  // namespace

int res_smem_bytes() { return SMEM_BYTES; }

static bool g_res_ready = false;

int res_engine_init() {
  if (g_res_ready) return 0;
  cudaError_t e = cudaFuncSetAttribute(res_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }
  e = cudaFuncSetAttribute(res_forward_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { cudaGetLastError(); return 2; }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(RES_CL);
  cfg.blockDim = dim3(RT);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RES_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, res_forward_kernel, &cfg);
  if (e != cudaSuccess || nclusters < 1) { cudaGetLastError(); return 3; }
  g_res_ready = true;
  return 0;
}

int launch_res_forward(const ResConsts* consts_dev, const ResCall& call, int B, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(RES_CL * B);
  cfg.blockDim = dim3(RT);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = RES_CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, res_forward_kernel, consts_dev, call);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ddh
