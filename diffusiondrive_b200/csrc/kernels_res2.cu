// Group-resident engine of the ddh planning head (sm_100a): the whole
// TrajectoryHead.forward_test (transfuser_model_v2.py:578-641) of one scene runs in ONE kernel on
// ONE 16-CTA thread-block cluster, so that a batch-1 forward costs one launch and the decoder
// chain never touches global memory.
//
// Work split.  The cluster is a 4 x 4 grid: CTA (ag, fg) = (rank / 4, rank % 4) belongs to anchor
// group ag (anchors [ag*NAG, +NAG), NAG = ceil(A/4) <= 7) and owns feature slice fg of it.  Every
// Linear of the chain is, per CTA,
//     out^T[f, n] = sum_k W[f, k] * x[n, k]       (tcgen05.mma, M = 64 weight rows, N = 8)
// with the CTA's 64 weight rows (F/4 for the FFN up-projection) as the A operand -- pre-swizzled
// shared-memory images (pack_sw128_kernel) fetched 32 KiB at a time by bulk copies into a two-slot
// ring, by two free-running threads, ahead of the math -- and the group's <= 8 activation rows as
// the B operand.  Epilogue: one thread per output feature reads TMEM, applies bias / ReLU / residual
// and writes the CTA's slice, which the bulk-copy engine pushes into the three peers of the group
// (cp.async.bulk shared::cta -> shared::cluster, completion bytes counted by the peers' exchange
// barrier); every CTA then applies the row operations (LayerNorm, +ego, FiLM) to the full rows,
// one warp per anchor, and writes the bf16 operand of the next stage.  Stages whose consumer needs
// no row operation (FFN hidden, reg hidden, attention output) push bf16 straight into the peers'
// next operand.  Each CTA runs the attention of two heads (its q slice) on K|V staged in shared
// memory.  What crosses groups:
//   * the sampling plan needs every anchor's points and attention weights: they are pushed into
//     all 16 CTAs (st.shared::cluster), every CTA then builds the identical plan (bitmap + popcount
//     compaction, no atomics on hot words);
//   * the on-demand value_proj conv (modules/blocks.py:68-76,114) runs as (64 unique pixels) x
//     (64 output columns) tcgen05 tiles, one per CTA, gathered from the NHWC bf16 map with
//     cp.async in steps of two k-chunks over a four-stage pipeline that borrows the idle weight
//     ring; each CTA pushes its [A x 64] slice of the sampled features to the anchors' groups;
//   * the step-invariant agent K|V and ego projections are computed once, 16-way feature-split,
//     and exchanged through L2;
//   * NCHW callers: the BEV rows a conv call reads are converted on demand, dealt over the CTAs.
// Cluster-wide synchronisation is an mbarrier per CTA that one thread of every CTA arrives on
// remotely (release.cluster), ~13 times per forward.
//
// Three measured properties of the part shape this design (tools/ubench_*.cu, profiles/):
//   * a thread's bulk / TMA copies run one at a time, ~750 cycles each whatever their size
//     -> large (32 KiB) requests from several issuing threads;
//   * a tcgen05.mma or tcgen05.commit costs its issuing thread ~160 cycles whatever its shape,
//     and threads of different warps issue concurrently -> four issuers with private accumulators
//     (fixed k-step assignment keeps sums deterministic);
//   * handing a pipeline stage over costs ~300-400 cycles whatever it carries -> two k-chunks per
//     conv step.
//
// Numerics are those of the bf16 tensor engine (kernels_tc.cu): bf16 operands, fp32 accumulate,
// fp32 LayerNorm / softmax / embeddings / residuals / regression tail.
#include <stdio.h>
#include <stdlib.h>

#include "geom.cuh"
#include "kernels_res2.h"
#include "tc_ptx.cuh"

namespace ddh {
namespace {

constexpr int NT = 384;                      // 8 compute warps + 4 engine warps
constexpr int NPROD = 2;                     // a thread's bulk copies run one at a time (~750 cycles each,
                                             // measured: tools/ubench_ingest.cu), so two threads issue them
constexpr int NCT = 256;                     // compute threads
constexpr int GF = 4;                        // feature groups (= CTAs per anchor group)
constexpr int NROW = 8;                      // activation rows of a chain B operand (anchors per group <= 7)
constexpr int SLOT = 32 * 1024;              // weight ring slot: 64 rows x 256 k bf16, one bulk copy of the
                                             // pre-swizzled image (pack_sw128_kernel)
constexpr int NSLOT = 2;
constexpr int FKC = 4;                       // k-chunks per fill
constexpr int RING = NSLOT * SLOT;
constexpr int CROWS = 64;                    // conv tile: 64 unique pixels x 64 output columns per CTA
constexpr int CCOLS = 64;
constexpr int CA_TILE = CROWS * 128;
constexpr int CNS = 4;                       // conv pipeline stages of TWO k-chunks each (a stage hand-over costs
                                             // ~400 cycles whatever it carries): [A0 | A1 | B0 | B1], 32 KiB; the conv
                                             // borrows the weight ring (idle while it runs): RING + PIPE
constexpr int CSTAGE = 2 * (CA_TILE + CCOLS * 128);
constexpr int PIPE = 2 * CSTAGE;
static_assert(CNS * CSTAGE == RING + PIPE, "conv stages span the weight ring and the pipeline buffers");
constexpr int KC_CONV = 9 * (D / 64);        // 36 k-chunks: (tap, 64-channel chunk)
constexpr int KP_CONV = KC_CONV / 2;         // pipeline steps (pairs of k-chunks)
constexpr int BCH = 1024;                    // chain B operand: 8 rows x 128 B per k-chunk; rows 8..15 of
                                             // the N = 16 operand alias rows 0..7 (descriptor SBO = 0)
constexpr int BCH32 = 4096;                  // hoisted stage: 32 rows x 128 B
// A tcgen05.mma costs its ISSUING THREAD ~160 cycles whatever its shape, and threads of different
// warps issue concurrently (measured: tools/ubench_mma.cu, 159 -> 85 -> 48 cycles per instruction with
// 1 / 2 / 4 issuers).  NMMA threads therefore share every stage: issuer j takes k-step j of each
// 64-wide k-chunk into its own accumulator (a fixed summation order keeps results deterministic)
// and the epilogues add the NMMA accumulators.  The linear stages are issued by lane 0 of compute
// warps 4..7 (idle while the tensor core works), the conv by the four engine warps.
constexpr int NMMA = 4;
constexpr uint32_t ACC_CONV = 0, ACC_LIN = 256;   // TMEM columns: conv 4 x 64; linear tile t, issuer j at ACC_LIN + 32 t + 8 j
                                                  // (N = 8: one 32-column TMEM load fetches a tile's four accumulators)
constexpr int TMEM_COLS = 512;
constexpr int VS_LD = CCOLS + 4;
constexpr int KS_LD = 64 + 4;                // padded K rows (2 heads): conflict-free 128-bit reads, lane = agent

// exchange region X behind RING + PIPE: two chain B operands and two fp32 row buffers that the
// CTAs of an anchor group push into; during the conv phase the same bytes hold the sampled-feature
// partials (pushed by the conv CTAs) and the conv drain staging
constexpr int X_BOP = 0;                     // 2 x 16 KiB
constexpr int X_ACT = 32768;                 // 2 x float [NROW][256]
constexpr int X_SP = 0;                      // float [4 tiles][NAG][256]   (conv phase, <= 28 KiB)
constexpr int X_VS = 30720;                  // float [64][VS_LD]           (conv phase)
constexpr int XBYTES = 49152;
// chain-phase use of the idle conv pipeline buffers
constexpr int P_KV = 0;                      // float Ks[32][KS_LD] | Vv[32][64]
constexpr int P_QL = 20480;                  // float q of my two heads [NROW][64]
constexpr int P_ACT2 = 32768;                // cls branch: 2 x float [NROW][256]
constexpr int P_BOP2 = 49152;                // cls branch B operand (4 KiB)
// fixed region behind X
constexpr int F_CONSTS = 0;                  // R2Consts
constexpr int F_Q0 = 4096;                   // float [NROW][256]
constexpr int F_X1 = F_Q0 + 8192;
constexpr int F_EGO = F_X1 + 8192;           // float [L][256]
constexpr int F_ENT = F_EGO + 4096;          // EntPair [A*P*4] (<= 1024)
constexpr int F_UPIX = F_ENT + 8192;         // int [rcap] (<= 1024)
constexpr int F_BM = F_UPIX + 4096;          // uint [HW/32] pixel bitmap | int [HW/32] prefix
constexpr int F_AW = F_BM + 1024;            // float [L][A*P]
constexpr int F_PTS = F_AW + 4096;           // float [A*P*2] every anchor's current points
constexpr int F_OWN = F_PTS + 2048;          // diffused sample of my outputs: float [NROW][24]
constexpr int F_MISC = F_OWN + 1024;         // conv bias slice [64] | ints [16]
constexpr int F_FIN = F_MISC + 512;          // rank 0: scores [32] | modes [A*3P] (<= 768)
constexpr int F_BAR = F_FIN + 3328;
constexpr int F_END = F_BAR + 320;          // (256..263: anchor-group barrier of the dense mode)
constexpr int SMEM_BYTES = RING + PIPE + XBYTES + F_END + 1024;
static_assert(sizeof(R2Consts) <= F_Q0 - F_CONSTS, "R2Consts must fit its shared-memory slot");
static_assert(sizeof(R2Consts) % 16 == 0, "R2Consts is copied as uint4");
static_assert(X_VS + CROWS * VS_LD * 4 <= XBYTES, "drain staging must fit the exchange region");
static_assert(4 * 7 * D * 4 <= X_VS, "sampled-feature partials must not reach the drain staging");
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
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// D = f32, A = B = bf16, K-major, M = 64, N = n.  An M = 64 instruction occupies the tensor pipe
// for half the time of an M = 128 one whatever N is (measured); its accumulator row r lives in
// TMEM lane 32 * (r / 16) + r % 16 (probed: tools/ubench_m64.cu).
__device__ __forceinline__ constexpr uint32_t idesc_m64(uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((64u >> 4) << 24);
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
// byte offset of element (row n, column k) of a K-major, 128-byte-swizzled bf16 operand whose
// 64-wide k-chunks are `chunk_bytes` apart
__device__ __forceinline__ uint32_t sw_off(int n, int k, int chunk_bytes) {
  return (uint32_t)((k >> 6) * chunk_bytes + n * 128 + ((((k & 63) >> 3) ^ (n & 7)) << 4) + (k & 7) * 2);
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
// shared-memory tile (same scheme as bev_rows_to_nhwc_kernel); 256 compute threads.  Split in a
// load half and a store half so that the loads of the next item are in flight while this one is
// transposed and written.
template <typename TI>
__device__ __forceinline__ void layout_load(const TI* __restrict__ src, int HW, int px0, int tid, float4 (&v)[8]) {
  const int px4 = tid & 7, cl = tid >> 3;
  const TI* s = src + px0 + px4 * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ld4_bev<TI>(s + (size_t)(i * 32 + cl) * HW);
}
__device__ __forceinline__ void layout_store(__nv_bfloat16* __restrict__ dst, int px0, uint32_t* tile_u32, int tid,
                                             const float4 (&v)[8]) {
  constexpr int LDW = 129;
  constexpr int LDE = LDW * 2;
  __nv_bfloat16* tile = reinterpret_cast<__nv_bfloat16*>(tile_u32);
  const int px4 = tid & 7, cl = tid >> 3, lane = tid & 31, warp = tid >> 5;
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

__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                  uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst_cluster), "r"(src_cta), "r"(bytes), "r"(mbar_cluster) : "memory");
}
// global -> my shared memory, one contiguous chunk; the mbarrier receives the byte count
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// LayerNorm of a row held by one warp (8 values per lane).  Sum and sum of squares share the
// five shuffle rounds (the row operation sits on the critical path of every stage); the variance
// E[x^2] - mean^2 is formed in fp32 from O(1) activations.
__device__ __forceinline__ void ln_row_reg(float (&v)[8], const float (&g)[8], const float (&b)[8]) {
  float s = 0.f, q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += v[i];
    q = fmaf(v[i], v[i], q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  const float mean = s * (1.0f / D);
  const float var = fmaxf(q * (1.0f / D) - mean * mean, 0.f);
  const float rstd = rsqrtf(var + LN_EPS);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = (v[i] - mean) * rstd * g[i] + b[i];
}
__device__ __forceinline__ void ldg8(const float* p, int lane, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p + lane * 4));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p + 128 + lane * 4));
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
// fp32 row buffers are slice-major [feature group][row][64] so that a CTA's slice is contiguous:
// lane's 8 values of row n (features lane*4+{0..3} and 128+lane*4+{0..3})
__device__ __forceinline__ void act_load8(const float* act, int n, int lane, float (&o)[8]) {
  const float* p = act + (lane >> 4) * (NROW * 64) + n * 64 + (lane & 15) * 4;
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 2 * NROW * 64);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
  o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

// lane's 8 values of row n (columns lane*4+{0..3} and 128+lane*4+{0..3}) -> bf16 chain B operand
__device__ __forceinline__ void bt_store8(uint8_t* bt, int n, int lane, const float (&v)[8]) {
#pragma unroll
  for (int h2 = 0; h2 < 2; ++h2) {
    const int k = h2 * 128 + lane * 4;
    __nv_bfloat162 p0 = __floats2bfloat162_rn(v[4 * h2 + 0], v[4 * h2 + 1]);
    __nv_bfloat162 p1 = __floats2bfloat162_rn(v[4 * h2 + 2], v[4 * h2 + 3]);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&p0);
    u.y = *reinterpret_cast<uint32_t*>(&p1);
    *reinterpret_cast<uint2*>(bt + sw_off(n, k, BCH)) = u;
  }
}
__device__ __forceinline__ void store8(float* p, int lane, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 128 + lane * 4) = make_float4(v[4], v[5], v[6], v[7]);
}


// ---------------------------------------------------------------------------------------------
// Helper clusters of a dense-mode launch (DenseArgs, kernels_res2.h): whole-map value_proj + ReLU
// (modules/blocks.py:68-76,114) for every layer, ahead of the scene cluster that samples it.
//   layout job (b, y):   BEV row y of scene b, NCHW f32/bf16 -> NHWC bf16 (256 threads), flag DC_ROW
//   conv job (l, b, t, half):  pixels [128 t, +128) (two map rows) x 128 output channels, K = 9 x 256 as
//                        36 k-chunks of (tap, 64 channels).  A tile = TMA box {64 ch, 64 px, 2 rows} of the
//                        NHWC map at (dx, y0 + dy): the zero padding of the conv is the TMA's out-of-bounds
//                        fill; B tile = [128 x 64] of the packed weights.  Four 32 KiB stages, one A and one
//                        B producer thread per stage (a thread's TMA copies run one at a time), two MMA
//                        issuers with private accumulators (chunks of equal parity; a thread spends
//                        ~160 cycles per tcgen05.mma), epilogue on all twelve warps: sum, + bias, ReLU,
//                        bf16, 256-byte half lines to V; counter DC_VDONE.  Layer l starts when layer l - 1
//                        is complete: the 64 jobs of one layer already pull what the chip's L2 delivers
//                        (~6.3 KB/clk), so the first map is ready at ~15 us instead of ~21 us.
// Jobs are claimed from DC_JOB in order (layout first), so a conv job only ever waits for layout
// jobs that running CTAs hold.
constexpr int HSTAGE = 32768, HNS = 4, HKC = 36;   // stage = A tile [128 px x 64] + B tile [128 ch x 64]
static_assert(HNS * HSTAGE <= RING + PIPE + XBYTES + F_UPIX, "helper stages end below its small shared-memory items");
__device__ __forceinline__ void dense_helper_role(const ResCall& call, const DenseArgs& da, uint8_t* sm,
                                                  uint32_t sm_addr, uint8_t* fix, uint32_t hbar, uint32_t tmem, int tid,
                                                  long long* hdbg) {
  const int warp = tid >> 5, lane = tid & 31;
  const long long t_enter = clock64();
  const int H = da.H, W = da.W, L = da.L, B = da.B, HW = H * W;
  const int tiles = HW / 128;
  const int hpr = W / 32;                              // 32-pixel layout jobs per BEV row
  const int n_layout = da.nhwc ? B * H * hpr : 0;
  const int total = n_layout + B * L * tiles * 2;
  int* ctrl = da.ctrl;
  volatile int* job_slot = reinterpret_cast<volatile int*>(fix + F_BAR + 248);
  float* bias_s = reinterpret_cast<float*>(fix + F_UPIX);
  auto full = [&](int s) { return hbar + s * 8; };
  auto empty = [&](int s) { return hbar + (HNS + s) * 8; };
  const uint32_t accf = hbar + 2 * HNS * 8;
  uint32_t g = 0, acc_par = 0;   // k-chunks this CTA has pipelined so far; accumulator phase
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&da.amap);
    for (int l = 0; l < L; ++l) tma_prefetch_desc(&da.wmap[l]);
  }
  // tid 0 keeps one claim in flight: the job after the current one is requested before the current
  // one starts, so the atomic's round trip never sits between two jobs (the first job was claimed
  // in the kernel prologue)
  int next_job = 0;
  if (tid == 0) next_job = atomicAdd(ctrl + DC_JOB, 1);
  for (bool first = true;; first = false) {
    if (tid == 0 && !first) {
      *job_slot = next_job;
      if (next_job < total) next_job = atomicAdd(ctrl + DC_JOB, 1);
    }
    __syncthreads();
    const int job = *job_slot;
    __syncthreads();
    if (job >= total) break;
    if (job < n_layout) {
      const int b = job / (H * hpr), r = job - b * (H * hpr), y = r / hpr, x0 = (r - y * hpr) * 32;
      if (tid < NCT) {
        __nv_bfloat16* dst = da.nhwc + (size_t)b * HW * D;
        uint32_t* tile_u32 = reinterpret_cast<uint32_t*>(sm);
        float4 v[8];
        long long* hl = (hdbg && first && tid == 0) ? hdbg + 13 : nullptr;   // debug timeline of the CTA's first layout job
        if (hl) { hl[0] = job; hl[1] = clock64() - t_enter; }
        if (call.bev_dtype == 0)
          layout_load<float>(reinterpret_cast<const float*>(call.bev) + (size_t)b * D * HW, HW, y * W + x0, tid, v);
        else
          layout_load<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)b * D * HW, HW, y * W + x0, tid, v);
        layout_store(dst, y * W + x0, tile_u32, tid, v);
        if (hl) hl[2] = clock64() - t_enter;
        __threadfence();   // (the generic -> async proxy fence is the reader's: it precedes the TMA loads)
        named_bar_sync(1, NCT);
        if (tid == 0) atomicAdd(ctrl + DC_ROW + b * H + y, 1);   // (release: every writer fenced before the barrier)
        if (hl) hl[3] = clock64() - t_enter;
      }
      continue;
    }
    const int idx = job - n_layout;
    // debug timeline of the CTA that runs conv job 0: clock64 per label, globaltimer at the end
    long long* hd = (hdbg && idx == 0) ? hdbg : nullptr;
    if (hd && tid == 0) { hd[0] = t_enter; hd[1] = clock64(); }
    // conv job = (layer, scene, 128-pixel tile, 128-channel half), the two halves of a tile adjacent in
    // the order (their A tiles hit in L2)
    const int jpl = B * tiles * 2;
    const int l = idx / jpl, r = idx - l * jpl, b = r / (tiles * 2), t = (r - b * (tiles * 2)) >> 1, hf = r & 1;
    const int y0 = t * (128 / W);
    // Layer l waits for layer l - 1: all 64 CTAs of a layer together pull 64 x 32 KiB per k-chunk through
    // the chip's ~6.3 KB/clk of L2 bandwidth, so running both layers at once would only make the first
    // map -- the one the scene cluster is waiting for -- finish later
    if (l > 0) {
      if (tid == 0) wait_flag_ge(ctrl + DC_VDONE + b * L + l - 1, tiles * 2);
      __syncthreads();
    }
    if (warp < HNS) {
      if (lane == 0) {   // A tiles (map patches), stage = warp
        fence_proxy_async();   // the layout tile / the previous job's staging (generic writes) alias the stages
        for (int c = warp; c < HKC; c += HNS) {
          const uint32_t gi = g + c;
          const int s = (int)(gi % HNS);
          const uint32_t use = gi / HNS;
          if (use > 0) mbar_wait(empty(s), (use - 1) & 1u);
          const int tap = c >> 2, dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
          const uint32_t st_addr = sm_addr + s * HSTAGE;
          mbar_arrive_expect_tx(full(s), 16384);
          if (c == warp && da.nhwc) {   // the rows (with halo) this tile reads
            // (the <= 4 flags are read together and acquired with one fence: four dependent
            // ld.acquire round trips cost ~2.8 k cycles even when every row is ready)
            const int ylo = max(0, y0 - 1), yhi = min(H - 1, y0 + 128 / W);
            const volatile int* rf = ctrl + DC_ROW + b * H;
#ifdef DDH_CHECKED
            unsigned long long spins = 0;
#endif
            for (;;) {
              int lo = hpr;
              for (int y = ylo; y <= yhi; ++y) lo = min(lo, rf[y]);
              if (lo >= hpr) break;
              __nanosleep(32);
#ifdef DDH_CHECKED
              if (++spins > (1ull << 26)) { printf("DDH_CHECKED: row flag wait timed out (block %d)\n", (int)blockIdx.x); __trap(); }
#endif
            }
            __threadfence();
            fence_proxy_async_all();
            if (hd && warp == 0) hd[2] = clock64();
          }
          tma_load_4d(st_addr, &da.amap, full(s), (c & 3) * 64, dx, y0 + dy, b);
        }
      }
    } else if (warp < HNS + 2) {
      if (lane == 0) {   // MMA issuers: k-chunks of equal parity, private accumulators
        const int j = warp - HNS;
        constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // M 128, N 128
        for (int c = j; c < HKC; c += 2) {
          const uint32_t gi = g + c;
          const int s = (int)(gi % HNS);
          mbar_wait(full(s), (gi / HNS) & 1u);
          tc_fence_after();
          if (hd && c == 0) hd[3] = clock64();
          const uint32_t st_addr = sm_addr + s * HSTAGE;
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4)
            umma_bf16(tmem + j * 128, umma_desc_sw128(st_addr + k4 * 32), umma_desc_sw128(st_addr + 16384 + k4 * 32),
                      idesc, (c >= 2 || k4 > 0) ? 1u : 0u);
          umma_commit(empty(s));
        }
        umma_commit(accf);
        if (hd && j == 0) hd[4] = clock64();
      }
    } else if (warp >= 8) {
      if (lane == 0) {   // B tiles (128 weight rows of this half), stage = warp - 8: a thread's TMA copies run one
                         // at a time, so the two operands of a stage come from two threads
        for (int c = warp - 8; c < HKC; c += HNS) {
          const uint32_t gi = g + c;
          const int s = (int)(gi % HNS);
          const uint32_t use = gi / HNS;
          if (use > 0) mbar_wait(empty(s), (use - 1) & 1u);
          mbar_arrive_expect_tx(full(s), 16384);
          tma_load_2d(sm_addr + s * HSTAGE + 16384, &da.wmap[l], full(s), c * 64, hf * 128);
        }
      }
    } else {
      const int et = tid - 192;   // warps 6, 7: the bias of this half
      bias_s[et] = __ldg(da.bias[l] + hf * 128 + et);
      bias_s[et + 64] = __ldg(da.bias[l] + hf * 128 + et + 64);
    }
    __syncwarp();
    __syncthreads();   // the bias is staged; producers and issuers have issued everything
    mbar_wait(accf, acc_par);
    tc_fence_after();
    if (hd && tid == 0) hd[5] = clock64();
    {
      // Epilogue on all twelve warps: warp w reads TMEM lane quarter w % 4 (pixel rows 32 (w % 4) + lane),
      // column blocks {0,1} / {2} / {3} by w / 4: sum of the two issuers' accumulators + bias, ReLU, bf16
      // into a shared-memory line (pitch 272 B: conflict-free 16-byte writes at a 256-byte lane stride;
      // the pipeline stages are idle, every MMA of the job has completed), then the 256-byte half lines
      // of two pixels per warp instruction to V
      constexpr int EPITCH = 272;
      const int q = warp & 3, wg = warp >> 2;
      const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16);
      uint8_t* stg = sm + (size_t)(q * 32 + lane) * EPITCH;
      const int cb0 = wg == 0 ? 0 : wg + 1, cb1 = wg == 0 ? 2 : wg + 2;
#pragma unroll 1
      for (int cb = cb0; cb < cb1; ++cb) {
        uint32_t u0[32], u1[32];
        tmem_ld32(tl + cb * 32, u0);
        tmem_ld32(tl + 128 + cb * 32, u1);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaxf(__uint_as_float(u0[2 * i]) + __uint_as_float(u1[2 * i]) + bias_s[cb * 32 + 2 * i], 0.f);
          const float x1 = fmaxf(__uint_as_float(u0[2 * i + 1]) + __uint_as_float(u1[2 * i + 1]) + bias_s[cb * 32 + 2 * i + 1], 0.f);
          const __nv_bfloat162 pr = __floats2bfloat162_rn(x0, x1);
          pk[i] = *reinterpret_cast<const uint32_t*>(&pr);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
          *reinterpret_cast<uint4*>(stg + cb * 64 + i * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
      }
      tc_fence_before();
      __syncthreads();
      if (hd && tid == 0) hd[7] = clock64();
      uint4* vout = reinterpret_cast<uint4*>(da.V + (((size_t)b * L + l) * HW + (size_t)t * 128) * D + hf * 128);
      for (int m = warp * 2 + (lane >> 4); m < 128; m += NT / 16)
        vout[(size_t)m * 32 + (lane & 15)] = *reinterpret_cast<const uint4*>(sm + (size_t)m * EPITCH + (lane & 15) * 16);
      if (hd && tid == 0) hd[8] = clock64();
      __threadfence();
      __syncthreads();
      if (tid == 0) {
        atomicAdd(ctrl + DC_VDONE + b * L + l, 1);
        if (hd) {
          hd[6] = clock64();
          unsigned long long gt;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
          hd[10] = (long long)gt;
        }
      }
    }
    g += HKC;
    acc_par ^= 1u;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (tid == 0) {
    __threadfence();
    atomicAdd(ctrl + DC_EXIT, 1);
  }
}

// DBG: debug instantiation (clock64 timeline + taps, ResCall::dbg set); the production one carries none of it
template <bool DBG>
__global__ void __launch_bounds__(NT, 1)
res2_forward_kernel(const R2Consts* __restrict__ gconsts, const ResCall call_in,
                    const __grid_constant__ DenseArgs da) {
  ResCall call = call_in;
  if (!DBG) call.dbg = nullptr;
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
  uint8_t* xr = sm + RING + PIPE;
  const uint32_t x_addr = sm_addr + RING + PIPE;
  uint8_t* fix = xr + XBYTES;
  const uint32_t fix_addr = x_addr + XBYTES;
  const R2Consts& C = *reinterpret_cast<const R2Consts*>(fix + F_CONSTS);
  float* q0_s = reinterpret_cast<float*>(fix + F_Q0);
  float* x1_s = reinterpret_cast<float*>(fix + F_X1);
  float* ego_s = reinterpret_cast<float*>(fix + F_EGO);
  EntPair* ent = reinterpret_cast<EntPair*>(fix + F_ENT);
  int* upix_s = reinterpret_cast<int*>(fix + F_UPIX);
  unsigned int* bm_s = reinterpret_cast<unsigned int*>(fix + F_BM);
  int* pre_s = reinterpret_cast<int*>(fix + F_BM + 512);
  float* aw_s = reinterpret_cast<float*>(fix + F_AW);
  float* pts_s = reinterpret_cast<float*>(fix + F_PTS);
  float* img_o = reinterpret_cast<float*>(fix + F_OWN);          // [NROW][24], my outputs only
  float* cbias_s = reinterpret_cast<float*>(fix + F_MISC);       // [64]
  int* ints_s = reinterpret_cast<int*>(cbias_s + 64);            // [16]
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
  const uint32_t xbar0 = conv_acc + 40;   // two exchange barriers (stage parity): bytes pushed by my group land here
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(fix + F_BAR + (2 * NSLOT + 2 * CNS + 8) * 8);
  const uint32_t grp_bar = bar + 256;     // dense mode: barrier of the four CTAs of an anchor group
  const uint32_t gbar = bar + 168;        // dense mode: exchange of the sampled-feature k-chunks inside an anchor group
  const uint32_t hbar = bar + 176;        // helper role (dense mode): 9 barriers, then its job slot at F_BAR + 248
  static_assert((2 * NSLOT + 2 * CNS + 8) * 8 + 4 <= 176 && 176 + (2 * HNS + 1) * 8 <= 248, "barrier area layout");

  {  // constants -> shared memory
    const uint4* s = reinterpret_cast<const uint4*>(gconsts);
    uint4* d = reinterpret_cast<uint4*>(fix + F_CONSTS);
    for (int i = tid; i < (int)(sizeof(R2Consts) / 16); i += NT) d[i] = __ldg(s + i);
  }
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(ring_full(s), 1); mbar_init(ring_empty(s), NMMA); }
    for (int s = 0; s < CNS; ++s) { mbar_init(conv_full(s), NCT + 1); mbar_init(conv_empty(s), 2); }
    mbar_init(conv_acc, NMMA);
    mbar_init(acc_full, NMMA);
    mbar_init(b_ready, NCT);
    mbar_init(conv_go, 1);
    mbar_init(cl_bar, RES_CL);
    mbar_init(xbar0, 1);
    mbar_init(xbar0 + 8, 1);
    mbar_init(gbar, 1);
    mbar_init(grp_bar, GF);
    for (int i = 0; i < HNS; ++i) { mbar_init(hbar + i * 8, 2); mbar_init(hbar + (HNS + i) * 8, 1); }   // helper role: stage full (A and B producer) / empty
    mbar_init(hbar + 2 * HNS * 8, 2);                               // helper role: both issuers' accumulators
    fence_barrier_init();
  }
  // helper CTA of a dense-mode launch: claim the first job now, the atomic's round trip hides
  // under the rest of the prologue
  if (da.enabled && (int)cluster_id_x() >= da.B && tid == 0)
    *reinterpret_cast<volatile int*>(fix + F_BAR + 248) = atomicAdd(da.ctrl + DC_JOB, 1);
  if (warp == 8) tmem_alloc<TMEM_COLS>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  cluster_sync_hw();   // every CTA's barriers exist before anyone arrives remotely

  if (da.enabled && (int)cluster_id_x() >= da.B) {   // helper cluster of a dense-mode launch
    dense_helper_role(call, da, sm, sm_addr, fix, hbar, tmem, tid, (DBG && call.dbg) ? call.dbg + 920 : nullptr);
    tc_fence_before();
    __syncthreads();
    cluster_sync_hw();
    if (warp == 8) tmem_dealloc<TMEM_COLS>(tmem);
    return;
  }

  const int A = C.A, P = C.P, Na = C.Na, L = C.L, S = C.S, H = C.H, W = C.W;
  const int AP = A * P, HW = H * W;
  const int ag = rank >> 2, fg = rank & 3;          // anchor group, feature group (also conv row tile, column group)
  const int NAG = (A + 3) >> 2;                     // anchors per group

  // =============================================================== engine warps 8..11
  // Warps 8 and 9 stream the linear stages' weights (a thread's bulk copies run one at a time, so
  // two threads alternate); all four issue the conv's MMAs, k-step `me` of every k-chunk each.
  if (warp >= 8) {
    // warp-specialised register budget: the engine warpgroup hands its registers to the two compute
    // warpgroups, whose code is compiled for 224 registers instead of spilling at the 168 of a
    // 384-thread block
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (lane == 0) {
      const int me = warp - 8;
      int seq = 0, cg = 0;
      uint32_t gopar = 0, capar = 0;
      for (int si = 0; si < C.n_stages; ++si) {
        const R2Stage stg = C.stages[si];
        if (stg.flags & R2F_CONV) {
          if (call.dense) continue;   // value_proj comes from the helper clusters
          mbar_wait(conv_go, gopar);
          gopar ^= 1u;
          const int nu = *reinterpret_cast<volatile int*>(ints_s + 4);
          const int passes = (nu + 255) / 256;
          constexpr uint32_t idesc = idesc_m64(CCOLS);
          for (int pass = 0; pass < passes; ++pass) {
            if (pass * 256 + ag * CROWS >= nu) continue;
            // Step g's two k-chunks go to issuers 2 (g & 1) and 2 (g & 1) + 1: a thread spends ~160 cycles
            // per tcgen05 instruction (MMA or commit), and a stage is only free again once its issuers
            // are through, so the work of a step is split while every accumulator keeps a fixed set
            // of k-chunks (deterministic sums).
            for (int kp = 0; kp < KP_CONV; ++kp) {
              const int g = cg + kp, s = g % CNS;
              if ((me >> 1) != (kp & 1)) continue;
              const int c = me & 1;
              mbar_wait(conv_full(s), (uint32_t)((g / CNS) & 1));
              tc_fence_after();
              const uint32_t a_stage = sm_addr + s * CSTAGE;
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4)   // issuer 0's accumulator starts at the bias, the others at zero
                umma_bf16(tmem + ACC_CONV + me * CCOLS, umma_desc_sw128(a_stage + c * CA_TILE + k4 * 32),
                          umma_desc_sw128(a_stage + 2 * CA_TILE + c * (CCOLS * 128) + k4 * 32), idesc,
                          (me == 0 || kp >= 2 || k4 > 0) ? 1u : 0u);
              umma_commit(conv_empty(s));
            }
            umma_commit(conv_acc);
            cg += KP_CONV;
            // the conv stages alias the weight ring: nobody refills it before the last MMA has read them
            mbar_wait(conv_acc, capar);
            capar ^= 1u;
          }
          continue;
        }
        if (me >= NPROD) continue;
        const uint32_t fill_bytes = (uint32_t)stg.rows * 128u * FKC;
        const int nfill = (int)stg.mtiles * (int)stg.kchunks / FKC;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(stg.w) +
                              (size_t)((stg.flags & R2F_RANK16) ? rank : fg) * nfill * fill_bytes;
        for (int f = 0; f < nfill; ++f, ++seq) {
          if ((seq % NPROD) != me) continue;
          const int slot = seq % NSLOT, use = seq / NSLOT;
          if (use > 0) mbar_wait(ring_empty(slot), (uint32_t)((use - 1) & 1));
          mbar_arrive_expect_tx(ring_full(slot), fill_bytes);
          bulk_load(sm_addr + slot * SLOT, wsrc + (size_t)f * fill_bytes, fill_bytes, ring_full(slot));
        }
      }
    }
    __syncwarp();
  }
  // =============================================================== compute warps
  else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    const int quad = warp & 3;
    const uint32_t tlane = tmem + ((uint32_t)(quad * 32) << 16);
    const int a0 = ag * NAG;
    const int n_own = max(0, min(NAG, A - a0));
    uint32_t peer[GF];          // shared::cluster base address of the CTAs of my anchor group
#pragma unroll
    for (int j = 0; j < GF; ++j) peer[j] = mapa(sm_addr, (uint32_t)(ag * GF + j));
    uint32_t acc_par = 0, cl_par = 0, conv_par = 0, gpar = 0;
    int cg = 0, dbg_i = 0, k = 0;   // k: linear stage counter (B operand / row buffer parity)
    if (tid == 0) need_s[1] = 0ull;   // BEV rows already converted (kept in shared memory: a spilled copy costs ~5 k cycles per reload)
    const __nv_bfloat16* bevn =
        call.bev_nhwc_bf16 ? reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)scene * HW * D
                           : C.bev_nhwc + (size_t)scene * HW * D;
    float* kvg = C.kv + (size_t)scene * L * Na * 2 * D;
    float* egog = C.egov + (size_t)scene * L * D;

    auto mark = [&](int label) {
      if (DBG && call.dbg && scene == 0 && rank == 0 && tid == 0 && dbg_i < 900)
        call.dbg[dbg_i++] = ((long long)label << 48) | (clock64() & 0xFFFFFFFFFFFFll);
    };
    if (DBG && call.dbg && scene == 0 && rank == 0 && tid == 0) {
      unsigned long long gt;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      call.dbg[920 + 11] = (long long)gt;
    }
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
    // dense mode: after the hoisted stage the anchor groups are independent (each samples V for its own
    // anchors only), so the per-step / per-layer meetings shrink from the cluster to the group's four CTAs
    uint32_t grp_par = 0;
    auto gsync = [&]() {
      mark(104);
      bsync();
      if (tid < GF) mbar_arrive_remote_release(mapa(grp_bar, (uint32_t)(ag * GF + tid)));
      mbar_wait_acq_cluster(grp_bar, grp_par);
      grp_par ^= 1u;
      mark(105);
    };
    // The B operand(s) of the next stage are written: make them visible to the tensor core, meet,
    // and let lane 0 of warps 4..7 issue the stage (issuer j = warp - 4 takes k-step j of every
    // k-chunk into accumulator j).  `nent` schedule entries share the operand wait (cls branch: 2).
    int sidx = 0, wseq = 0;   // next schedule entry; ring fills consumed so far
    auto b_done = [&](int nent) {
      fence_proxy_async();
      tc_fence_before();
      bsync();
      for (int e = 0; e < nent; ++e) {
        const R2Stage stg = C.stages[sidx++];
        const int nfill = (int)stg.mtiles * (int)stg.kchunks / FKC;
        if (warp >= 4 && lane == 0) {
          const int mj = warp - 4;
          tc_fence_after();
          const bool n32 = (stg.flags & R2F_N32) != 0;
          const uint32_t ncol = n32 ? 32u : 8u;
          const uint32_t idesc = idesc_m64(ncol);
          const uint32_t b_addr = stg.bsel == 2 ? pipe_addr + P_BOP2 : x_addr + X_BOP + stg.bsel * 16384;
          const uint32_t bch = n32 ? BCH32 : BCH, sbo = n32 ? 1024u : 0u;
          int sq = wseq;
          for (int mt = 0; mt < (int)stg.mtiles; ++mt)
            for (int kg = 0; kg < (int)stg.kchunks / FKC; ++kg, ++sq) {
              const int slot = sq % NSLOT;
              mbar_wait(ring_full(slot), (uint32_t)((sq / NSLOT) & 1));
              tc_fence_after();
              const uint64_t adesc = umma_desc_sw128(sm_addr + slot * SLOT);
              const uint64_t bdesc = umma_desc_sw128_sbo(b_addr + kg * FKC * bch, sbo);
              const uint32_t astep = ((uint32_t)stg.rows * 128u) >> 4, bstep = bch >> 4;
              const uint32_t dcol = tmem + stg.acc_col + (mt * NMMA + mj) * ncol;
#pragma unroll
              for (int c = 0; c < FKC; ++c)
                umma_bf16(dcol, adesc + (uint64_t)(c * astep + mj * 2), bdesc + (uint64_t)(c * bstep + mj * 2), idesc,
                          (kg | c) ? 1u : 0u);
              umma_commit(ring_empty(slot));
            }
          if (stg.flags & R2F_COMMIT) umma_commit(acc_full);
        }
        __syncwarp();   // the issuer's warp mates wait here instead of spinning beside it
        wseq += nfill;
      }
    };
    auto wait_acc = [&]() {
      mbar_wait(acc_full, acc_par);
      acc_par ^= 1u;
      tc_fence_after();
    };
    auto bop_ptr = [&](int kk) { return xr + X_BOP + (kk & 1) * 16384; };
    auto act_ptr = [&](int kk) { return reinterpret_cast<float*>(xr + X_ACT + (kk & 1) * 8192); };
    // Exchange inside my anchor group.  Warps 0-1 have written this CTA's slice (`bytes` at byte
    // offset `off` from the shared-memory base, same offset in every CTA); it is copied into the
    // three peers by the bulk-copy engine, whose completion bytes the peers' exchange barrier of
    // this stage parity counts.  Optionally a second slice (cls branch) rides in the same phase.
    uint32_t xbits = 0u;   // bit i: parity of exchange barrier i
    auto xchg_send = [&](int kk, uint32_t off, uint32_t bytes, uint32_t off2, uint32_t bytes2, bool all_warps) {
      fence_proxy_async();
      if (all_warps) bsync(); else named_bar_sync(2, 128);
      if (tid == 0) {
        const uint32_t xb = xbar0 + (kk & 1) * 8;
        mbar_arrive_expect_tx(xb, (GF - 1) * (bytes + bytes2));
#pragma unroll
        for (int j = 0; j < GF; ++j) {
          if (j != fg) {
            bulk_copy_to_peer(peer[j] + off, sm_addr + off, bytes, peer[j] + (xb - sm_addr));
            if (bytes2) bulk_copy_to_peer(peer[j] + off2, sm_addr + off2, bytes2, peer[j] + (xb - sm_addr));
          }
        }
      }
    };
    auto xchg_wait = [&](int kk) {
      mbar_wait(xbar0 + (kk & 1) * 8, (xbits >> (kk & 1)) & 1u);
      xbits ^= 1u << (kk & 1);
    };
    // this thread's feature column (warps 0-1: local feature fl of tile mt) of the group's rows
    auto acc8 = [&](uint32_t acc_col, float (&v)[8]) {   // sum of the four issuers' accumulators, fixed order
      uint32_t u[32];
      tmem_ld32(tlane + acc_col, u);
      tmem_ld_wait();
#pragma unroll
      for (int n = 0; n < 8; ++n)
        v[n] = ((__uint_as_float(u[n]) + __uint_as_float(u[8 + n])) + __uint_as_float(u[16 + n])) + __uint_as_float(u[24 + n]);
    };
    static_assert(NMMA == 4, "acc8 adds four 8-column accumulators");
    const uint32_t off_x = (uint32_t)(RING + PIPE);   // offset of X from the shared-memory base

    // ---- img = sqrt(ac) * norm_odo(anchors) + sqrt(1-ac) * noise   (:591-597); every CTA derives
    // the step-0 points of every anchor itself and keeps the diffused sample of the outputs it owns
    for (int i = tid; i < AP * 2; i += NCT) {
      const float a = __ldg(C.anchors + i);
      const float nv = (i & 1) ? norm_y(a) : norm_x(a);
      const float im = __fadd_rn(__fmul_rn(C.sa_tr, nv), __fmul_rn(C.sb_tr, __ldg(call.noise + (size_t)scene * AP * 2 + i)));
      const float v = fminf(fmaxf(im, -1.0f), 1.0f);
      pts_s[i] = (i & 1) ? denorm_y(v) : denorm_x(v);
      const int a_i = i / (2 * P), r = i - a_i * 2 * P, p = r >> 1, comp = r & 1;
      if (a_i >= a0 && a_i < a0 + n_own) img_o[(a_i - a0) * 24 + p * 3 + comp] = im;
    }
    mark(1);

    // ================= hoisted agent K|V and ego vectors (step-invariant, :316-327,355-364),
    // feature-split over the cluster, exchanged through L2
    {
      uint8_t* bop = xr + X_BOP;
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
            const int kk = h2 * 128 + lane * 4;
            __nv_bfloat162 p0 = __floats2bfloat162_rn(vv[4 * h2 + 0], vv[4 * h2 + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(vv[4 * h2 + 2], vv[4 * h2 + 3]);
            uint2 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(bop + sw_off(a, kk, BCH32)) = u;
          }
        }
      }
      b_done(L);
      wait_acc();
      const int rows = 3 * D / RES_CL;   // 48 features of [K | V | ego] per CTA
      if (warp < 4 && quad * 16 < rows) {
        const int f = quad * 16 + lane;
        const int gfeat = rank * rows + f;
        for (int l = 0; l < L; ++l) {
          uint32_t u[32];
          tmem_ld32(tlane + ACC_LIN + 128 * l, u);
          tmem_ld_wait();
#pragma unroll
          for (int j = 1; j < NMMA; ++j) {
            uint32_t u2[32];
            tmem_ld32(tlane + ACC_LIN + 128 * l + 32 * j, u2);
            tmem_ld_wait();
#pragma unroll
            for (int a = 0; a < 32; ++a) u[a] = __float_as_uint(__uint_as_float(u[a]) + __uint_as_float(u2[a]));
          }
          if (lane < 16 && f < rows) {
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
      // ============ sine embedding of the group's anchors (blocks.py:22-40) -> B operand (K = 64 P);
      // features 2m and 2m+1 share their argument: one sincosf per pair
      {
        uint8_t* bop = bop_ptr(k);
        const float two_pi = 6.283185307179586f;
        for (int i = tid; i < n_own * P * 32; i += NCT) {
          const int n = i / (P * 32), r = i - n * P * 32;
          const int p = r >> 5, jj = r & 31, hf = jj >> 4, m = jj & 15;   // pair m of half hf of pose p
          const float v = hf ? pts_s[((a0 + n) * P + p) * 2 + 0] : pts_s[((a0 + n) * P + p) * 2 + 1];   // (pos_y | pos_x)
          // |arg| <= ~400 rad: reduce to [-pi, pi] in fp32 (error ~3e-5 rad, two orders below the bf16
          // rounding of the result) and use the fast intrinsics
          const float arg = __fdiv_rn(__fmul_rn(v, two_pi), __ldg(C.dim_t + 2 * m));
          const float red = fmaf(-two_pi, rintf(arg * 0.15915494309189535f), arg);
          float sv, cv;
          __sincosf(red, &sv, &cv);
          const __nv_bfloat162 pr = __floats2bfloat162_rn(sv, cv);
          *reinterpret_cast<__nv_bfloat162*>(bop + sw_off(n, p * 64 + hf * 32 + 2 * m, BCH)) = pr;
        }
        b_done(1);
      }
      mark(10);
      // ============ plan_anchor_encoder (:459-462): Linear(512->256)+ReLU+LN, Linear(256->256)
      {
        const float bias = (warp < 4 && lane < 16) ? __ldg(C.b_enc0 + fg * 64 + quad * 16 + lane) : 0.f;
        float g[8], bt[8];
        if (warp < n_own) { ldg8(C.enc_ln_g, lane, g); ldg8(C.enc_ln_b, lane, bt); }
        wait_acc();
        if (warp < 4) {
          float v[8];
          acc8(ACC_LIN, v);
          float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
          for (int n = 0; n < NROW; ++n)
            if (lane < 16) sl[n * 64] = fmaxf(v[n] + bias, 0.f);
          xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048, 0, 0, false);
        }
        xchg_wait(k);
        if (warp < n_own) {
          float v[8];
          act_load8(act_ptr(k), warp, lane, v);
          ln_row_reg(v, g, bt);
          bt_store8(bop_ptr(k + 1), warp, lane, v);
        }
        ++k;
        b_done(1);
      }
      mark(11);
      {
        const float bias = (warp < 4 && lane < 16) ? __ldg(C.b_enc3 + fg * 64 + quad * 16 + lane) : 0.f;
        float wa[8][8];   // rows of this CTA's attention-weight head, fetched under the MMA
        float ba = 0.f;
        if (warp < n_own && fg < L) {
#pragma unroll
          for (int p = 0; p < 8; ++p) ldg8(C.layer[fg].attw_w + (size_t)p * D, lane, wa[p]);
          if (lane < P) ba = __ldg(C.layer[fg].attw_b + lane);
        }
        wait_acc();
        if (warp < 4) {
          float v[8];
          acc8(ACC_LIN, v);
          float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
          for (int n = 0; n < NROW; ++n)
            if (lane < 16) sl[n * 64] = v[n] + bias;
          xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048, 0, 0, false);
        }
        xchg_wait(k);
        if (warp < n_own) {
          float v[8];
          act_load8(act_ptr(k), warp, lane, v);
          store8(q0_s + warp * D, lane, v);
          if (call.dbg && fg == 0) store8(C.tap_q0 + ((size_t)scene * A + a0 + warp) * D, lane, v);
          // attention weights (blocks.py:98-100): softmax_p(q0 . Wa^T + ba); CTA fg of the group
          // computes layer fg and hands it to the whole cluster
          if (fg < L) {
            const int l = fg;
            float dots[8];
#pragma unroll
            for (int p = 0; p < 8; ++p) {
              dots[p] = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) dots[p] = fmaf(v[i], wa[p][i], dots[p]);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)   // the eight reductions share the shuffle rounds
#pragma unroll
              for (int p = 0; p < 8; ++p) dots[p] += __shfl_xor_sync(0xffffffffu, dots[p], o);
            float dot = 0.f;
#pragma unroll
            for (int p = 0; p < 8; ++p)
              if (lane == p) dot = dots[p];
            const float lg = (lane < P) ? dot + ba : -INFINITY;
            float mx = lg;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float e = (lane < P) ? expf(lg - mx) : 0.f;
            const float wgt = e / warp_sum(e);
            // lane p holds weight p: lanes 0..P-1 write it to CTAs 0..15
            if (lane < P) {
              const uint32_t dst = fix_addr + F_AW + (uint32_t)((l * AP + (a0 + warp) * P + lane) * 4);
              if (call.dense) {   // only my group samples my anchors
#pragma unroll
                for (int c = 0; c < GF; ++c) st_cluster_f32(mapa(dst, (uint32_t)(ag * GF + c)), wgt);
              } else {
                for (int c = 0; c < RES_CL; ++c) st_cluster_f32(mapa(dst, (uint32_t)c), wgt);
              }
            }
          }
        }
        ++k;
        if (call.dense) gsync(); else csync();
      }
      mark(12);

      for (int l = 0; l < L; ++l) {
        const ResLayerC& LC = C.layer[l];
        const bool last_layer = (l == L - 1);
        const bool want_cls = last_layer && last_step;
        const bool do_ddim = last_layer && !last_step;
        // ============ sampling plan (blocks.py:98-125), identically in every CTA: bitmap of the
        // in-bounds bilinear corners, prefix popcount over its words (pixel order == memory order of
        // the NHWC map), and every corner looks its slot up; the unique-pixel list is written from the
        // corner side (duplicates store the same value), so nobody walks the bitmap bit by bit
        int nu = 0;
        if (call.dense) {
          // ============ dense mode: ReLU(value_proj) of the whole map was written by the helper clusters;
          // sample it at the 4 bilinear corners of my group's anchors (blocks.py:98-125).  Work item =
          // (anchor, half of its 8 points): a lane owns 8 channels, 16 corner lines (512 B each) in flight.
          ++sidx;   // the conv entry of the schedule
          const __nv_bfloat16* Vl = da.V + ((size_t)scene * L + l) * HW * D;
          {  // stage the agent K|V of my two heads (fp32) in the idle conv pipeline buffers
            const float* kvl = kvg + (size_t)l * Na * 2 * D;
            const int n16 = Na * 32;
            for (int i = tid; i < n16; i += NCT) {
              const int jrow = i >> 5, u = i & 31;
              const uint32_t dst = (u < 16) ? pipe_addr + P_KV + (uint32_t)((jrow * KS_LD + u * 4) * 4)
                                            : pipe_addr + P_KV + (uint32_t)(32 * KS_LD * 4 + (jrow * 64 + (u - 16) * 4) * 4);
              const float* src = kvl + (size_t)jrow * 2 * D + (u < 16 ? fg * 64 + u * 4 : D + fg * 64 + (u - 16) * 4);
              cp_async16(dst, src, 16u);
            }
            cp_async_commit();
          }
          // corner (pixel, bilinear x attention weight) pairs of my anchors' points: one thread per point
          if (tid < n_own * P) {
            const int ap = a0 * P + tid;
            const Corners c = corners_of(pts_s[ap * 2 + 0], pts_s[ap * 2 + 1], H, W, C.oc);
            const float a_w = aw_s[l * AP + ap];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              EntPair ep;
              ep.slot = c.pix[q] >= 0 ? c.pix[q] : 0;
              ep.w = c.pix[q] >= 0 ? c.w[q] * a_w : 0.f;
              ent[tid * 4 + q] = ep;
            }
          }
          if (DBG && call.dbg && scene == 0 && rank == 0 && tid == 0 && si == 0 && l == 0) {
            unsigned long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            call.dbg[920 + 12] = (long long)gt;   // the scene cluster reaches its first gather
          }
          if (si == 0) wait_flag_ge(da.ctrl + DC_VDONE + scene * L + l, 2 * (HW / 128));
          bsync();
          mark(120);
          // CTA (ag, fg) reads channels [64 fg, +64) of every corner line of its group's anchors: 8 lanes
          // per line (16 B each), four lines per warp instruction, warp w takes lines 4w..4w+3 of each
          // anchor -> one load per anchor and lane, all in flight together
          // Rolled, register-light loops on purpose (the chain code around it is register-bound and a shuffle
          // reduction here serialised into a 3.5 k-cycle chain): the CTA's 128-byte slices of the corner lines
          // go to shared memory with cp.async, 16 B per lane, a warp instruction fetching lines 4w..4w+3 of
          // one anchor; then lane = channel pair sums its warp's four lines straight from shared memory.
          float* gp = reinterpret_cast<float*>(xr + X_ACT);      // [8 warps][NROW][64] partial sums (row buffers idle)
          {
            const uint32_t wstage = pipe_addr + P_ACT2 + (uint32_t)(warp * 512);   // + j * 4096: [4 lines][128 B]
            const EntPair* em = ent + warp * 4;
            const __nv_bfloat16* vsrc = Vl + fg * 64 + (lane & 7) * 8;
#pragma unroll 1
            for (int j = 0; j < n_own; ++j)
              cp_async16(wstage + j * 4096 + lane * 16, vsrc + (size_t)em[j * 32 + (lane >> 3)].slot * D, 16u);
            cp_async_commit();
            mark(124);
            cp_async_wait_all();
            __syncwarp();
            mark(125);
            float2* pd = reinterpret_cast<float2*>(gp + (size_t)(warp * NROW) * 64) + lane;
            const uint8_t* wst = pipe + P_ACT2 + warp * 512 + lane * 4;
#pragma unroll 1
            for (int j = 0; j < n_own; ++j) {
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const __nv_bfloat162 pr = *reinterpret_cast<const __nv_bfloat162*>(wst + j * 4096 + e * 128);
                const float we = em[j * 32 + e].w;
                s0 = fmaf(we, __low2float(pr), s0);
                s1 = fmaf(we, __high2float(pr), s1);
              }
              pd[j * 32] = make_float2(s0, s1);
            }
          }
          mark(126);
          bsync();
          mark(122);
          // the eight warps' partials in a fixed order -> bf16 -> my k-chunk of the group's B operand,
          // which the bulk-copy engine pushes into the three peers (their chunks land in mine)
          {
            uint8_t* bop = bop_ptr(k);
            for (int i = tid; i < n_own * 32; i += NCT) {
              const int j = i >> 5, c2 = (i & 31) * 2;
              float s0 = 0.f, s1 = 0.f;
#pragma unroll
              for (int w8 = 0; w8 < 8; ++w8) {
                const float2 t2 = *reinterpret_cast<const float2*>(gp + (size_t)(w8 * NROW + j) * 64 + c2);
                s0 += t2.x;
                s1 += t2.y;
              }
              *reinterpret_cast<__nv_bfloat162*>(bop + sw_off(j, fg * 64 + c2, BCH)) = __floats2bfloat162_rn(s0, s1);
            }
            fence_proxy_async();
            bsync();
            mark(123);
            if (tid == 0) {
              const uint32_t off = off_x + X_BOP + (uint32_t)(k & 1) * 16384u + (uint32_t)fg * BCH;
              mbar_arrive_expect_tx(gbar, (GF - 1) * BCH);
#pragma unroll
              for (int j = 0; j < GF; ++j)
                if (j != fg) bulk_copy_to_peer(peer[j] + off, sm_addr + off, BCH, peer[j] + (gbar - sm_addr));
            }
            mbar_wait(gbar, gpar);
            gpar ^= 1u;
          }
          mark(121);
        } else {
          unsigned long long todo;
          {
            const int nwords = HW / 32;
            if (tid < nwords) bm_s[tid] = 0u;
            if (tid < CCOLS) cbias_s[tid] = __ldg(LC.b_conv + fg * CCOLS + tid);
            Corners c;
            float a_w = 0.f;
            if (tid < AP) {
              c = corners_of(pts_s[tid * 2 + 0], pts_s[tid * 2 + 1], H, W, C.oc);
              a_w = aw_s[l * AP + tid];
            }
            mark(110);
            bsync();
            if (tid < AP) {
  #pragma unroll
              for (int q = 0; q < 4; ++q)
                if (c.pix[q] >= 0) atomicOr(bm_s + (c.pix[q] >> 5), 1u << (c.pix[q] & 31));
            }
            bsync();
            mark(111);
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
              // BEV rows (with their 3x3 halo) this warp's words touch: one shared-memory word per warp
              // (a 64-bit atomic per thread on one word costs ~4 k cycles)
              const int y = (tid * 32) / W;
              const unsigned long long m = bits ? ((y > 0) ? (7ull << (y - 1)) : 3ull) : 0ull;
              const unsigned int lo = __reduce_or_sync(0xffffffffu, (unsigned int)m);
              const unsigned int hi = __reduce_or_sync(0xffffffffu, (unsigned int)(m >> 32));
              if (lane == 0) need_s[2 + warp] = ((unsigned long long)hi << 32) | lo;
              named_bar_sync(3, 128);
              int base = incl - cnt;
              for (int w = 0; w < warp; ++w) base += ints_s[w];
              if (tid == 127) ints_s[4] = base + cnt;
              if (tid < nwords) pre_s[tid] = base;
            }
            bsync();
            mark(112);
            nu = ints_s[4];
            const unsigned long long need_any = (need_s[2] | need_s[3]) | (need_s[4] | need_s[5]);
            const unsigned long long need_all = (H >= 64) ? need_any : (need_any & ((1ull << H) - 1ull));
            const unsigned long long done_rows = need_s[1];
            todo = call.bev_nhwc_bf16 ? 0ull : (need_all & ~done_rows);
            if (tid < AP) {
  #pragma unroll
              for (int q = 0; q < 4; ++q) {
                EntPair ep;
                ep.slot = -1;
                ep.w = 0.f;
                if (c.pix[q] >= 0) {
                  const int wd = c.pix[q] >> 5;
                  ep.slot = pre_s[wd] + __popc(bm_s[wd] & ((1u << (c.pix[q] & 31)) - 1u));
                  ep.w = c.w[q] * a_w;
                  const int y = c.pix[q] / W;
                  upix_s[ep.slot] = (y << 16) | (c.pix[q] - y * W);
                }
                ent[tid * 4 + q] = ep;
              }
            }
          }
          bsync();
          if (tid == 0) need_s[1] |= (need_s[2] | need_s[3]) | (need_s[4] | need_s[5]);
          mark(20);
          // ============ on-demand BEV layout: the rows this conv call reads, not converted yet
          if (todo) {
            uint32_t* tile_u32 = reinterpret_cast<uint32_t*>(pipe);
            __nv_bfloat16* dst = C.bev_nhwc + (size_t)scene * HW * D;
            const int tpr = W / 32;
            // my items: (needed row, 32-pixel block) number rank, rank + 16, ...
            auto item_px0 = [&](int it) {   // the it-th (row, block) of `todo`, or -1
              const int row_i = it / tpr;
              if (row_i >= __popcll(todo)) return -1;
              unsigned long long r = todo;
              for (int q = 0; q < row_i; ++q) r &= r - 1;
              return (__ffsll((long long)r) - 1) * W + (it - row_i * tpr) * 32;
            };
            const bool f32 = call.bev_dtype == 0;
            const float* src32 = reinterpret_cast<const float*>(call.bev) + (size_t)scene * D * HW;
            const __nv_bfloat16* src16 = reinterpret_cast<const __nv_bfloat16*>(call.bev) + (size_t)scene * D * HW;
            float4 va[8], vb[8];
            int px_cur = item_px0(rank);
            if (px_cur >= 0) { if (f32) layout_load<float>(src32, HW, px_cur, tid, va); else layout_load<__nv_bfloat16>(src16, HW, px_cur, tid, va); }
            for (int it = rank; px_cur >= 0; it += RES_CL) {
              const int px_next = item_px0(it + RES_CL);
              if (px_next >= 0) { if (f32) layout_load<float>(src32, HW, px_next, tid, vb); else layout_load<__nv_bfloat16>(src16, HW, px_next, tid, vb); }
              layout_store(dst, px_cur, tile_u32, tid, va);
  #pragma unroll
              for (int i = 0; i < 8; ++i) va[i] = vb[i];
              px_cur = px_next;
            }
            csync();
            mark(21);
          } else {
            bsync();
          }
          // ============ value_proj conv at the unique pixels + bilinear/attention combine:
          // CTA (ag, fg) = (64-row tile, 64-column group)
          {
            if (tid == 0) mbar_arrive(conv_go);
            // the conv's weight tiles (8 KiB per k-chunk) are fetched by lane 0 of compute warp g % 8, so
            // that eight threads share the bulk copies (a thread's copies run one at a time)
            const uint8_t* wconv = reinterpret_cast<const uint8_t*>(C.stages[sidx++].w) + (size_t)fg * KC_CONV * (CCOLS * 128);
            const int passes = (nu + 255) / 256;
            const int a_c = tid >> 4, cqd = tid & 15;         // combine: anchors a_c, a_c + 16; 4-column group
            float4 sacc[2];
            sacc[0] = sacc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
            float* Vs = reinterpret_cast<float*>(xr + X_VS);
            for (int pass = 0; pass < passes; ++pass) {
              const int row_base = pass * 256 + ag * CROWS;
              if (row_base >= nu) continue;
              const int rows_valid = min(CROWS, nu - row_base);
              const int j = tid & 7, rb = tid >> 3;
              int rowoff[2];
              uint32_t vmask[2];
  #pragma unroll
              for (int i = 0; i < 2; ++i) {
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
              for (int kp = 0; kp < KP_CONV; ++kp) {
                const int g = cg + kp, s = g % CNS;
                mbar_wait(conv_empty(s), (uint32_t)(((g / CNS) & 1) ^ 1));
                if (lane == 0 && warp == (g & 7)) {   // both weight tiles of the step are contiguous in the packed image
                  mbar_arrive_expect_tx(conv_full(s), 2 * CCOLS * 128);
                  bulk_load(sm_addr + s * CSTAGE + 2 * CA_TILE, wconv + (size_t)kp * (2 * CCOLS * 128), 2 * CCOLS * 128, conv_full(s));
                }
  #pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const int kc = kp * 2 + c;
                  const uint32_t a_dst = sm_addr + s * CSTAGE + c * CA_TILE + dst_base;
                  const int tap = kc >> 2;
                  const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
                  const int tapoff = (dy * W + dx) * D + (kc & 3) * 64;
  #pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    const bool ok = (vmask[i] >> tap) & 1u;
                    const int off = ok ? rowoff[i] + tapoff : 0;
                    cp_async16(a_dst + i * 4096, bevn + off, ok ? 16u : 0u);
                  }
                }
                const int kc = kp;   // (bias preload below runs on the first step)
                if (kc == 0 && warp < 4) {   // accumulators start at the conv bias
  #pragma unroll
                  for (int hh = 0; hh < 2; ++hh) {
                    uint32_t u[32];
  #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                      const uint4 t4 = *reinterpret_cast<const uint4*>(cbias_s + hh * 32 + 4 * q);
                      u[4 * q + 0] = t4.x; u[4 * q + 1] = t4.y; u[4 * q + 2] = t4.z; u[4 * q + 3] = t4.w;
                    }
                    tmem_st32(tlane + ACC_CONV + hh * 32, u);
                  }
                  tmem_st_wait();
                  tc_fence_before();
                }
                cp_async_mbar_arrive_noinc(conv_full(s));
              }
              cg += KP_CONV;
              mark(125);
              mbar_wait(conv_acc, conv_par);
              conv_par ^= 1u;
              tc_fence_after();
              mark(126);
              if (warp < 4) {   // drain (ReLU) into the staging area: row quad*16 + lane lives in lanes 0..15
                float* vrow = Vs + (size_t)(quad * 16 + (lane & 15)) * VS_LD;
  #pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  uint32_t u0[32];
                  tmem_ld32(tlane + ACC_CONV + hh * 32, u0);
                  tmem_ld_wait();
  #pragma unroll
                  for (int j = 1; j < NMMA; ++j) {
                    uint32_t u1[32];
                    tmem_ld32(tlane + ACC_CONV + j * CCOLS + hh * 32, u1);
                    tmem_ld_wait();
  #pragma unroll
                    for (int q = 0; q < 32; ++q) u0[q] = __float_as_uint(__uint_as_float(u0[q]) + __uint_as_float(u1[q]));
                  }
                  if (lane < 16) {
  #pragma unroll
                    for (int q = 0; q < 8; ++q)
                      *reinterpret_cast<float4*>(vrow + hh * 32 + 4 * q) = make_float4(
                          fmaxf(__uint_as_float(u0[4 * q]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 1]), 0.f),
                          fmaxf(__uint_as_float(u0[4 * q + 2]), 0.f), fmaxf(__uint_as_float(u0[4 * q + 3]), 0.f));
                  }
                }
                tc_fence_before();
              }
              bsync();
              mark(127);
  #pragma unroll
              for (int i = 0; i < 2; ++i) {
                const int a = a_c + 16 * i;
                if (a < A) {
                  const EntPair* ea = ent + a * P * 4;
                  for (int q0 = 0; q0 < P * 4; q0 += 8) {   // 8 entries in flight, no branches
                    float wq[8];
                    float4 vq[8];
  #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                      const EntPair e = ea[q0 + q];
                      const int rr = e.slot - row_base;
                      const bool ok = rr >= 0 && rr < rows_valid;
                      wq[q] = ok ? e.w : 0.f;
                      vq[q] = *reinterpret_cast<const float4*>(Vs + (size_t)(ok ? rr : 0) * VS_LD + cqd * 4);
                    }
  #pragma unroll
                    for (int q = 0; q < 8; ++q) {
                      sacc[i].x = fmaf(wq[q], vq[q].x, sacc[i].x); sacc[i].y = fmaf(wq[q], vq[q].y, sacc[i].y);
                      sacc[i].z = fmaf(wq[q], vq[q].z, sacc[i].z); sacc[i].w = fmaf(wq[q], vq[q].w, sacc[i].w);
                    }
                  }
                }
              }
              bsync();
            }
            mark(128);
            // this CTA's [A x 64] slice of the sampled features -> the four CTAs of each anchor's group
  #pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int a = a_c + 16 * i;
              if (a < A) {
                const int g_a = a / NAG, n = a - g_a * NAG;
                const uint32_t d = off_x + X_SP + (uint32_t)(((ag * NAG + n) * D + fg * CCOLS + cqd * 4) * 4);
  #pragma unroll
                for (int jj = 0; jj < GF; ++jj) st_cluster_v4(mapa(sm_addr + d, (uint32_t)(g_a * GF + jj)), sacc[i]);
              }
            }
            mark(129);
            // stage the agent K|V of my two heads (fp32) in the idle conv pipeline buffers
            {
              const float* kvl = kvg + (size_t)l * Na * 2 * D;
              const int n16 = Na * 32;   // per agent: 16 units of K, 16 units of V
              for (int i = tid; i < n16; i += NCT) {
                const int jrow = i >> 5, u = i & 31;
                const uint32_t dst = (u < 16) ? pipe_addr + P_KV + (uint32_t)((jrow * KS_LD + u * 4) * 4)
                                              : pipe_addr + P_KV + (uint32_t)(32 * KS_LD * 4 + (jrow * 64 + (u - 16) * 4) * 4);
                const float* src = kvl + (size_t)jrow * 2 * D + (u < 16 ? fg * 64 + u * 4 : D + fg * 64 + (u - 16) * 4);
                cp_async16(dst, src, 16u);
              }
              cp_async_commit();
            }
            csync();
          }
        }
        mark(22);
        // ============ output_proj + residual (blocks.py:127-129): x1 = S.Wo + b + q0
        {
          const int ntile = min(4, (nu + CROWS - 1) / CROWS);   // tiles of the first pass hold everything a tile CTA accumulated
          if (!call.dense) {   // (dense mode: the gather wrote and exchanged the operand already)
            float sv[8];
            if (warp < n_own) {
              const float* sp = reinterpret_cast<const float*>(xr + X_SP) + warp * D;
              load8(sp, lane, sv);
              for (int t = 1; t < ntile; ++t) {
                float u[8];
                load8(sp + (size_t)t * NAG * D, lane, u);
#pragma unroll
                for (int i = 0; i < 8; ++i) sv[i] += u[i];
              }
            }
            bsync();   // the partials live where the B operand goes
            if (warp < n_own) bt_store8(bop_ptr(k), warp, lane, sv);
          }
          b_done(1);
          const int f = fg * 64 + quad * 16 + lane;
          const float bias = (warp < 4 && lane < 16) ? __ldg(LC.b_bev_out + f) : 0.f;
          wait_acc();
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) sl[n * 64] = v[n] + bias + q0_s[n * D + f];
            xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048, 0, 0, false);
          }
          xchg_wait(k);
          if (warp < n_own) {
            float v[8];
            act_load8(act_ptr(k), warp, lane, v);
            store8(x1_s + warp * D, lane, v);
            if (call.dbg && fg == 0) store8(C.tap_x1 + ((size_t)scene * A + a0 + warp) * D, lane, v);
            bt_store8(bop_ptr(k + 1), warp, lane, v);
          }
          ++k;
          b_done(1);
        }
        mark(23);
        // ============ cross_agent_attention (:316-321,355-357): q projection of my two heads,
        // softmax(qK^T)V, head outputs straight into the group's next B operand
        {
          float* ql = reinterpret_cast<float*>(pipe + P_QL);      // [NROW][64]
          const float bias = (warp < 4 && lane < 16) ? __ldg(LC.b_q + fg * 64 + quad * 16 + lane) : 0.f;
          wait_acc();
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            const float scale = 0.17677669529663687f;   // 1/sqrt(32)
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) ql[n * 64 + quad * 16 + lane] = (v[n] + bias) * scale;
          }
          cp_async_wait_all();
          bsync();
          const float* Ks = reinterpret_cast<const float*>(pipe + P_KV);
          const float* Vv = Ks + 32 * KS_LD;
          uint8_t* bnext = bop_ptr(k + 1);
          // one (anchor, head) pair per HALF-warp, so that the group's <= 14 pairs run in one round:
          // a lane scores agents l16 and l16 + 16 and accumulates output dims l16 and l16 + 16
          {
            const int pr = warp * 2 + (lane >> 4), l16 = lane & 15;
            const bool live = pr < n_own * 2;
            const int n = live ? pr >> 1 : 0, hh = pr & 1;
            const float* qr = ql + n * 64 + hh * 32;
            float s0 = -INFINITY, s1 = -INFINITY;
            if (l16 < Na) {
              s0 = 0.f;
              const float* kr = Ks + l16 * KS_LD + hh * 32;
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 kk = *reinterpret_cast<const float4*>(kr + c4 * 4);
                const float4 qq = *reinterpret_cast<const float4*>(qr + c4 * 4);
                s0 = fmaf(qq.x, kk.x, s0); s0 = fmaf(qq.y, kk.y, s0); s0 = fmaf(qq.z, kk.z, s0); s0 = fmaf(qq.w, kk.w, s0);
              }
            }
            if (l16 + 16 < Na) {
              s1 = 0.f;
              const float* kr = Ks + (l16 + 16) * KS_LD + hh * 32;
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) {
                const float4 kk = *reinterpret_cast<const float4*>(kr + c4 * 4);
                const float4 qq = *reinterpret_cast<const float4*>(qr + c4 * 4);
                s1 = fmaf(qq.x, kk.x, s1); s1 = fmaf(qq.y, kk.y, s1); s1 = fmaf(qq.z, kk.z, s1); s1 = fmaf(qq.w, kk.w, s1);
              }
            }
            float mx = fmaxf(s0, s1);
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float e0 = (l16 < Na) ? expf(s0 - mx) : 0.f;
            const float e1 = (l16 + 16 < Na) ? expf(s1 - mx) : 0.f;
            float den = e0 + e1;
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) den += __shfl_xor_sync(0xffffffffu, den, off);
            const float p0 = e0 / den, p1 = e1 / den;
            float acc0 = 0.f, acc1 = 0.f;
            const float* vcol = Vv + hh * 32 + l16;
            const int nlo = min(Na, 16);
            for (int jj = 0; jj < nlo; ++jj) {
              const float pj = __shfl_sync(0xffffffffu, p0, jj, 16);
              acc0 = fmaf(pj, vcol[jj * 64], acc0);
              acc1 = fmaf(pj, vcol[jj * 64 + 16], acc1);
            }
            for (int jj = 16; jj < Na; ++jj) {
              const float pj = __shfl_sync(0xffffffffu, p1, jj - 16, 16);
              acc0 = fmaf(pj, vcol[jj * 64], acc0);
              acc1 = fmaf(pj, vcol[jj * 64 + 16], acc1);
            }
            if (live) {
              const int f0 = fg * 64 + hh * 32 + l16;
              *reinterpret_cast<__nv_bfloat16*>(bnext + sw_off(n, f0, BCH)) = __float2bfloat16_rn(acc0);
              *reinterpret_cast<__nv_bfloat16*>(bnext + sw_off(n, f0 + 16, BCH)) = __float2bfloat16_rn(acc1);
            }
          }
          xchg_send(k, off_x + X_BOP + ((k + 1) & 1) * 16384 + fg * BCH, BCH, 0, 0, true);
          xchg_wait(k);
          ++k;
          b_done(1);
        }
        mark(24);
        // ============ attention out_proj + residual, norm1, + ego, norm2   (:355-364)
        {
          const int f = fg * 64 + quad * 16 + lane;
          const float bias = (warp < 4 && lane < 16) ? __ldg(LC.b_attn_out + f) : 0.f;
          float g1[8], b1[8], g2[8], b2[8];
          if (warp < n_own) {
            ldg8(LC.norm1_g, lane, g1); ldg8(LC.norm1_b, lane, b1);
            ldg8(LC.norm2_g, lane, g2); ldg8(LC.norm2_b, lane, b2);
          }
          mark(130);
          wait_acc();
          mark(131);
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) sl[n * 64] = v[n] + bias + x1_s[n * D + f];
            mark(132);
            xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048, 0, 0, false);
          }
          mark(133);
          xchg_wait(k);
          mark(134);
          if (warp < n_own) {
            float v[8], eg[8];
            act_load8(act_ptr(k), warp, lane, v);
            ln_row_reg(v, g1, b1);
            load8(ego_s + l * D, lane, eg);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += eg[i];
            ln_row_reg(v, g2, b2);
            bt_store8(bop_ptr(k + 1), warp, lane, v);
          }
          ++k;
          mark(135);
          b_done(1);
        }
        mark(25);
        // ============ FFN up: my F/4 hidden features, h = relu(x2.W1 + b), straight into the
        // group's next B operand (:366-368)
        {
          const int fq = C.F / GF, ntl = fq / 64;          // hidden features per CTA (<= 256), 64-row tiles
          float bias[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            bias[t] = (warp < 4 && lane < 16 && t < ntl) ? __ldg(LC.b_ffn0 + fg * fq + t * 64 + quad * 16 + lane) : 0.f;
          wait_acc();
          if (warp < 4) {
            uint8_t* bnext = bop_ptr(k + 1);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              if (t < ntl) {
                float v[8];
                acc8(ACC_LIN + t * 32, v);
                const int f = fg * fq + t * 64 + quad * 16 + lane;
#pragma unroll
                for (int n = 0; n < NROW; ++n)
                  if (lane < 16) *reinterpret_cast<__nv_bfloat16*>(bnext + sw_off(n, f, BCH)) = __float2bfloat16_rn(fmaxf(v[n] + bias[t], 0.f));
              }
            }
            xchg_send(k, off_x + X_BOP + ((k + 1) & 1) * 16384 + fg * ntl * BCH, ntl * BCH, 0, 0, false);
          }
          xchg_wait(k);
          ++k;
          b_done(1);
        }
        mark(26);
        // ============ FFN down, norm3, time FiLM   (:368-373)
        {
          const float bias = (warp < 4 && lane < 16) ? __ldg(LC.b_ffn2 + fg * 64 + quad * 16 + lane) : 0.f;
          float g3[8], b3[8], sc[8], sh[8];
          if (warp < n_own) {
            const float* film = C.film + ((size_t)si * L + l) * 2 * D;
            ldg8(LC.norm3_g, lane, g3); ldg8(LC.norm3_b, lane, b3);
            ldg8(film, lane, sc); ldg8(film + D, lane, sh);
          }
          wait_acc();
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) sl[n * 64] = v[n] + bias;
            xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048, 0, 0, false);
          }
          xchg_wait(k);
          if (warp < n_own) {
            float v[8];
            act_load8(act_ptr(k), warp, lane, v);
            ln_row_reg(v, g3, b3);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] * (1.0f + sc[i]) + sh[i];
            bt_store8(bop_ptr(k + 1), warp, lane, v);
          }
          ++k;
          b_done(want_cls ? 2 : 1);
        }
        mark(27);
        // ============ reg / cls hidden 1   (:221-231)
        {
          const int f = fg * 64 + quad * 16 + lane;
          float bias_r = 0.f, bias_c = 0.f;
          if (warp < 4) {
            bias_r = __ldg(LC.b_reg0 + f);
            if (want_cls) bias_c = __ldg(LC.b_cls0 + f);
          }
          float gc[8], bc[8];
          if (want_cls && warp < n_own) { ldg8(LC.cls_ln2_g, lane, gc); ldg8(LC.cls_ln2_b, lane, bc); }
          float* act2 = reinterpret_cast<float*>(pipe + P_ACT2);
          wait_acc();
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            uint8_t* bnext = bop_ptr(k + 1);
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) *reinterpret_cast<__nv_bfloat16*>(bnext + sw_off(n, f, BCH)) = __float2bfloat16_rn(fmaxf(v[n] + bias_r, 0.f));
            if (want_cls) {
              acc8(ACC_LIN + 128, v);
              float* sl = act2 + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
              for (int n = 0; n < NROW; ++n)
                if (lane < 16) sl[n * 64] = fmaxf(v[n] + bias_c, 0.f);
            }
            xchg_send(k, off_x + X_BOP + ((k + 1) & 1) * 16384 + fg * BCH, BCH,
                      (uint32_t)(RING + P_ACT2) + fg * 2048, want_cls ? 2048u : 0u, false);
          }
          xchg_wait(k);
          if (want_cls && warp < n_own) {
            float v[8];
            act_load8(act2, warp, lane, v);
            ln_row_reg(v, gc, bc);
            bt_store8(pipe + P_BOP2, warp, lane, v);
          }
          ++k;
          b_done(want_cls ? 2 : 1);
        }
        mark(28);
        // ============ reg / cls hidden 2, regression head (256 -> 3P, fp32), cls score
        {
          const int f = fg * 64 + quad * 16 + lane;
          float bias_r = 0.f, bias_c = 0.f;
          if (warp < 4) {
            bias_r = __ldg(LC.b_reg2 + f);
            if (want_cls) bias_c = __ldg(LC.b_cls3 + f);
          }
          // regression-head rows of this CTA: outputs c = fg*6 .. fg*6+5 (3P = 24)
          const int nout = 3 * P / GF;
          float w4[6][8];
          float b4 = 0.f;   // fetched under the MMA: a global load after the exchange would sit on the critical path
          if (warp < n_own) {
#pragma unroll
            for (int ci = 0; ci < 6; ++ci)
              if (ci < nout) ldg8(LC.reg4_w + (size_t)(fg * nout + ci) * D, lane, w4[ci]);
            if (lane < nout) b4 = __ldg(LC.reg4_b + fg * nout + lane);
          }
          float* act2 = reinterpret_cast<float*>(pipe + P_ACT2 + 8192);
          wait_acc();
          if (warp < 4) {
            float v[8];
            acc8(ACC_LIN, v);
            float* sl = act_ptr(k) + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
            for (int n = 0; n < NROW; ++n)
              if (lane < 16) sl[n * 64] = fmaxf(v[n] + bias_r, 0.f);
            if (want_cls) {
              acc8(ACC_LIN + 128, v);
              float* s2 = act2 + fg * (NROW * 64) + quad * 16 + lane;
#pragma unroll
              for (int n = 0; n < NROW; ++n)
                if (lane < 16) s2[n * 64] = fmaxf(v[n] + bias_c, 0.f);
            }
            xchg_send(k, off_x + X_ACT + (k & 1) * 8192 + fg * 2048, 2048,
                      (uint32_t)(RING + P_ACT2 + 8192) + fg * 2048, want_cls ? 2048u : 0u, false);
          }
          xchg_wait(k);
          mark(29);
          if (warp < n_own) {
            const int n = warp, a = a0 + n;
            float x[8];
            act_load8(act_ptr(k), n, lane, x);
            float dots[6];
#pragma unroll
            for (int ci = 0; ci < 6; ++ci) {
              dots[ci] = 0.f;
              if (ci < nout) {
#pragma unroll
                for (int i = 0; i < 8; ++i) dots[ci] = fmaf(x[i], w4[ci][i], dots[ci]);
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)   // the six reductions share the shuffle rounds
#pragma unroll
              for (int ci = 0; ci < 6; ++ci) dots[ci] += __shfl_xor_sync(0xffffffffu, dots[ci], o);
            float mine = 0.f;
#pragma unroll
            for (int ci = 0; ci < 6; ++ci)
              if (lane == ci) mine = dots[ci];
            float score = 0.f;
            if (want_cls) {   // scores = LN(c2).w6 + b6   (:221-224)
              float c2[8], w6[8], g5[8], b5[8];
              ldg8(LC.cls_ln5_g, lane, g5); ldg8(LC.cls_ln5_b, lane, b5); ldg8(LC.cls6_w, lane, w6);
              act_load8(act2, n, lane, c2);
              ln_row_reg(c2, g5, b5);
              float s = 0.f;
#pragma unroll
              for (int i = 0; i < 8; ++i) s = fmaf(c2[i], w6[i], s);
              score = warp_sum(s) + __ldg(LC.cls6_b);
            }
            // reg[..., :2] += points; heading = tanh(.)*pi; next points; DDIM update (:378-380,424,632-636)
            if (lane < nout) {
              const int cidx = fg * nout + lane, p = cidx / 3, comp = cidx - p * 3;
              mine += b4;
              if (call.dbg) C.tap_regraw[((size_t)scene * A + a) * 3 * P + cidx] = mine;
              float out;
              if (comp < 2) {
                const int pi = (a * P + p) * 2 + comp;
                out = __fadd_rn(mine, pts_s[pi]);
                float nxt = out;
                if (do_ddim) {
                  const DdimCoef dc = C.dc[si];
                  const float x0 = comp ? norm_y(out) : norm_x(out);
                  const float sample = img_o[n * 24 + cidx];
                  const float eps = __fdiv_rn(__fsub_rn(sample, __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
                  const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
                  const float im = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
                  img_o[n * 24 + cidx] = im;
                  const float vc = fminf(fmaxf(im, -1.0f), 1.0f);
                  nxt = comp ? denorm_y(vc) : denorm_x(vc);
                }
                if (!want_cls) {   // every CTA's next plan (and the group's next embedding) needs them
                  const uint32_t dst = fix_addr + F_PTS + (uint32_t)(pi * 4);
                  if (call.dense) {   // only my group embeds / samples my anchors
#pragma unroll
                    for (int cta = 0; cta < GF; ++cta) st_cluster_f32(mapa(dst, (uint32_t)(ag * GF + cta)), nxt);
                  } else {
                    for (int cta = 0; cta < RES_CL; ++cta) st_cluster_f32(mapa(dst, (uint32_t)cta), nxt);
                  }
                }
              } else {
                out = __fmul_rn(tanhf(mine), 3.14159265358979323846f);
              }
              if (want_cls) {
                if (call.out_modes) call.out_modes[((size_t)scene * A + a) * 3 * P + cidx] = out;
                st_cluster_f32(mapa(fix_addr + F_FIN + (uint32_t)((32 + a * 3 * P + cidx) * 4), 0u), out);
                if (cidx == 0) {
                  if (call.out_scores) call.out_scores[(size_t)scene * A + a] = score;
                  st_cluster_f32(mapa(fix_addr + F_FIN + (uint32_t)(a * 4), 0u), score);
                }
              }
            }
          }
          ++k;
          if (call.dense && !want_cls) gsync(); else csync();   // (the final scores / modes go to rank 0: cluster-wide)
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
    // dense mode: the last scene cluster to finish waits until every helper CTA has left its job
    // loop and clears the control words for the next launch
    if (call.dense && rank == 0) {
      int* ctrl = da.ctrl;
      if (tid == 0) {
        const bool last = da.B == 1 || atomicAdd(ctrl + DC_CHAINS, 1) == da.B - 1;
        if (last) wait_flag_ge(ctrl + DC_EXIT, da.n_helper_ctas);
        ints_s[0] = last ? 1 : 0;
      }
      bsync();
      if (ints_s[0])
        for (int i = tid; i < DC_WORDS; i += NCT) ctrl[i] = 0;
    }
    mark(99);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_hw();   // nobody leaves while a peer may still push into its shared memory
  if (warp == 8) tmem_dealloc<TMEM_COLS>(tmem);
}

// bf16 [N][K] (K contiguous) -> the shared-memory image of its 64-wide k-chunks: tiles of `rows` rows,
// [tile][k-chunk][row][128 B], 16-byte units XOR-swizzled by row & 7 (the UMMA 128-byte swizzle),
// so that a CTA's slice of a stage is one contiguous piece a single bulk copy can fetch.
__global__ void pack_sw128_kernel(const uint4* __restrict__ W, uint4* __restrict__ out, int N, int K, int rows) {
  const int upr = K / 8;
  const long long total = (long long)N * upr;
  for (long long u = blockIdx.x * (long long)blockDim.x + threadIdx.x; u < total; u += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(u / upr), j8 = (int)(u - (long long)n * upr);
    const int kc = j8 >> 3, j = j8 & 7, T = n / rows, r = n - T * rows;
    out[((size_t)(T * (K / 64) + kc) * rows + r) * 8 + (j ^ (r & 7))] = W[u];
  }
}

}  // namespace

void launch_pack_sw128(const __nv_bfloat16* W, __nv_bfloat16* out, int N, int K, int rows, cudaStream_t st) {
  const long long total = (long long)N * (K / 8);
  const int blocks = (int)((total + 255) / 256);
  pack_sw128_kernel<<<blocks < 1184 ? blocks : 1184, 256, 0, st>>>(reinterpret_cast<const uint4*>(W),
                                                                    reinterpret_cast<uint4*>(out), N, K, rows);
}

int res2_smem_bytes() { return SMEM_BYTES; }

static bool g_res2_ready[64] = {};   // per device: function attributes are per-device state
static int g_res2_clusters[64] = {};  // co-resident clusters of the engine (occupancy query)

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
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 4;
  if (g_res2_ready[dev]) return 0;
  cudaError_t e = cudaFuncSetAttribute(res2_forward_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       SMEM_BYTES);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(res2_forward_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
  if (e != cudaSuccess) { cudaGetLastError(); return 1; }
  e = cudaFuncSetAttribute(res2_forward_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(res2_forward_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  if (e != cudaSuccess) { cudaGetLastError(); return 2; }
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  res2_cfg(cfg, attr, 1, nullptr);
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, res2_forward_kernel<false>, &cfg);
  if (e != cudaSuccess || nclusters < 1) {
    if (getenv("DDH_VERBOSE")) fprintf(stderr, "ddh: res2 occupancy query: %s, clusters %d\n", cudaGetErrorString(e), nclusters);
    cudaGetLastError();
    return 3;
  }
  g_res2_clusters[dev] = nclusters;
  g_res2_ready[dev] = true;
  return 0;
}

int res2_max_clusters() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  return g_res2_clusters[dev];
}

int launch_res2_forward(const R2Consts* consts_dev, const ResCall& call_in, int B, cudaStream_t st,
                        const DenseArgs* dense) {
  static const DenseArgs no_dense = {};
  ResCall call = call_in;
  call.dense = (dense && dense->enabled) ? 1 : 0;
  const DenseArgs& da = call.dense ? *dense : no_dense;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  res2_cfg(cfg, attr, B + (call.dense ? da.n_helper_ctas / RES_CL : 0), st);
  cudaError_t e = call.dbg ? cudaLaunchKernelEx(&cfg, res2_forward_kernel<true>, consts_dev, call, da)
                           : cudaLaunchKernelEx(&cfg, res2_forward_kernel<false>, consts_dev, call, da);
  return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace ddh
