// Scene-tile chain engine of the ddh planning head (sm_100a): the throughput path's decoder chain.
//
// One persistent CTA per SM walks 128-row tiles (spt scenes x A anchors, 6 x 20 = 120 rows for the
// reference shape).  For one decoder-layer call (CustomTransformerDecoderLayer.forward after the
// BEV sampling, transfuser_model_v2.py:343-382, task decoder :208-256) a tile runs
//
//   x1 = S.Wbev_out + b + q0          (modules/blocks.py:127-129)
//   q  = x1.Wq + bq ; o = softmax(q K^T / sqrt(32)) V per scene and head       (:355-357)
//   x2 = LN2(LN1(x1 + o.Wout + b) + ego)                                       (:358-365)
//   x3 = FiLM(LN3(ReLU(x2.W0 + b0).W2 + b2))                                   (:368-373)
//   reg = reg4(ReLU(reg2(ReLU(reg0(x3))))) ; + points ; tanh*pi ; DDIM          (:376-380, :632-636)
//   cls = cls6(LN(ReLU(cls3(LN(ReLU(cls0(x3)))))))        (last call only, :244-251)
//
// with every activation resident on the SM: GEMM accumulators in TMEM (two 256-column buffers),
// bf16 A operands written by the epilogues straight into the 128-byte-swizzled K-major layout
// tcgen05.mma reads (two 64 KiB operand regions), weights streamed from L2 by two TMA threads
// through a ring of 32 KiB [256 x 64] tiles.  Only S in and points / modes / scores out touch global
// memory.  A second program of the same kernel (mode 1) computes the sine embedding in place
// and runs plan_anchor_encoder (:459-462, modules/blocks.py:22-40).
//
// Roles (384 threads):
//   warps 0-7   compute: epilogues (thread = row x column half; TMEM lane quarter = warp % 4) and the
//               agent attention on mma.sync (warp = head), K|V of one scene at a time staged by
//               bulk copies into the idle operand region
//   warps 8-9   weight producers (a thread's TMA copies run one at a time: two issuers)
//   warp  10    MMA issuer (one lane) + TMEM allocation
//   warp  11    loader: sampled-feature tile of the NEXT tile (TMA, under the current tile's tail)
//               and the per-scene K|V bulk copies
// The program (which GEMMs, which epilogue, where operands live) is a table built on the host
// (ChainArgs), walked in lock step by the three roles.
//
// Residuals ride in TMEM: the epilogue of bev_out writes x1 + b_attn_out back into the accumulator
// buffer that attn_out then accumulates onto; ffn2 accumulates its four k-blocks over the streamed
// FFN hidden blocks.
#include "geom.cuh"
#include "kernels_chain.h"
#include "tc_ptx.cuh"

namespace ddh {
namespace {

constexpr int CH_NT = 384;
constexpr int CH_NCT = 256;
constexpr int CH_CHUNK = 16384;                       // [128 rows][64 k] bf16
constexpr int CH_REGION = 8 * CH_CHUNK;               // two 64 KiB operand regions (chunks 0-3, 4-7)
constexpr int CH_NS = 2;
constexpr int CH_STAGE = 32768;                       // [256 rows][64 k] bf16
constexpr int OFF_RING = CH_REGION;
constexpr int OFF_PAR = OFF_RING + CH_NS * CH_STAGE;
constexpr int OFF_XCH = OFF_PAR + CH_PAR_FLOATS * 4;  // float2 [2][2][128]
constexpr int OFF_BAR = OFF_XCH + 4096;
constexpr int CH_SMEM = OFF_BAR + 256 + 1024;
static_assert(CH_SMEM <= 227 * 1024, "shared memory budget");

// barrier slots (8 bytes each) behind OFF_BAR
// Timeline (args.dbg != nullptr): CTA 0 stamps clock64 during its second tile.  Layout (long long):
//   [0] tile start (compute warp 0), [1 + 4 j + {0,1,2}] compute step j: wait begin, accumulators ready,
//   epilogue done; [128 + 2 j + {0,1}] MMA thread step j: operands ready, all MMAs issued.
#define CH_STAMP(cond, idx) do { if (args.dbg && blockIdx.x == 0 && (cond)) args.dbg[idx] = clock64(); } while (0)

constexpr int B_WFULL = 0, B_WEMPTY = 2, B_AREADY = 4, B_ACC = 5, B_SFULL = 6, B_SAFREE = 7,
              B_KVGO = 8, B_KVFULL = 9, B_KVEMPTY = 11, B_TMEM = 14;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t idesc_bf16(int n) {   // D=f32, A=B=bf16, K-major, M=128
  return (1u << 4) | (1u << 7) | (1u << 10) | (((uint32_t)n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---- compute-warp context
struct CW {
  uint8_t* sm;
  const float* par;
  float2* xch;
  uint32_t trow;     // TMEM address of this thread's lane quarter, column 0
  int r, hf, q, lane, warp;
  int xk;
};

__device__ __forceinline__ void ld_blk(uint32_t taddr, float (&v)[32]) {
  uint32_t u[32];
  tmem_ld32(taddr, u);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void st_blk(uint32_t taddr, const float (&v)[32]) {
  uint32_t u[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(v[i]);
  tmem_st32(taddr, u);
}
__device__ __forceinline__ void add_par(const float* p, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * i);
    v[4 * i + 0] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
  }
}
// 32 columns of row r -> bf16 A operand (128-byte swizzle, K-major: chunk c0/64, row r at r*128,
// 16-byte unit u at (u ^ (r & 7)) * 16)
__device__ __forceinline__ void st_operand(uint8_t* region, int r, int c0, const float (&v)[32]) {
  DDH_ASSERT(r >= 0 && r < 128 && c0 >= 0 && c0 + 32 <= 256);
  uint8_t* base = region + (c0 >> 6) * CH_CHUNK + r * 128;
  const int u0 = (c0 & 63) >> 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 w;
    w.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
    w.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
    w.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
    w.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
    *reinterpret_cast<uint4*>(base + (((u0 + i) ^ (r & 7)) << 4)) = w;
  }
}
// the two threads of a row (column halves, warps q and q + 4) exchange partial sums
__device__ __forceinline__ float2 row_exchange(CW& c, float a, float b) {
  float2* buf = c.xch + (c.xk & 1) * 256;
  ++c.xk;
  buf[c.hf * 128 + c.r] = make_float2(a, b);
  named_bar_sync(2 + c.q, 64);
  return buf[(c.hf ^ 1) * 128 + c.r];
}
__device__ __forceinline__ void ln_stats(CW& c, float s, float q, float& mean, float& rstd) {
  const float2 o = row_exchange(c, s, q);
  mean = (s + o.x) * (1.0f / D);
  const float var = fmaxf((q + o.y) * (1.0f / D) - mean * mean, 0.f);
  rstd = 1.0f / sqrtf(var + LN_EPS);
}

}  // namespace

__global__ void __launch_bounds__(CH_NT, 1) chain_kernel(const __grid_constant__ ChainArgs args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* sm = smem_raw + pad;
  const uint32_t sm_addr = raw_addr + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = sm_addr + OFF_BAR;
  auto BAR = [&](int i) { return bar + i * 8; };
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(sm + OFF_BAR + B_TMEM * 8);
  float* par = reinterpret_cast<float*>(sm + OFF_PAR);

  const int A = args.A, Na = args.Na, B = args.B, spt = args.spt;
  const int rows_per_tile = spt * A;
  const int M = B * A;
  const int n_steps = args.n_steps;
  const int kv_bytes = Na * CH_KV_LD * 2;
  const int kv_nslot = (2 * kv_bytes <= 4 * CH_CHUNK) ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < CH_NS; ++s) { mbar_init(BAR(B_WFULL + s), 1); mbar_init(BAR(B_WEMPTY + s), 1); }
    mbar_init(BAR(B_AREADY), CH_NCT);
    mbar_init(BAR(B_ACC), 1);
    mbar_init(BAR(B_SFULL), 1);
    mbar_init(BAR(B_SAFREE), 1);
    mbar_init(BAR(B_KVGO), 1);
    for (int s = 0; s < 2; ++s) { mbar_init(BAR(B_KVFULL + s), 1); mbar_init(BAR(B_KVEMPTY + s), 8); }
    fence_barrier_init();
  }
  if (warp == 10) tmem_alloc<512>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
  if (warp == 8 && lane == 0) {
    for (int i = 0; i < CH_MAX_MAPS; ++i) tma_prefetch_desc(&args.maps[i]);
    tma_prefetch_desc(&args.smap);
  }
  for (int v = 0; v < args.n_par; ++v) {
    const ChainParSrc ps = args.par[v];
    for (int i = threadIdx.x; i < ps.n; i += CH_NT) par[ps.dst + i] = __ldg(ps.src + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp <= 9) {
      // ===================== weight producers: chunk g of the program goes to producer g % 2
      if (lane == 0) {
        const int me = warp - 8;
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x) {
          for (int j = 0; j < n_steps; ++j) {
            const ChainStep st = args.steps[j];
            for (int o = 0; o < st.nops; ++o) {
              const ChainOp op = args.ops[st.op0 + o];
              const uint32_t bytes = (op.flags & CO_N64) ? 8192u : (uint32_t)CH_STAGE;
              for (int kc = 0; kc < op.nk; ++kc, ++g) {
                if ((int)(g & 1u) != me) continue;
                const uint32_t s = g % CH_NS;
                mbar_wait(BAR(B_WEMPTY + s), ((g / CH_NS) & 1u) ^ 1u);
                mbar_arrive_expect_tx(BAR(B_WFULL + s), bytes);
                tma_load_2d(sm_addr + OFF_RING + s * CH_STAGE, &args.maps[op.map], BAR(B_WFULL + s),
                            (op.k0 + kc) * 64, op.n0);
              }
            }
          }
        }
      }
    } else if (warp == 10) {
      // ===================== MMA issuer
      if (lane == 0) {
        const uint32_t idesc256 = idesc_bf16(256), idesc64 = idesc_bf16(64);
        uint32_t g = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
          for (int j = 0; j < n_steps; ++j) {
            const ChainStep st = args.steps[j];
            mbar_wait(BAR(B_AREADY), (uint32_t)(it * n_steps + j) & 1u);
            if (st.flags & CS_WAIT_S) mbar_wait(BAR(B_SFULL), (uint32_t)it & 1u);
            tc_fence_after();
            CH_STAMP(it == 1, 128 + 2 * j);
            for (int o = 0; o < st.nops; ++o) {
              DDH_ASSERT(st.op0 + o < CH_MAX_OPS);
              const ChainOp op = args.ops[st.op0 + o];
              DDH_ASSERT(op.map < CH_MAX_MAPS && op.a_chunk + op.nk <= 8 && op.acc_col + ((op.flags & CO_N64) ? 64 : 256) <= 512);
              const uint32_t idesc = (op.flags & CO_N64) ? idesc64 : idesc256;
              for (int kc = 0; kc < op.nk; ++kc, ++g) {
                const uint32_t s = g % CH_NS;
                mbar_wait(BAR(B_WFULL + s), (g / CH_NS) & 1u);
                tc_fence_after();
                const uint32_t a = sm_addr + (op.a_chunk + kc) * CH_CHUNK;
                const uint32_t b = sm_addr + OFF_RING + s * CH_STAGE;
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  umma_bf16(tmem_base + op.acc_col, umma_desc_sw128(a + k4 * 32), umma_desc_sw128(b + k4 * 32),
                            idesc, ((op.flags & CO_ACCUM) || kc > 0 || k4 > 0) ? 1u : 0u);
                umma_commit(BAR(B_WEMPTY + s));
              }
            }
            umma_commit(BAR(B_ACC));
            CH_STAMP(it == 1, 128 + 2 * j + 1);
            if (st.flags & CS_KVGO) umma_commit(BAR(B_KVGO));
            if (st.flags & CS_SAFREE) umma_commit(BAR(B_SAFREE));
          }
        }
      }
    } else {
      // ===================== loader: S tile of each tile (TMA) and the per-scene K|V bulk copies
      if (lane == 0 && args.mode == 0) {
        int it = 0;
        uint32_t n = 0;   // running scene count (K|V slot ring)
        // K|V (HBM-resident: 128 MB per layer at 4096 scenes) and q0 of a tile are pulled into L2 one
        // tile ahead, so that the attention's bulk copies and the x1 epilogue see L2 latency
        auto prefetch_tile = [&](int tile) {
          if (tile >= args.n_tiles) return;
          const int nsc = min(spt, B - tile * spt);
          l2_prefetch(args.kv16 + (size_t)tile * spt * Na * CH_KV_LD, (uint32_t)(nsc * kv_bytes));
          l2_prefetch(args.q0t + (size_t)tile * 128 * D, 128 * D * 4);
        };
        prefetch_tile(blockIdx.x);
        for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
          prefetch_tile(tile + gridDim.x);
          if (it > 0) mbar_wait(BAR(B_SAFREE), (uint32_t)(it - 1) & 1u);
          mbar_arrive_expect_tx(BAR(B_SFULL), 4 * CH_CHUNK);
          for (int kc = 0; kc < 4; ++kc)
            tma_load_2d(sm_addr + kc * CH_CHUNK, &args.smap, BAR(B_SFULL), kc * 64, tile * rows_per_tile);
          mbar_wait(BAR(B_KVGO), (uint32_t)it & 1u);
          const int nsc = min(spt, B - tile * spt);
          for (int sl = 0; sl < nsc; ++sl, ++n) {
            const uint32_t slot = n % kv_nslot, use = n / kv_nslot;
            DDH_ASSERT((slot + 1) * (uint32_t)kv_bytes <= 4u * CH_CHUNK || kv_nslot == 1);
            DDH_ASSERT(tile * spt + sl < B);
            mbar_wait(BAR(B_KVEMPTY + slot), (use & 1u) ^ 1u);
            mbar_arrive_expect_tx(BAR(B_KVFULL + slot), kv_bytes);
            bulk_load(sm_addr + slot * (2 * CH_CHUNK),
                      args.kv16 + (size_t)(tile * spt + sl) * Na * CH_KV_LD, kv_bytes, BAR(B_KVFULL + slot));
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================== compute warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    CW c;
    c.sm = sm; c.par = par; c.xch = reinterpret_cast<float2*>(sm + OFF_XCH);
    c.warp = warp; c.lane = lane; c.q = warp & 3; c.hf = warp >> 2;
    c.r = c.q * 32 + lane;
    c.trow = tmem_base + ((uint32_t)(c.q * 32) << 16);
    c.xk = 0;
    const int r = c.r, hf = c.hf;
    const int cbase = hf * 128;
    uint32_t acc_par = 0;
    uint32_t kvn = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < args.n_tiles; tile += gridDim.x, ++it) {
      const long long m = (long long)tile * rows_per_tile + r;
      const bool valid = r < rows_per_tile && m < M;
      const long long mc = valid ? m : 0;    // clamped row for loads of rows that do not exist
      CH_STAMP(it == 1 && threadIdx.x == 0, 0);
      if (args.mode == 1) {
        // ---- clamp + denorm_odo (:601-602), sine embedding (modules/blocks.py:22-40) of poses
        // 4 hf .. 4 hf + 3 of row r, written as chunks p of the A operand (64 features per pose:
        // [embed(y) | embed(x)], feature pairs (2k, 2k+1) = (sin, cos) of the same argument)
        const float* dimt = par + args.steps[0].par[5];
        const int a_idx = (int)(mc % A);
        float iv[8];
        if (args.first_step) {   // img = sqrt(ac) * norm_odo(anchor) + sqrt(1 - ac) * noise   (:591-597)
          const float4 n0 = __ldg(reinterpret_cast<const float4*>(args.noise + mc * 16 + hf * 8));
          const float4 n1 = __ldg(reinterpret_cast<const float4*>(args.noise + mc * 16 + hf * 8) + 1);
          const float4 a0 = __ldg(reinterpret_cast<const float4*>(args.anchors + a_idx * 16 + hf * 8));
          const float4 a1 = __ldg(reinterpret_cast<const float4*>(args.anchors + a_idx * 16 + hf * 8) + 1);
          const float nz[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
          const float an[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float nv = (i & 1) ? norm_y(an[i]) : norm_x(an[i]);
            iv[i] = __fadd_rn(__fmul_rn(args.sa, nv), __fmul_rn(args.sb, nz[i]));
          }
          if (valid) {
            float4* d = reinterpret_cast<float4*>(args.img + m * 16 + hf * 8);
            d[0] = make_float4(iv[0], iv[1], iv[2], iv[3]);
            d[1] = make_float4(iv[4], iv[5], iv[6], iv[7]);
          }
        } else {
          const float4 i0 = *reinterpret_cast<const float4*>(args.img + mc * 16 + hf * 8);
          const float4 i1 = *(reinterpret_cast<const float4*>(args.img + mc * 16 + hf * 8) + 1);
          iv[0] = i0.x; iv[1] = i0.y; iv[2] = i0.z; iv[3] = i0.w;
          iv[4] = i1.x; iv[5] = i1.y; iv[6] = i1.z; iv[7] = i1.w;
        }
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float cl = fminf(fmaxf(iv[i], -1.0f), 1.0f);
          pv[i] = (i & 1) ? denorm_y(cl) : denorm_x(cl);
        }
        if (valid) {
          float4* d = reinterpret_cast<float4*>(args.pts + m * 16 + hf * 8);
          d[0] = make_float4(pv[0], pv[1], pv[2], pv[3]);
          d[1] = make_float4(pv[4], pv[5], pv[6], pv[7]);
        }
        const float two_pi = 6.283185307179586f;
#pragma unroll 1
        for (int pp = 0; pp < 4; ++pp) {
          uint8_t* base = sm + (hf * 4 + pp) * CH_CHUNK + r * 128;
#pragma unroll 1
          for (int half = 0; half < 2; ++half) {   // output order (pos_y, pos_x), blocks.py:39
            const float v = half ? pv[2 * pp] : pv[2 * pp + 1];
            const float num = __fmul_rn(v, two_pi);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                // one Cody-Waite step to [-pi, pi], then the SFU: |error| ~1e-6 for arguments up to
                // ~360 rad, far below the bf16 rounding of the operand (the fp32 engine keeps sinf/cosf)
                const float arg = __fdiv_rn(num, dimt[2 * (4 * u + k)]);
                const float kk = rintf(arg * 0.15915494309189535f);
                float red = fmaf(-kk, 6.28318548202514648f, arg);
                red = fmaf(-kk, -1.7484556000744883e-7f, red);
                float sn, cs;
                __sincosf(red, &sn, &cs);
                w[k] = pack_bf16(sn, cs);
              }
              *reinterpret_cast<uint4*>(base + (((half * 4 + u) ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(BAR(B_AREADY));

      for (int j = 0; j < n_steps; ++j) {
        const ChainStep st = args.steps[j];
        DDH_ASSERT(st.dst_chunk <= 4 && st.acc_col <= 256 && st.epi <= CE_Q0);
        DDH_ASSERT(st.par[0] + 256 <= CH_PAR_FLOATS + 1024);
        CH_STAMP(it == 1 && threadIdx.x == 0, 1 + 4 * j);
        mbar_wait(BAR(B_ACC), acc_par);
        acc_par ^= 1u;
        tc_fence_after();
        CH_STAMP(it == 1 && threadIdx.x == 0, 1 + 4 * j + 1);
        const uint32_t tacc = c.trow + st.acc_col + cbase;
        uint8_t* dst = sm + st.dst_chunk * CH_CHUNK;
        switch (st.epi) {
          case CE_X1: {
            const float* b0 = par + st.par[0] + cbase;
            const float* b1 = par + st.par[1] + cbase;
            const float4* q0 = reinterpret_cast<const float4*>(args.q0t) + ((size_t)tile * 64 + hf * 32) * 128 + r;
            float4 qv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) qv[i] = __ldg(q0 + (size_t)i * 128);
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float4 qn[8];
              if (b < 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) qn[i] = __ldg(q0 + (size_t)((b + 1) * 8 + i) * 128);
              }
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(b0 + b * 32, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                v[4 * i + 0] += qv[i].x; v[4 * i + 1] += qv[i].y; v[4 * i + 2] += qv[i].z; v[4 * i + 3] += qv[i].w;
              }
              st_operand(dst, r, cbase + b * 32, v);
              add_par(b1 + b * 32, v);
              st_blk(tacc + b * 32, v);
              if (b < 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) qv[i] = qn[i];
              }
            }
            tmem_st_wait();
          } break;
          case CE_ATTN: {
            // ---- q (pre-scaled by log2(e) / sqrt(head_dim)) -> bf16 in the operand layout
            const float* bq = par + st.par[0] + cbase;
            const float qscale = 0.17677669529663687f * 1.4426950408889634f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(bq + b * 32, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] *= qscale;
              st_operand(dst, r, cbase + b * 32, v);
            }
            CH_STAMP(it == 1 && threadIdx.x == 0, 64);
            named_bar_sync(1, CH_NCT);
            CH_STAMP(it == 1 && threadIdx.x == 0, 65);
            // ---- attention: warp = head, one scene at a time (K|V staged by the loader in region 0),
            // 16 query rows per m-tile; the output overwrites q in place (same warp, same columns)
            const int h = warp;
            const uint32_t qreg = sm_addr + st.dst_chunk * CH_CHUNK + (h >> 1) * CH_CHUNK;
            uint8_t* oreg = dst + (h >> 1) * CH_CHUNK;
            const int hcol = (h & 1) * 32;       // column of the head inside its 64-wide chunk
            const int g = lane >> 2, t = lane & 3;
            const int nsc = min(spt, B - tile * spt);
            const int MT = (A + 15) >> 4;
            for (int sl = 0; sl < nsc; ++sl, ++kvn) {
              const uint32_t slot = kvn % kv_nslot, use = kvn / kv_nslot;
              CH_STAMP(it == 1 && threadIdx.x == 0, 66 + 3 * sl);
              mbar_wait(BAR(B_KVFULL + slot), use & 1u);
              CH_STAMP(it == 1 && threadIdx.x == 0, 67 + 3 * sl);
              const uint32_t kvb = sm_addr + slot * (2 * CH_CHUNK);
              uint32_t kf[4][4], vf[4][4];
#pragma unroll
              for (int jn = 0; jn < 4; ++jn) {
                const int key = min(8 * jn + (lane & 7), Na - 1);
                ldsm_x4(kvb + key * (CH_KV_LD * 2) + (h * 32 + (lane >> 3) * 8) * 2, kf[jn]);
              }
#pragma unroll
              for (int jd = 0; jd < 4; ++jd) {
                const int key = min(lane, Na - 1);
                ldsm_x4_t(kvb + key * (CH_KV_LD * 2) + (D + h * 32 + 8 * jd) * 2, vf[jd]);
              }
              for (int mt = 0; mt < MT; ++mt) {
                const int row0 = sl * A + 16 * mt;
                uint32_t qf[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const int lrow = min(row0 + ((lane >> 3) & 1) * 8 + (lane & 7), 127);
                  const int unit = (hcol + 16 * ks + (lane >> 4) * 8) >> 3;
                  ldsm_x4(qreg + lrow * 128 + ((unit ^ (lrow & 7)) << 4), qf[ks]);
                }
                float sacc[4][4];
#pragma unroll
                for (int jn = 0; jn < 4; ++jn) {
#pragma unroll
                  for (int e = 0; e < 4; ++e) sacc[jn][e] = 0.f;
#pragma unroll
                  for (int ks = 0; ks < 2; ++ks) mma16816(sacc[jn], qf[ks], kf[jn][2 * ks], kf[jn][2 * ks + 1]);
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int jn = 0; jn < 4; ++jn)
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const int key = 8 * jn + 2 * t + (e & 1);
                    if (key >= Na) sacc[jn][e] = -INFINITY;
                    if (e < 2) mx0 = fmaxf(mx0, sacc[jn][e]); else mx1 = fmaxf(mx1, sacc[jn][e]);
                  }
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
                mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
                mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
                float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
                for (int jn = 0; jn < 4; ++jn)
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float pz = exp2f(sacc[jn][e] - ((e < 2) ? mx0 : mx1));
                    sacc[jn][e] = pz;
                    if (e < 2) sum0 += pz; else sum1 += pz;
                  }
                sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
                sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
                sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
                sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
                const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
                uint32_t pf[2][4];
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) {
                  pf[kk][0] = pack_bf16(sacc[2 * kk][0], sacc[2 * kk][1]);
                  pf[kk][1] = pack_bf16(sacc[2 * kk][2], sacc[2 * kk][3]);
                  pf[kk][2] = pack_bf16(sacc[2 * kk + 1][0], sacc[2 * kk + 1][1]);
                  pf[kk][3] = pack_bf16(sacc[2 * kk + 1][2], sacc[2 * kk + 1][3]);
                }
                const int ra = 16 * mt + g, rb = ra + 8;
#pragma unroll
                for (int jd = 0; jd < 4; ++jd) {
                  float oacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk) mma16816(oacc, pf[kk], vf[jd][2 * kk], vf[jd][2 * kk + 1]);
                  const int col = hcol + 8 * jd + 2 * t;
                  if (ra < A) {
                    const int rr = sl * A + ra;
                    DDH_ASSERT(rr < 128 && col + 1 < 64);
                    *reinterpret_cast<uint32_t*>(oreg + rr * 128 + (((col >> 3) ^ (rr & 7)) << 4) + (col & 7) * 2) =
                        pack_bf16(oacc[0] * inv0, oacc[1] * inv0);
                  }
                  if (rb < A) {
                    const int rr = sl * A + rb;
                    *reinterpret_cast<uint32_t*>(oreg + rr * 128 + (((col >> 3) ^ (rr & 7)) << 4) + (col & 7) * 2) =
                        pack_bf16(oacc[2] * inv1, oacc[3] * inv1);
                  }
                }
              }
              __syncwarp();
              CH_STAMP(it == 1 && threadIdx.x == 0, 68 + 3 * sl);
              if (lane == 0) mbar_arrive(BAR(B_KVEMPTY + slot));
            }
          } break;
          case CE_LN2EGO: {
            float s = 0.f, qq = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) { s += v[i]; qq = fmaf(v[i], v[i], qq); }
            }
            const long long scene = mc / A;
            const float4* ego = reinterpret_cast<const float4*>(args.egov + scene * D + cbase);
            float4 ev[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ev[i] = __ldg(ego + i);
            float mean, rstd;
            ln_stats(c, s, qq, mean, rstd);
            const float* g1 = par + st.par[0] + cbase;
            const float* be1 = par + st.par[1] + cbase;
            s = 0.f; qq = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
              float4 en[8];
              if (b < 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) en[i] = __ldg(ego + (b + 1) * 8 + i);
              }
              float v[32];
              ld_blk(tacc + b * 32, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 gg = *reinterpret_cast<const float4*>(g1 + b * 32 + 4 * i);
                const float4 bb = *reinterpret_cast<const float4*>(be1 + b * 32 + 4 * i);
                v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * gg.x + bb.x + ev[i].x;
                v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y + ev[i].y;
                v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z + ev[i].z;
                v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w + ev[i].w;
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) { s += v[i]; qq = fmaf(v[i], v[i], qq); }
              st_blk(tacc + b * 32, v);
              if (b < 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) ev[i] = en[i];
              }
            }
            tmem_st_wait();
            ln_stats(c, s, qq, mean, rstd);
            const float* g2 = par + st.par[2] + cbase;
            const float* be2 = par + st.par[3] + cbase;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 gg = *reinterpret_cast<const float4*>(g2 + b * 32 + 4 * i);
                const float4 bb = *reinterpret_cast<const float4*>(be2 + b * 32 + 4 * i);
                v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * gg.x + bb.x;
                v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
                v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z;
                v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
              }
              st_operand(dst, r, cbase + b * 32, v);
            }
          } break;
          case CE_RELU: {
            const float* b0 = par + st.par[0] + cbase;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(b0 + b * 32, v);
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              st_operand(dst, r, cbase + b * 32, v);
            }
          } break;
          case CE_LN_FILM:
          case CE_RELU_LN:
          case CE_SCORE: {
            const bool relu = st.epi != CE_LN_FILM;
            const float* b0 = par + st.par[0] + cbase;
            float s = 0.f, qq = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(b0 + b * 32, v);
              if (relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) { s += v[i]; qq = fmaf(v[i], v[i], qq); }
            }
            float mean, rstd;
            ln_stats(c, s, qq, mean, rstd);
            const float* gg_ = par + st.par[1] + cbase;
            const float* bb_ = par + st.par[2] + cbase;
            float dot = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(b0 + b * 32, v);
              if (relu) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = fmaxf(v[i], 0.f);
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float4 gg = *reinterpret_cast<const float4*>(gg_ + b * 32 + 4 * i);
                const float4 bb = *reinterpret_cast<const float4*>(bb_ + b * 32 + 4 * i);
                v[4 * i + 0] = (v[4 * i + 0] - mean) * rstd * gg.x + bb.x;
                v[4 * i + 1] = (v[4 * i + 1] - mean) * rstd * gg.y + bb.y;
                v[4 * i + 2] = (v[4 * i + 2] - mean) * rstd * gg.z + bb.z;
                v[4 * i + 3] = (v[4 * i + 3] - mean) * rstd * gg.w + bb.w;
              }
              if (st.epi == CE_LN_FILM) {
                const float* sc = par + st.par[3] + cbase + b * 32;
                const float* sh = sc + D;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 a4 = *reinterpret_cast<const float4*>(sc + 4 * i);
                  const float4 h4 = *reinterpret_cast<const float4*>(sh + 4 * i);
                  v[4 * i + 0] = v[4 * i + 0] * (1.0f + a4.x) + h4.x;
                  v[4 * i + 1] = v[4 * i + 1] * (1.0f + a4.y) + h4.y;
                  v[4 * i + 2] = v[4 * i + 2] * (1.0f + a4.z) + h4.z;
                  v[4 * i + 3] = v[4 * i + 3] * (1.0f + a4.w) + h4.w;
                }
              }
              if (st.epi == CE_SCORE) {
                const float* w6 = par + st.par[3] + cbase + b * 32;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const float4 w4 = *reinterpret_cast<const float4*>(w6 + 4 * i);
                  dot = fmaf(v[4 * i + 0], w4.x, dot); dot = fmaf(v[4 * i + 1], w4.y, dot);
                  dot = fmaf(v[4 * i + 2], w4.z, dot); dot = fmaf(v[4 * i + 3], w4.w, dot);
                }
              } else {
                st_operand(dst, r, cbase + b * 32, v);
              }
            }
            if (st.epi == CE_SCORE) {
              const float2 o = row_exchange(c, dot, 0.f);
              if (hf == 0 && valid) args.scores[m] = (dot + o.x) + par[st.par[4]];
            }
          } break;
          case CE_TAIL: {
            // reg = r2 . W4^T as [hi | lo] bf16 halves of W4 (columns 0-31 / 32-63), fp32 sum; the
            // thread of column half hf finishes poses 4 hf .. 4 hf + 3 (outputs 12 hf .. 12 hf + 11)
            float hi[32], lo[32];
            ld_blk(c.trow + st.acc_col, hi);
            ld_blk(c.trow + st.acc_col + 32, lo);
            if (valid) {
              DDH_ASSERT(m >= 0 && m < M);
              const float* b4 = par + st.par[0] + 12 * hf;
              float raw[12];
#pragma unroll
              for (int o = 0; o < 12; ++o) {
                float h_ = 0.f, l_ = 0.f;
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // compile-time indices: no local-memory arrays
                  if (hf == k) { h_ = hi[12 * k + o]; l_ = lo[12 * k + o]; }
                }
                raw[o] = (h_ + l_) + b4[o];
              }
              float pt[8], im[8];
              {
                const float4 p0 = *reinterpret_cast<const float4*>(args.pts + m * 16 + 8 * hf);
                const float4 p1 = *(reinterpret_cast<const float4*>(args.pts + m * 16 + 8 * hf) + 1);
                pt[0] = p0.x; pt[1] = p0.y; pt[2] = p0.z; pt[3] = p0.w;
                pt[4] = p1.x; pt[5] = p1.y; pt[6] = p1.z; pt[7] = p1.w;
              }
              if (args.do_ddim) {
                const float4 p0 = *reinterpret_cast<const float4*>(args.img + m * 16 + 8 * hf);
                const float4 p1 = *(reinterpret_cast<const float4*>(args.img + m * 16 + 8 * hf) + 1);
                im[0] = p0.x; im[1] = p0.y; im[2] = p0.z; im[3] = p0.w;
                im[4] = p1.x; im[5] = p1.y; im[6] = p1.z; im[7] = p1.w;
              }
              float out[12];
              const DdimCoef dc = args.dc;
#pragma unroll
              for (int o = 0; o < 12; ++o) {
                const int p = o / 3, comp = o - p * 3;
                if (comp < 2) {
                  const float val = __fadd_rn(raw[o], pt[p * 2 + comp]);
                  pt[p * 2 + comp] = val;
                  out[o] = val;
                  if (args.do_ddim) {
                    const float x0 = comp ? norm_y(val) : norm_x(val);
                    const float eps = __fdiv_rn(__fsub_rn(im[p * 2 + comp], __fmul_rn(dc.sqrt_ac_t, x0)), dc.sqrt_1m_ac_t);
                    const float x0c = fminf(fmaxf(x0, -1.0f), 1.0f);
                    im[p * 2 + comp] = __fadd_rn(__fmul_rn(dc.sqrt_ac_prev, x0c), __fmul_rn(dc.sqrt_1m_ac_prev, eps));
                  }
                } else {
                  out[o] = __fmul_rn(tanhf(raw[o]), 3.14159265358979323846f);
                }
              }
              float4* pd = reinterpret_cast<float4*>(args.pts + m * 16 + 8 * hf);
              pd[0] = make_float4(pt[0], pt[1], pt[2], pt[3]);
              pd[1] = make_float4(pt[4], pt[5], pt[6], pt[7]);
              if (args.do_ddim) {
                float4* id = reinterpret_cast<float4*>(args.img + m * 16 + 8 * hf);
                id[0] = make_float4(im[0], im[1], im[2], im[3]);
                id[1] = make_float4(im[4], im[5], im[6], im[7]);
              }
              float4* md = reinterpret_cast<float4*>(args.modes + m * 24 + 12 * hf);
              md[0] = make_float4(out[0], out[1], out[2], out[3]);
              md[1] = make_float4(out[4], out[5], out[6], out[7]);
              md[2] = make_float4(out[8], out[9], out[10], out[11]);
            }
          } break;
          case CE_Q0: {
            const float* b0 = par + st.par[0] + cbase;
            float4* q0 = reinterpret_cast<float4*>(args.q0t) + ((size_t)tile * 64 + hf * 32) * 128 + r;
            // hoisted attention-weight logits: this thread's half of the dot products of its row (fp32, the
            // plan kernel adds the two halves and the bias)
            const int nlog = args.logit_part ? (int)st.par[2] : 0;
            const float* lw = par + st.par[1] + cbase;
            float lacc[CH_LOGITS];
#pragma unroll
            for (int o = 0; o < CH_LOGITS; ++o) lacc[o] = 0.f;
#pragma unroll 1
            for (int b = 0; b < 4; ++b) {
              float v[32];
              ld_blk(tacc + b * 32, v);
              add_par(b0 + b * 32, v);
#pragma unroll
              for (int i = 0; i < 8; ++i)
                q0[(size_t)(b * 8 + i) * 128] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
              if (nlog) {
#pragma unroll
                for (int o = 0; o < CH_LOGITS; ++o) {
                  if (o < nlog) {
                    const float4* w4 = reinterpret_cast<const float4*>(lw + o * 256 + b * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                      const float4 w = w4[i];
                      lacc[o] = fmaf(v[4 * i], w.x, lacc[o]);
                      lacc[o] = fmaf(v[4 * i + 1], w.y, lacc[o]);
                      lacc[o] = fmaf(v[4 * i + 2], w.z, lacc[o]);
                      lacc[o] = fmaf(v[4 * i + 3], w.w, lacc[o]);
                    }
                  }
                }
              }
            }
            if (nlog) {
              float4* lp = reinterpret_cast<float4*>(args.logit_part + (((size_t)tile * 2 + hf) * 128 + r) * CH_LOGITS);
#pragma unroll
              for (int o = 0; o < CH_LOGITS / 4; ++o)
                lp[o] = make_float4(lacc[4 * o], lacc[4 * o + 1], lacc[4 * o + 2], lacc[4 * o + 3]);
            }
          } break;
          default: break;
        }
        CH_STAMP(it == 1 && threadIdx.x == 0, 1 + 4 * j + 2);
        if (j + 1 < n_steps) {   // the arrival after a tile's last epilogue is the next tile's first one
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(BAR(B_AREADY));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 10) tmem_dealloc<512>(tmem_base);
}

int chain_smem_bytes() { return CH_SMEM; }

int chain_engine_init() {
  return (int)cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM);
}

void launch_chain(const ChainArgs& a, cudaStream_t st) {
  static int num_sms = 0;
  if (!num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (num_sms <= 0) num_sms = 148;
  }
  const int grid = a.n_tiles < num_sms ? a.n_tiles : num_sms;
  chain_kernel<<<grid, CH_NT, CH_SMEM, st>>>(a);
}

}  // namespace ddh
