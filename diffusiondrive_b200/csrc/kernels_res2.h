// Anchor-resident engine of the ddh planning head (sm_100a), second generation of the
// one-launch small-batch path: one 16-CTA thread-block cluster per scene runs the whole
// TrajectoryHead.forward_test (transfuser_model_v2.py:578-641) in ONE kernel, and the decoder
// chain of an anchor never leaves the SM that owns it.  See kernels_res2.cu for the design.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "kernels.h"

namespace ddh {

constexpr int RES_CL = 16;        // CTAs per cluster (non-portable size, opt-in)
constexpr int RES_MAX_L = 4;      // decoder layers
constexpr int RES_MAX_S = 4;      // denoise steps
constexpr int RES_MAX_B = 24;     // scenes per call served by the one-launch engines


struct ResLayerC {
  const float *b_kvego, *b_bev_out, *b_q, *b_attn_out, *b_ffn0, *b_ffn2, *b_reg0, *b_reg2, *b_cls0,
      *b_cls3, *b_conv;
  const float *attw_w, *attw_b, *norm1_g, *norm1_b, *norm2_g, *norm2_b, *norm3_g, *norm3_b;
  const float *cls_ln2_g, *cls_ln2_b, *cls_ln5_g, *cls_ln5_b, *cls6_w, *cls6_b, *reg4_w, *reg4_b;
  const CUtensorMap* conv_map;
};


struct ResCall {
  const float* ego;       // [B][1][256]
  const float* agents;    // [B][Na][256]
  const void* bev;        // NCHW f32/bf16 or NHWC bf16
  int bev_dtype;          // 0 f32, 1 bf16
  int bev_nhwc_bf16;      // 1: gather straight from the caller's NHWC bf16 map
  const float* noise;     // [B][A][P][2]
  float* out_traj;        // [B][P][3] or null
  float* out_modes;       // [B][A][P][3]
  float* out_scores;      // [B][A]
  long long* out_mode_idx;
  long long* dbg;         // optional clock64 stamps of cluster 0 / rank 0
  int dense;              // 1: value_proj comes from the whole-map conv of the helper clusters (DenseArgs)
  int pad0;
};

// Whole-map value_proj for very small batches (B <= DENSE_MAX_B).  One scene's 3x3 conv over the
// full 64x64 map is 4.8 GFLOP per layer -- a few microseconds on the SMs the 16-CTA scene cluster
// leaves idle -- while the on-demand conv (plan, dedup, layout, gathered GEMM, combine) sits four
// times on the scene cluster's critical path.  In dense mode the SAME launch carries extra "helper"
// clusters that (1) convert the NCHW map to NHWC bf16 and (2) run value_proj + ReLU for every pixel
// of every layer on tcgen05 (TMA-fed implicit GEMM, 128 pixels x 256 channels per job), writing
// V[b][l][pixel][256] bf16 to L2; the scene clusters only gather the 4 bilinear corners of their own
// anchors' sample points.  Jobs are claimed from a counter (work stealing: any resident helper CTA
// makes progress, so the scheme cannot deadlock on co-residency); completion flags live in `ctrl`.
constexpr int DENSE_MAX_B = 2;
enum : int {
  DC_JOB = 0,       // next job to claim
  DC_EXIT = 1,      // helper CTAs that have left their job loop
  DC_CHAINS = 2,    // scene clusters that have finished
  DC_VDONE = 8,     // [B][L] half tiles (128 pixels x 128 channels) of V written
  DC_ROW = 16,      // [B][H] BEV row converted to NHWC bf16
  DC_WORDS = DC_ROW + DENSE_MAX_B * 64,
};
struct DenseArgs {
  CUtensorMap amap;                // NHWC bf16 map [B][H][W][256]: box {64 channels, 64 pixels, 2 rows, 1 scene}
  CUtensorMap wmap[RES_MAX_L];     // value_proj weights [256][9*256] (tap, channel), box {64 k, 128 channels}
  const float* bias[RES_MAX_L];
  __nv_bfloat16* V;                // [B][L][H*W][256]
  __nv_bfloat16* nhwc;             // destination of the layout jobs; null: the caller's map is NHWC bf16 already
  int* ctrl;                       // [DC_WORDS], all zero between launches
  int enabled, n_helper_ctas, B, L, H, W;
};


constexpr int R2_MAX_STAGES = 160;

// One entry of the static schedule that the weight-ring TMA thread and the MMA thread walk in
// lock-step with the compute warps' program order.
enum : unsigned char {
  R2F_WAITB = 1,      // MMA thread waits for the B operand(s) of this stage group first
  R2F_COMMIT = 2,     // last entry of a stage group: commit the accumulators to the compute warps
  R2F_CONV = 4,       // marker: the value_proj conv of a layer call runs here (map = conv weights)
  R2F_N32 = 8,        // 32 activation rows (hoisted K|V|ego stage), standard operand layout
  R2F_RANK16 = 16,    // 16-way feature split (hoisted stage): this CTA streams rows [rank*rows, +rows);
                      // every other stage is 4-way: rows [fg*rows*mtiles, +rows*mtiles), fg = rank & 3
};
struct R2Stage {
  const void* w;            // weights packed by launch_pack_sw128: [tile of `rows` rows][k-chunk][row][128 B]
  unsigned short rows;      // weight rows per tile (64; 48 for the hoisted stage)
  unsigned short acc_col;   // TMEM column of tile 0
  unsigned char mtiles;     // tiles per CTA
  unsigned char kchunks;    // K / 64
  unsigned char flags;
  unsigned char bsel;       // B operand: 0 / 1 = the two chain buffers (stage parity), 2 = cls branch
};

struct alignas(16) R2Consts {
  ResLayerC layer[RES_MAX_L];
  const float *b_enc0, *b_enc3, *enc_ln_g, *enc_ln_b, *anchors, *dim_t, *film;   // film [S][L][2D]
  int A, P, Na, F, L, S, H, W, heads, rcap, n_stages, pad0;
  OdoConsts oc;
  float sa_tr, sb_tr;              // sqrt(ac[t_trunc]), sqrt(1 - ac[t_trunc])
  DdimCoef dc[RES_MAX_S];
  // exchange through L2 (library-owned workspace, per scene)
  float *kv, *egov;                 // [B][L][Na][2D], [B][L][D]
  __nv_bfloat16* bev_nhwc;          // [B][H*W][256] working copy (NCHW callers)
  // debug taps (written only when ResCall::dbg is set)
  float *tap_q0, *tap_x1, *tap_regraw;
  R2Stage stages[R2_MAX_STAGES];
};

int res2_smem_bytes();
int res2_engine_init();   // 0 when a 16-CTA cluster of this kernel can be co-scheduled
// pack-time: bf16 [N][K] -> pre-swizzled shared-memory images of (rows x 64) tiles
void launch_pack_sw128(const __nv_bfloat16* W, __nv_bfloat16* out, int N, int K, int rows, cudaStream_t st);
int res2_max_clusters();  // co-resident 16-CTA clusters of the engine on an idle device (after res2_engine_init)
// dense: null or DenseArgs with enabled = 1 (the launch then carries n_helper_ctas / 16 helper clusters)
int launch_res2_forward(const R2Consts* consts_dev, const ResCall& call, int B, cudaStream_t st,
                        const DenseArgs* dense);

}  // namespace ddh
