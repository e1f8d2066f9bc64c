// Resident ("one launch per forward") engine of the ddh planning head for a handful of scenes:
// one 16-CTA thread-block cluster per scene runs the whole TrajectoryHead.forward_test
// (transfuser_model_v2.py:578-641) in ONE kernel.  See kernels_res.cu for the design.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "kernels.h"

namespace ddh {

constexpr int RES_CL = 16;        // CTAs per cluster (non-portable size, opt-in)
constexpr int RES_MAX_L = 4;      // decoder layers
constexpr int RES_MAX_S = 4;      // denoise steps
constexpr int RES_MAX_ITEMS = 160;
constexpr int RES_MAX_B = 24;     // scenes per call served by the one-launch engines

// Tensor maps of the weight matrices (bf16 [N][K], box {64 k, slice rows}, 128-byte swizzle) and
// the order in which a CTA consumes them.  Lives in global memory only (TMA descriptors).
enum ResMapKind { RM_KVEGO = 0, RM_BEV_OUT, RM_Q, RM_ATTN_OUT, RM_FFN0, RM_FFN2, RM_REG0, RM_REG2,
                  RM_CLS0, RM_CLS3, RM_CONV, RM_PER_LAYER };
struct alignas(64) ResMaps {
  CUtensorMap enc0, enc3;
  CUtensorMap layer[RES_MAX_L][RM_PER_LAYER];
};

// One streamed weight item of the linear stages: rows [rank*rows, +rows) of a bf16 [n_total][K]
// matrix, K/64 TMA boxes of rows x 128 B.
struct ResItem {
  const CUtensorMap* map;
  unsigned short rows, kchunks;
  int n_total;
};

struct ResLayerC {
  const float *b_kvego, *b_bev_out, *b_q, *b_attn_out, *b_ffn0, *b_ffn2, *b_reg0, *b_reg2, *b_cls0,
      *b_cls3, *b_conv;
  const float *attw_w, *attw_b, *norm1_g, *norm1_b, *norm2_g, *norm2_b, *norm3_g, *norm3_b;
  const float *cls_ln2_g, *cls_ln2_b, *cls_ln5_g, *cls_ln5_b, *cls6_w, *cls6_b, *reg4_w, *reg4_b;
  const CUtensorMap* conv_map;
};

// Everything the kernel needs besides the per-call buffers; copied to shared memory at kernel
// start.  Workspace pointers address scene 0; scene s adds s * (the per-scene element count).
struct alignas(16) ResConsts {
  ResLayerC layer[RES_MAX_L];
  const float *b_enc0, *b_enc3, *enc_ln_g, *enc_ln_b, *anchors, *dim_t, *film;   // film [S][L][2D]
  int A, P, Na, F, L, S, H, W, heads, rcap, n_items, tiles_max;
  OdoConsts oc;
  float sa_tr, sb_tr;              // sqrt(ac[t_trunc]), sqrt(1 - ac[t_trunc])
  DdimCoef dc[RES_MAX_S];
  // exchange buffers (library-owned workspace)
  __nv_bfloat16 *emb16, *o16, *h16, *r1_16;
  float *e1, *q0, *spart, *x1, *y2, *y3, *c1, *r2, *c2, *regraw, *kv, *egov;
  __nv_bfloat16* bev_nhwc;          // [RES_MAX_B][H*W][256] working copy (NCHW callers)
  ResItem items[RES_MAX_ITEMS];
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
};

int res_smem_bytes();
// returns 0 when a 16-CTA cluster of this kernel can be co-scheduled on the device
int res_engine_init();
int launch_res_forward(const ResConsts* consts_dev, const ResCall& call, int B, cudaStream_t st);

}  // namespace ddh
