// Host-side launchers of the ddh kernels.  All launches are asynchronous on `st`.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"

namespace ddh {

// scalar constants of the elementwise stages
struct OdoConsts {
  float lidar_max_x, lidar_max_y;
};
struct DdimCoef {
  float sqrt_ac_t, sqrt_1m_ac_t, sqrt_ac_prev, sqrt_1m_ac_prev;
};

// ---- kernels_simt.cu -------------------------------------------------------------
void launch_simt_gemm(const GemmParams& p, int n_total, cudaStream_t st);
void launch_simt_conv(const GemmParams& p, int B, cudaStream_t st);

void launch_bev_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype, int B, int C,
                        int HW, cudaStream_t st);
void launch_cast_f32_bf16(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t st);
void launch_cast_bf16_f32(const __nv_bfloat16* src, float* dst, size_t n, cudaStream_t st);

void launch_init_img(const float* anchors, const float* noise, float* img, int B, int AP,
                     float sqrt_ac, float sqrt_1m_ac, cudaStream_t st);
void launch_embed(const float* img, float* pts, float* emb32, __nv_bfloat16* emb16, int M, int P,
                  const float* dim_t_dev, cudaStream_t st);
// Value-row reuse across denoise steps (forward_fused): value_proj(bev) of a layer does not depend on
// the denoise step, so the rows of V a step produced are kept ([B][vcap][256] bf16 per layer) and a
// later step only evaluates the pixels that have no row yet.
//   mode 0  off;  1  first step: the scene's unique pixels go to upix as before and the pixel -> slot
//   table is saved;  2  later step: sampled pixels without a row are appended to the cross-scene list
//   new_list (x = pixel index in the batch, y = value row it fills) and get the scene's next slots;
//   3  as 2, but the saved table is ignored: every sampled pixel of the step is evaluated again (validation /
//   worst case of the schedule: option conv_reuse = 2)
struct PlanReuse {
  int mode = 0;
  int keep = 1;                           // write the slot table back (0 in the last step)
  unsigned short* slot_tab = nullptr;     // [B][H*W] slot + 1 of a pixel's value row, 0 = none
  int* slot_cnt = nullptr;                // [B] rows a scene holds
  int2* new_list = nullptr;
  int* new_count = nullptr;               // rows in new_list (zeroed before the forward)
  int vcap = 0;                           // row capacity per scene
  // independent of mode: attention-weight logits hoisted into the chain engine's encoder program
  // (ChainArgs::logit_part; needs q0_spt > 0 and 8 poses): the layer's 8 logits start at logit_off of logit_ld
  const float* logit_part = nullptr;
  int logit_ld = 0, logit_off = 0;
};
void launch_plan(const float* q0, const float* attw_w, const float* attw_b, const float* pts,
                 int* upix, int* nuniq, int* ent_slot, float* ent_w, int* rows_total,
                 unsigned int* need_seg, unsigned int* done_seg, int seg_shift, int nw32, int B, int A,
                 int P, int H, int W, int rcap, OdoConsts oc, cudaStream_t st, int q0_spt = 0,
                 PlanReuse ru = PlanReuse());
// S[scene, a, :] = sum_k w[scene, a, k] * V[scene * vcap + slot[scene, a, k], :]  (bf16 rows, fp32 sum)
void launch_combine_rows(const __nv_bfloat16* V, const int* ent_slot, const float* ent_w,
                         __nv_bfloat16* s16, int B, int A, int ent_per_anchor, int vcap, cudaStream_t st);
// on-demand layout conversion of the BEV segments (seg = 8 or 16 pixels of a row) flagged in todo
void launch_bev_segs_to_nhwc(const void* src, int src_dtype, void* dst, int dst_dtype,
                             const unsigned int* todo, int nw32, int seg, int B, int C, int H, int W,
                             cudaStream_t st, void* dst_lo = nullptr);
void launch_split_tf32(const float* src, float* hi, float* lo, size_t n, cudaStream_t st);
void launch_pack_conv_tf32x2(const float* w, float* dst, int Cout, int Cin, cudaStream_t st);
void launch_combine(const float* V, const int* ent_slot, const float* ent_w, float* s32,
                    __nv_bfloat16* s16, int B, int A, int P, int rcap, cudaStream_t st);
void launch_attn_core(const float* qh, const float* kv, float* o32, __nv_bfloat16* o16, int B,
                      int A, int Na, int heads, cudaStream_t st);
void launch_reg_finish(const float* r2, const float* w4, const float* b4, float* pts, float* img,
                       float* modes, int M, int P, int do_ddim, DdimCoef dc, cudaStream_t st);
void launch_select(const float* scores, const float* modes, float* traj, long long* mode_idx,
                   int B, int A, int P, cudaStream_t st);

// query decoder pieces (kernels_simt.cu)
int launch_mha_small(const float* q, int ldq, const float* kv, int ldkv, float* o32, __nv_bfloat16* o16,
                     int B, int nq, int nk, cudaStream_t st);
void launch_rowdot(const float* x, int ldx, const float* w, const float* b, float* y, int M, int K, int n_out,
                   int rows_per_group, int skip_first, int states, cudaStream_t st);
void launch_broadcast_rows(const float* emb, float* x32, __nv_bfloat16* x16, int rows_per_group, size_t n,
                           cudaStream_t st);

// pack-time helpers
void launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t st);
void launch_pack_conv_f32(const float* w, float* dst, int Cout, int Cin, cudaStream_t st);
void launch_pack_conv_bf16(const float* w, __nv_bfloat16* dst, int Cout, int Cin,
                           cudaStream_t st);
void launch_matvec(const float* W, const float* x, const float* b, float* y, int n_out, int k,
                   int act_in_mish, cudaStream_t st);
void launch_time_sinemb(float* emb, int dim, int timestep, cudaStream_t st);
void launch_pack_hilo(const float* w, __nv_bfloat16* dst, int n_out, int k, cudaStream_t st);

// ---- kernels_producer.cu (cross_bev_feature producer, transfuser_model_v2.py:121-140)
size_t bev_producer_scratch_bytes(int B, int g, int cb);
int launch_bev_producer(const float* tok, const float* map, const float* w, const float* bias,
                        const float* ln_g, const float* ln_b, void* out, int out_bf16, int B, int H, int W,
                        int g, int cb, float* scratch, cudaStream_t st);

// ---- kernels_tc.cu (tcgen05 / TMEM / TMA engine) ----------------------------------
// W is described by a TMA tensor map over a bf16 [N_total][K] matrix (box 64 x 256,
// 128-byte swizzle).  A is bf16 [M][lda] (dense) or gathered from the NHWC bf16 BEV map.
void launch_tc_gemm(const GemmParams& p, const CUtensorMap& wmap, int n_total, cudaStream_t st);
// mode 0: one CTA per scene; 1: persistent, CUDA-core combine; 2: persistent, combine on the tensor
// core (n_anchor <= 64, ent_per_anchor == 32; other shapes fall back to mode 1)
void launch_tc_conv(const GemmParams& p, const CUtensorMap& wmap, int B, cudaStream_t st, int mode = 1);
void launch_tc_convv(const GemmParams& p, const CUtensorMap& wmap, cudaStream_t st);
void launch_tc_conv_tf32(const GemmParams& p, const CUtensorMap& wmap, int B, cudaStream_t st);
int tc_conv_smem_bytes(int A, int ent_per_anchor);
int tc_engine_init();   // sets max dynamic smem attributes; returns cudaError_t as int

}  // namespace ddh
