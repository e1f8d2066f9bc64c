/*
 * ddh.h — C ABI of the B200-native DiffusionDrive planning head ("ddh").
 *
 * The reference (seulbinHwang/DiffusionDrive) is pure Python and has no FFI layer
 * of its own; the boundary this library replaces is the Python method
 *
 *   TrajectoryHead.forward_test(ego_query, agents_query, bev_feature,
 *                               bev_spatial_shape, status_encoding, global_img)
 *   navsim/agents/diffusiondrive/transfuser_model_v2.py:578-641
 *
 * reached through TrajectoryHead.forward (:502-518), called by
 * V2TransfuserModel.forward (:150-156).  Each entry point below cites the
 * reference lines whose work it takes over.  The Python binding a maintainer
 * adds on the reference side is shown in INTEGRATION.md; the one shipped in this
 * repo is diffusiondrive_b200/_lib.py (ctypes).
 *
 * Conventions
 *   - plain C: no C++ types, no torch types, pointers + sizes only;
 *   - every function returning int returns DDH_OK (0) or a negative ddh_status;
 *     the text of the last failure is available from ddh_last_error();
 *   - all tensors are dense row-major ("contiguous" in torch terms);
 *   - the caller owns every buffer it passes in; the library borrows device
 *     pointers for the duration of one call and never frees them;
 *   - the library owns packed weights, workspace and TMA tensor maps, all
 *     released by ddh_destroy();
 *   - ddh_forward() is asynchronous on the given CUDA stream and never
 *     synchronises the host (workspace growth on a larger batch is the one
 *     exception: call ddh_reserve() first to avoid it);
 *   - one in-flight call per handle; different handles are independent;
 *     cudaSetDevice() is the caller's job;
 *   - there is no CPU fallback: without a CUDA device every compute entry
 *     point fails with DDH_ERR_CUDA.
 */
#ifndef DDH_H_
#define DDH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDH_ABI_VERSION 1

#if defined(__GNUC__)
#define DDH_API __attribute__((visibility("default")))
#else
#define DDH_API
#endif

typedef enum ddh_status {
  DDH_OK = 0,
  DDH_ERR_BAD_ARG = -1,      /* null pointer, bad enum, B <= 0 ...                 */
  DDH_ERR_UNSUPPORTED = -2,  /* shape outside what the kernels are built for       */
  DDH_ERR_NOT_PACKED = -3,   /* ddh_forward before ddh_pack_weights                */
  DDH_ERR_CUDA = -4,         /* CUDA runtime / driver error (text in last_error)   */
  DDH_ERR_ALIGNMENT = -5,    /* a pointer is not aligned as documented             */
  DDH_ERR_NOMEM = -6
} ddh_status;

/* arithmetic of the GEMM-shaped stages (everything between them is fp32) */
typedef enum ddh_precision {
  DDH_PREC_FP32 = 0,  /* CUDA-core fp32 FMA everywhere: <= 1e-4 m vs the reference  */
  DDH_PREC_BF16 = 1   /* bf16 operands on tcgen05 tensor cores, fp32 accumulate     */
} ddh_precision;

typedef enum ddh_dtype { DDH_F32 = 0, DDH_BF16 = 1 } ddh_dtype;

/* layout of bev_feature: the reference hands the head NCHW
 * (transfuser_model_v2.py:138-140); one line earlier it holds NHWC (:136-137). */
typedef enum ddh_layout { DDH_NCHW = 0, DDH_NHWC = 1 } ddh_layout;

/* Shape of the path.  Defaults of the reference in brackets. */
typedef struct ddh_shape {
  int32_t num_anchors;     /* A  [20]  plan_anchor.shape[0], :453-458                */
  int32_t num_poses;       /* P  [8]   must be 8                                     */
  int32_t d_model;         /* D  [256] must be 256                                   */
  int32_t d_ffn;           /* F  [1024] multiple of 256                              */
  int32_t num_heads;       /*    [8]   head_dim must be 32                           */
  int32_t num_agents;      /* Na [30]  <= 32                                         */
  int32_t bev_channels;    /* C  [256] must be 256 (in_bev_dims, :314)               */
  int32_t bev_h;           /* H  [64]                                                */
  int32_t bev_w;           /* W  [64]  H*W <= 65535                                  */
  int32_t num_layers;      /* L  [2]   decoder depth, literal at :476                */
  int32_t num_steps;       /* S  [2]   step_num, literal at :581                     */
  int32_t trunc_timestep;  /*    [8]   literal at :594                               */
  float lidar_max_x;       /*    [32]  transfuser_config.py:29-32                    */
  float lidar_max_y;       /*    [32]                                                */
} ddh_shape;

/* Device pointers to the fp32 parameters of one CustomTransformerDecoderLayer
 * (transfuser_model_v2.py:297-341), in torch's native layouts. */
typedef struct ddh_layer_weights {
  /* cross_bev_attention (modules/blocks.py:49-78) */
  const float *bev_attw_w, *bev_attw_b;   /* attention_weights  [P,D], [P]           */
  const float *bev_out_w, *bev_out_b;     /* output_proj        [D,D], [D]           */
  const float *bev_conv_w, *bev_conv_b;   /* value_proj.0       [256,C,3,3], [256]   */
  /* cross_agent_attention / cross_ego_attention (nn.MultiheadAttention) */
  const float *agent_in_w, *agent_in_b;   /* in_proj q|k|v      [3D,D], [3D]         */
  const float *agent_out_w, *agent_out_b; /* out_proj           [D,D], [D]           */
  const float *ego_in_w, *ego_in_b;
  const float *ego_out_w, *ego_out_b;
  const float *ffn0_w, *ffn0_b;           /* ffn.0              [F,D], [F]           */
  const float *ffn2_w, *ffn2_b;           /* ffn.2              [D,F], [D]           */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b, *norm3_w, *norm3_b; /* [D]     */
  const float *film_w, *film_b;           /* time_modulation.scale_shift_mlp.1 [2D,D]*/
  /* task_decoder (DiffMotionPlanningRefinementModule, :208-256) */
  const float *cls0_w, *cls0_b, *cls_ln2_w, *cls_ln2_b;
  const float *cls3_w, *cls3_b, *cls_ln5_w, *cls_ln5_b;
  const float *cls6_w, *cls6_b;           /* [1,D], [1]                              */
  const float *reg0_w, *reg0_b, *reg2_w, *reg2_b;
  const float *reg4_w, *reg4_b;           /* [3P,D], [3P]                            */
} ddh_layer_weights;

/* Device pointers to all fp32 parameters of TrajectoryHead (:455-476). */
typedef struct ddh_weight_ptrs {
  const float *plan_anchor;                /* [A,P,2] metres                         */
  const float *enc0_w, *enc0_b;            /* plan_anchor_encoder.0 [D,512]          */
  const float *enc_ln_w, *enc_ln_b;        /* plan_anchor_encoder.2 [D]              */
  const float *enc3_w, *enc3_b;            /* plan_anchor_encoder.3 [D,D]            */
  const float *time1_w, *time1_b;          /* time_mlp.1 [4D,D]                      */
  const float *time3_w, *time3_b;          /* time_mlp.3 [D,4D]                      */
  const ddh_layer_weights *layers;         /* HOST array of num_layers entries       */
} ddh_weight_ptrs;

typedef struct ddh_handle ddh_handle;

/* Library / ABI version (DDH_ABI_VERSION of the build). */
DDH_API int ddh_abi_version(void);

/* Build-time facts: "sm_100a", kernel families present. Static string. */
DDH_API const char *ddh_build_info(void);

/* Replaces TrajectoryHead.__init__ shape bookkeeping (:431-478). No device work. */
DDH_API int ddh_create(const ddh_shape *shape, ddh_handle **out);
DDH_API void ddh_destroy(ddh_handle *h);

/* Text of the last error on this handle (h may be NULL for create-time errors). */
DDH_API const char *ddh_last_error(const ddh_handle *h);

/* Override the DDIM alphas_cumprod table (n >= 21 floats, host pointer).  By default the
 * library computes DDIMScheduler(1000, "scaled_linear") itself (diffusers, :447-451);
 * a Python caller passes torch's own table so the constants are bit-identical. */
DDH_API int ddh_set_alphas_cumprod(ddh_handle *h, const float *table_host, int n);
/* Copy the table in use to the caller (host), for tests. */
DDH_API int ddh_get_alphas_cumprod(const ddh_handle *h, float *table_host, int n);

/* Copy + repack the parameters into the layouts the kernels read, fold what is
 * weight-only (time_mlp + FiLM vectors :463-468,:276-294; ego out_proj o v_proj :322-327)
 * and build the TMA tensor maps.  Must be called again after any change of the
 * source parameters (load_state_dict, .to(), ...).  Asynchronous on `stream`. */
DDH_API int ddh_pack_weights(ddh_handle *h, const ddh_weight_ptrs *w, int precision, void *stream);

/* Workspace the library will hold for a batch of B scenes, in bytes. */
DDH_API size_t ddh_workspace_bytes(const ddh_handle *h, int B);
/* Pre-allocate for up to B scenes (synchronises; optional). */
DDH_API int ddh_reserve(ddh_handle *h, int B);

/* TrajectoryHead.forward_test (:578-641) for B scenes, device buffers.
 *   ego     [B,1,D] f32          agents [B,Na,D] f32
 *   bev     [B,C,H,W] (NCHW) or [B,H,W,C] (NHWC), f32 or bf16, 16-byte aligned
 *   noise   [B,A,P,2] f32        host-generated N(0,1), replaces torch.randn (:593)
 * outputs (any may be NULL):
 *   out_traj     [B,P,3] f32     "trajectory" (:641)
 *   out_modes    [B,A,P,3] f32   poses_reg of the last step (:630)
 *   out_scores   [B,A] f32       poses_cls logits of the last step (:631)
 *   out_mode_idx [B] int64       argmax (:637)
 */
DDH_API int ddh_forward(ddh_handle *h, const float *ego, const float *agents, const void *bev,
                int bev_dtype, int bev_layout, const float *noise, float *out_traj,
                float *out_modes, float *out_scores, int64_t *out_mode_idx, int B,
                void *stream);

/* Same call with HOST buffers (pinned memory recommended): copies the inputs to the
 * device (a pinned NCHW map is read in place, see option "host_zero_copy"), runs ddh_forward,
 * copies the outputs back and synchronises `stream`.
 * This is the end-to-end path a CPU-resident caller (abstract_agent.py:65-86) uses. */
DDH_API int ddh_forward_host(ddh_handle *h, const float *ego, const float *agents, const void *bev,
                     int bev_dtype, int bev_layout, const float *noise, float *out_traj,
                     float *out_modes, float *out_scores, int64_t *out_mode_idx, int B,
                     void *stream);

/* cross_bev_feature producer: the stage of V2TransfuserModel.forward
 * that builds the head's BEV input (transfuser_model_v2.py:121-140 with bev_proj, :96):
 *   out[b,y,x,:] = LayerNorm(ReLU(W . cat(bilinear_up(keyval tokens)[b,:,y,x], bev_map[b,:,y,x]) + bias))
 *   keyval_tokens [B, grid*grid, 256] f32   keyval[:, :-1] of :115-119 (token t = y*grid + x)
 *   bev_map       [B, Cb, H, W] f32         bev_feature_upscale, NCHW (:110)
 *   weight [256, 256+Cb], bias / ln_weight / ln_bias [256]   bev_proj.{0,2}
 *   out           [B, H, W, 256] f32 or bf16 (NHWC: what ddh_forward takes with DDH_NHWC; the
 *                 reference's permute to NCHW, :138-140, is not materialised)
 *   scratch       ddh_bev_producer_scratch_bytes(B, grid, Cb) bytes of device memory
 * Device pointers, 16-byte aligned; asynchronous on `stream`; W % 32 == 0. */
DDH_API size_t ddh_bev_producer_scratch_bytes(int B, int grid, int bev_channels);
DDH_API int ddh_bev_producer(const float *keyval_tokens, const float *bev_map, const float *weight,
                     const float *bias, const float *ln_weight, const float *ln_bias, void *out,
                     int out_dtype, int B, int H, int W, int grid, int bev_channels, void *scratch,
                     void *stream);

/* ---- query decoder + AgentHead ------------------------------------
 * V2TransfuserModel.forward :141-146,159-160: query = _query_embedding.weight repeated per scene,
 * query_out = nn.TransformerDecoder(3 x nn.TransformerDecoderLayer(d 256, 8 heads, ffn 1024, post-norm,
 * ReLU), norm=None)(query, keyval); agent_states / agent_labels = AgentHead(query_out[:, 1:]) (:165-205).
 * Same engines as the head: DDH_PREC_FP32 = CUDA-core GEMMs, DDH_PREC_BF16 = tcgen05 GEMMs; the
 * attention (<= 32 queries) and LayerNorms are fp32 in both. */
typedef struct ddh_qdec_shape {
  int32_t num_queries;   /* [31] 1 ego + num_bounding_boxes, <= 32                    */
  int32_t num_keys;      /* [65] 8x8 BEV tokens + status token, <= 96                 */
  int32_t d_model;       /* [256] must be 256                                         */
  int32_t d_ffn;         /* [1024] multiple of 256                                    */
  int32_t num_heads;     /* [8] head_dim must be 32                                   */
  int32_t num_layers;    /* [3] tf_num_layers                                         */
} ddh_qdec_shape;

typedef struct ddh_qdec_layer_weights {   /* nn.TransformerDecoderLayer, torch layouts, device f32 */
  const float *self_in_w, *self_in_b;     /* self_attn.in_proj       [3D,D], [3D]      */
  const float *self_out_w, *self_out_b;   /* self_attn.out_proj      [D,D], [D]        */
  const float *cross_in_w, *cross_in_b;   /* multihead_attn.in_proj  [3D,D], [3D]      */
  const float *cross_out_w, *cross_out_b; /* multihead_attn.out_proj [D,D], [D]        */
  const float *lin1_w, *lin1_b;           /* linear1 [F,D], [F]                        */
  const float *lin2_w, *lin2_b;           /* linear2 [D,F], [D]                        */
  const float *norm1_w, *norm1_b, *norm2_w, *norm2_b, *norm3_w, *norm3_b;
} ddh_qdec_layer_weights;

typedef struct ddh_qdec_weight_ptrs {
  const float *query_embedding;           /* _query_embedding.weight [Q,D]             */
  const ddh_qdec_layer_weights *layers;   /* HOST array of num_layers entries          */
  const float *states0_w, *states0_b;     /* _agent_head._mlp_states.0 [F,D], [F]      */
  const float *states2_w, *states2_b;     /* _agent_head._mlp_states.2 [5,F], [5]      */
  const float *label_w, *label_b;         /* _agent_head._mlp_label.0  [1,D], [1]      */
} ddh_qdec_weight_ptrs;

typedef struct ddh_qdec ddh_qdec;
DDH_API int ddh_qdec_create(const ddh_qdec_shape *shape, ddh_qdec **out);
DDH_API void ddh_qdec_destroy(ddh_qdec *q);
DDH_API const char *ddh_qdec_last_error(const ddh_qdec *q);
DDH_API int ddh_qdec_pack_weights(ddh_qdec *q, const ddh_qdec_weight_ptrs *w, int precision, void *stream);
/* keyval [B,Nk,D] f32 (device) -> query_out [B,Q,D], agent_states [B,Q-1,5], agent_labels [B,Q-1]
 * (any output may be NULL); asynchronous on `stream` (workspace growth on a larger B synchronises). */
DDH_API int ddh_qdec_forward(ddh_qdec *q, const float *keyval, float *query_out, float *agent_states,
                     float *agent_labels, int B, void *stream);

/* Number of kernel launches issued by the last ddh_forward on this handle. */
DDH_API int ddh_last_launch_count(const ddh_handle *h);

/* Scene-chunk concurrency of ddh_forward: a call of B >= chunks*min_chunk_scenes scenes is cut
 * into `chunks` contiguous scene blocks issued alternately on the caller's stream and one
 * internal stream (fork/join with events, CUDA-graph capturable), so that one block's HBM-bound
 * BEV layout pass overlaps another block's tensor-bound stages.  Results do not depend on the
 * setting (scenes are independent).  Default: chunks = 1 (everything on the caller's stream);
 * 2 chunks gain ~1 % at 4096 scenes on a B200. */
DDH_API int ddh_set_concurrency(ddh_handle *h, int chunks, int min_chunk_scenes);

/* Execution options (engine selection and debugging), by name; none of them changes results
 * beyond the engines' documented tolerances.  Engine selection is fixed when the weights are
 * packed: after changing "chain_engine" or "resident_engine" ddh_forward fails with
 * DDH_ERR_NOT_PACKED until ddh_pack_weights is called again.
 *   "resident_engine"    1  B <= 24, bf16: whole forward as one launch (kernels_res2.cu)
 *   "dense_conv"         1  resident engine, B <= value (0..2): value_proj + ReLU of the WHOLE map runs on
 *                           helper clusters of the same launch (TMA-fed tcgen05 implicit GEMM) under the
 *                           embedding / encoder, and the scene cluster only gathers bilinear corners;
 *                           0: on-demand conv at the unique sampled pixels for every batch size
 *   "chain_engine"       1  bf16: scene-tile chain kernel per decoder-layer call (kernels_chain.cu);
 *                           0: one tcgen05 GEMM launch per Linear (kernels_tc.cu)
 *   "lazy_layout"        1  NCHW input: convert BEV segments on demand; 0: whole map up front
 *   "layout_segment"     8  pixels per on-demand layout segment (8 or 16)
 *   "host_zero_copy"     1  ddh_forward_host: a pinned NCHW bev_feature is read in place across PCIe by
 *                           the on-demand layout pass (only the needed segments) instead of copied whole
 *   "host_segment"      64  pixels per segment when the map is read in place from pinned host memory
 *                           (16, 32 or 64: 256-byte PCIe reads measured fastest)
 *   "persistent_conv"    2  value_proj conv as one persistent CTA per SM: 2 = bilinear x attention combine
 *                           on the tensor core (tc_conv3_kernel), 1 = CUDA-core combine
 *                           (tc_conv2_kernel); 0: one CTA per scene (tc_conv_kernel)
 *   "fp32_tensor_conv"   1  fp32 engine: value_proj runs on the tensor core as 3xTF32 (fp32 operands split in
 *                           a 10-bit-mantissa high part and an exact remainder, three tcgen05 kind::tf32
 *                           products, fp32 accumulate: fp32-level accuracy); 0: CUDA-core conv.  Fixed at
 *                           pack time like the engine options
 *   "conv_dynamic"       1  tc_conv3_kernel deals scenes to its persistent CTAs on demand (global counter) instead
 *                           of round-robin: removes the end-of-launch tail of unequal scenes
 *   "conv_reuse"         1  chain engine, >= 2 denoise steps: value_proj(bev) of a layer (blocks.py:114) does not
 *                           depend on the denoise step, so the value rows the first step evaluated are kept
 *                           (bf16 [layers][B][min(H*W, steps * rows)][256], allocated only when it takes less
 *                           than a third of the free device memory) and later steps evaluate only the pixels
 *                           no earlier step sampled (cross-scene tc_conv_kernel<true>) and combine over the
 *                           kept rows (combine_rows_kernel); exact up to the bf16 rounding of the combine
 *                           weights; 0: every step runs the full on-demand conv; 2: validation / worst case:
 *                           the later steps use the same kernels but treat every sampled pixel as new.
 *                           Switching between 0 and non-zero frees the workspace (synchronises)
 *   "chain_timeline"    -1  index (step * layers + layer) of the chain launch that stamps clock64
 *                           into the "dbg" tap (CTA 0, second tile)
 *   "debug_taps"         0  keep fp32 copies of intermediate activations for ddh_debug_copy
 *   "timeline_gemm"     -1  index of the dense GEMM launch to stamp (DDH_TIMELINE builds) */
DDH_API int ddh_set_option(ddh_handle *h, const char *name, int value);

/* Optional per-stage device timing: when on, ddh_forward brackets each stage with CUDA events
 * on the caller's stream.  ddh_get_profile synchronises and returns the summed duration and
 * the number of timed spans of one stage of the LAST forward.  Stages: "bev_layout",
 * "hoist_kv_ego", "init", "embed_encode", "plan", "conv", "combine", "gemm_chain",
 * "attn_core", "reg_finish", "select", "conv_new" (value rows of later denoise steps, conv_reuse). */
DDH_API int ddh_set_profiling(ddh_handle *h, int on);
DDH_API int ddh_get_profile(ddh_handle *h, const char *stage, float *total_ms, int *spans);

/* Test hook: copy a named internal buffer of the last forward to the host (synchronises).
 * Returns the number of bytes copied (>= 0) or a negative ddh_status.  Names are listed in
 * DESIGN.md ("debug taps"); sizes depend on B. */
DDH_API long long ddh_debug_copy(ddh_handle *h, const char *name, void *host_dst, size_t max_bytes);

/* Test hook: C[M,N] = A[M,K] * W[N,K]^T + bias on the GEMM engine of the given
 * precision (device pointers, f32 in/out), to validate the tensor-core tiles in isolation. */
DDH_API int ddh_test_gemm(ddh_handle *h, const float *A, const float *W, const float *bias, float *C,
                  int M, int N, int K, int precision, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DDH_H_ */
