// C ABI of the ddh planning head (include/ddh.h): handle, weight packing, workspace and the
// launch sequence of TrajectoryHead.forward_test (transfuser_model_v2.py:578-641).
//
// Algebra applied on top of the reference (each verified against the live module):
//   * agent K|V projections are step-invariant -> computed once per (scene, layer);
//   * cross_ego_attention has ONE key, so softmax == 1 and the block collapses to the vector
//     out_proj(v_proj(ego)); the two linears are folded into one at pack time;
//   * time_mlp / FiLM vectors are identical for every scene -> computed at pack time;
//   * value_proj (3x3 conv + ReLU) is evaluated only at the unique pixels grid_sample reads;
//   * the DDIM update of the final step is dead (never read) and skipped, as are the cls
//     branches of every decoder call but the last (only poses_cls_list[-1] of the last step
//     is read, :630-631).
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/ddh.h"
#include "kernels.h"
#include "kernels_chain.h"
#include "kernels_res2.h"

using namespace ddh;

namespace {

struct PackedLinear {
  float* wt32 = nullptr;          // [K][N] fp32 (SIMT engine)
  __nv_bfloat16* w16 = nullptr;   // [N][K] bf16 (tensor engine)
  CUtensorMap map;
  CUtensorMap map64;              // conv only: 64-row box for the small-batch kernel
  float* w2_32 = nullptr;         // conv only, fp32 engine: [N][w_hi (K) | w_lo (K)] for the 3xTF32 conv
  CUtensorMap map_tf32;           // over w2_32: box {32 k, 256 rows}
  float* bias = nullptr;          // [N]
  int N = 0, K = 0;
};

struct PackedLayer {
  PackedLinear conv, bev_out, q, kv, attn_out, ego, ffn0, ffn2, cls0, cls3, reg0, reg2;
  float *attw_w = nullptr, *attw_b = nullptr;
  float *norm1_g = nullptr, *norm1_b = nullptr, *norm2_g = nullptr, *norm2_b = nullptr;
  float *norm3_g = nullptr, *norm3_b = nullptr;
  float *cls_ln2_g = nullptr, *cls_ln2_b = nullptr, *cls_ln5_g = nullptr, *cls_ln5_b = nullptr;
  float *cls6_w = nullptr, *cls6_b = nullptr, *reg4_w = nullptr, *reg4_b = nullptr;
  __nv_bfloat16* reg4_hl = nullptr;   // [64][256]: bf16 hi rows 0.., lo rows 32.. (chain engine tail)
  CUtensorMap reg4_map;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

std::string g_create_error;

}  // namespace

struct ddh_handle {
  ddh_shape shp;
  int precision = -1;
  bool packed = false;
  std::string err;
  std::vector<void*> owned_w;    // packed weights
  std::vector<void*> owned_ws;   // workspace
  std::vector<float> ac;         // alphas_cumprod (host)
  std::vector<int> roll;         // denoise timesteps
  EncodeTiledFn encode = nullptr;
  bool tc_ready = false;

  // packed globals
  PackedLinear enc0, enc3;
  float *enc_ln_g = nullptr, *enc_ln_b = nullptr;
  float* anchors = nullptr;
  float* dim_t = nullptr;
  float* film = nullptr;         // [S][L][2D]
  std::vector<PackedLayer> layers;

  // workspace
  int cap_B = 0;
  int last_B = 0;
  int rcap = 0;
  void* bev_nhwc = nullptr;
  float *img = nullptr, *pts = nullptr;
  float* emb32 = nullptr; __nv_bfloat16* emb16 = nullptr;
  float* e1_32 = nullptr; __nv_bfloat16* e1_16 = nullptr;
  float* q0_32 = nullptr; __nv_bfloat16* q0_16 = nullptr;
  float* kv32 = nullptr;   // [L][B*Na][2D]
  float* egov = nullptr;   // [L][B][D]
  __nv_bfloat16 *agents16 = nullptr, *ego16 = nullptr;
  int *upix = nullptr, *nuniq = nullptr, *ent_slot = nullptr;
  float* ent_w = nullptr;
  float* V = nullptr;
  float* s32 = nullptr; __nv_bfloat16* s16 = nullptr;
  float* x1_32 = nullptr; __nv_bfloat16* x1_16 = nullptr;
  float* qh32 = nullptr;
  float* o32 = nullptr; __nv_bfloat16* o16 = nullptr;
  float* x2_32 = nullptr; __nv_bfloat16* x2_16 = nullptr;
  float* h32 = nullptr; __nv_bfloat16* h16 = nullptr;
  float* x3_32 = nullptr; __nv_bfloat16* x3_16 = nullptr;
  float* c1_32 = nullptr; __nv_bfloat16* c1_16 = nullptr;
  float* r1_32 = nullptr; __nv_bfloat16* r1_16 = nullptr;
  float* r2_32 = nullptr;
  float *modes_buf = nullptr, *scores_buf = nullptr;
  // host-call staging (ddh_forward_host)
  int host_cap_B = 0;
  size_t host_bev_bytes = 0;
  std::vector<void*> owned_host;
  float *hs_ego = nullptr, *hs_agents = nullptr, *hs_noise = nullptr, *hs_traj = nullptr,
        *hs_modes = nullptr, *hs_scores = nullptr;
  void* hs_bev = nullptr;
  long long* hs_idx = nullptr;

  // anchor-resident engine (kernels_res2.cu): same one-launch contract, activations stay on-chip
  int res_mode = 2;                            // option "resident_engine": 2 on, 0 off
  bool res2_ok = false;
  std::vector<void*> owned_res2;
  R2Consts* res2_consts = nullptr;
  R2Consts res2_host;
  // dense mode of the resident engine (whole-map value_proj on helper clusters, kernels_res2.h)
  int dense_max_b = 1;                         // option "dense_conv": largest batch served this way (0: off)
  bool dense_ok = false;
  DenseArgs dense_host;
  const void* dense_amap_base = nullptr;       // base address the A tensor map currently describes
  // scene-tile chain engine (kernels_chain.cu): B > RES_MAX_B in bf16 mode
  int chain_enabled = 1;
  bool chain_ok = false;
  int chain_spt = 0;                           // scenes per 128-row tile
  __nv_bfloat16* kv16 = nullptr;               // [L][B*Na][CH_KV_LD]
  float* q0t = nullptr;                        // [tiles][64][128] float4
  CUtensorMap smap;                            // S16 [B*A][256], box {64, 128}
  std::vector<ChainArgs> chain_prog;           // [S] encoder programs, then [S][L] layer programs
  std::map<std::string, std::pair<const void*, size_t>> taps;
  int launches = 0;
  // optional per-stage device timing (ddh_set_profiling)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  std::vector<std::pair<int, int>> ev_spans;   // (stage id, index of the begin event)
  int ev_used = 0;
  int* conv_rows = nullptr;                    // [S*L] unique value_proj rows per conv launch
  int* conv_sched = nullptr;                   // scene counter of the persistent conv's dynamic scene queue
  int fp32_tensor_conv = 1;                    // option "fp32_tensor_conv": fp32 engine runs value_proj as 3xTF32 on the tensor core
  bool tf32_conv_ok = false;                   // packed for it (fp32 precision, <= 64 anchors x 32 entries, smem fits)
  void* bev_nhwc_lo = nullptr;                 // low-order plane of the fp32 NHWC working copy
  int conv_dynamic = 1;                        // option "conv_dynamic": 1 scenes are dealt to the conv CTAs on demand, 0 round-robin
  // value rows kept across denoise steps (PlanReuse, kernels.h): value_proj(bev) of a layer is the same
  // in every step, so steps after the first evaluate only the pixels no earlier step sampled
  int conv_reuse = 1;                          // option "conv_reuse"
  __nv_bfloat16* vkeep = nullptr;              // [L][B * vcap][256] value rows
  int vcap = 0;                                // rows per scene and layer: min(H*W, steps * rcap)
  unsigned short* slot_tab = nullptr;          // [L][B][H*W] pixel -> slot + 1
  int* slot_cnt = nullptr;                     // [L][B]
  int2* new_list = nullptr;                    // [B * rcap] rows of the current call without a kept value
  int* new_count = nullptr;                    // [S*L] their number per conv call
  // attention-weight logits hoisted into the encoder program of the chain engine (ChainArgs::logit_part)
  float* attw_all = nullptr;                   // [L][P][256] the layers' attention_weights.weight, contiguous
  float* logit_part = nullptr;                 // [tiles][2][128][CH_LOGITS]
  bool logits_hoisted = false;
  unsigned int* need_seg = nullptr;            // [B][seg_nw32] BEV segments (+halo) the coming conv call reads
  unsigned int* done_seg = nullptr;            // [B][seg_nw32] BEV segments already converted to NHWC
  int lazy_layout = 1;                         // convert BEV segments on demand (NCHW input)
  int seg_px = 8;                              // pixels per segment (8 or 16), option "layout_segment"
  int seg_nw32 = 0;                            // 32-bit mask words per scene (0: map too large, eager layout)
  int host_seg_px = 64;                        // option "host_segment": 16 / 32 / 64 pixels when reading a host map
  int seg_px_call = 8;                         // granularity of the current call (pick_segments)
  bool host_map_call = false;                  // the current call reads a pinned host map in place
  int host_zero_copy = 1;                      // option "host_zero_copy": ddh_forward_host reads pinned maps in place
  int persistent_conv = 2;                     // option "persistent_conv": 2 tc_conv3_kernel (combine on the tensor core),
                                               // 1 tc_conv2_kernel (CUDA-core combine), 0 one CTA per scene
  int chain_timeline = -1;
  int conv_timeline = -1;                      // option "conv_timeline": index of the conv launch to stamp                     // option "chain_timeline": index of the chain launch to stamp
  bool profiling_eager = false;
  bool debug_taps = false;                     // env DDH_DEBUG_TAPS=1: keep fp32 copies of x2/x3 in bf16 mode
  long long* dbg = nullptr;                    // timeline stamps (DDH_TIMELINE builds)
  int tl_gemm = -1;                            // env DDH_TIMELINE_GEMM=<launch index>: stamp that dense GEMM
  // scene-chunk concurrency (ddh_set_concurrency)
  int chunks = 1;
  int min_chunk_scenes = 512;
  cudaStream_t aux_stream = nullptr;
  std::vector<cudaEvent_t> sync_events;
};

namespace {

const char* const kStageNames[] = {"bev_layout", "hoist_kv_ego", "embed_encode", "plan", "conv",
                                   "combine", "gemm_chain", "attn_core", "reg_finish", "select",
                                   "init", "conv_new"};
constexpr int kNumStages = sizeof(kStageNames) / sizeof(kStageNames[0]);
enum { ST_BEV = 0, ST_HOIST, ST_EMBED, ST_PLAN, ST_CONV, ST_COMBINE, ST_GEMM, ST_ATTN, ST_REG,
       ST_SELECT, ST_INIT, ST_CONV_NEW };

struct ProfSpan {
  ddh_handle* h;
  cudaStream_t st;
  bool on;
  ProfSpan(ddh_handle* h_, int stage, cudaStream_t st_);
  ~ProfSpan();
};

int fail(ddh_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

#define CU_TRY(h, expr)                                                              \
  do {                                                                               \
    cudaError_t e__ = (expr);                                                        \
    if (e__ != cudaSuccess)                                                          \
      return fail(h, DDH_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

ProfSpan::ProfSpan(ddh_handle* h_, int stage, cudaStream_t st_) : h(h_), st(st_), on(h_->profiling) {
  if (!on) return;
  while ((int)h->ev_pool.size() < h->ev_used + 2) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->ev_pool.push_back(e);
  }
  h->ev_spans.push_back({stage, h->ev_used});
  cudaEventRecord(h->ev_pool[h->ev_used], st);
  h->ev_used += 2;
}
ProfSpan::~ProfSpan() {
  if (on) cudaEventRecord(h->ev_pool[h->ev_spans.back().second + 1], st);
}

template <typename T>
int dev_alloc(ddh_handle* h, std::vector<void*>& owner, T** out, size_t count) {
  void* p = nullptr;
  const size_t bytes = (count * sizeof(T) + 255) / 256 * 256;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 256);
  if (e != cudaSuccess)
    return fail(h, DDH_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  owner.push_back(p);
  *out = reinterpret_cast<T*>(p);
  return DDH_OK;
}

void free_all(std::vector<void*>& owner) {
  for (void* p : owner) cudaFree(p);
  owner.clear();
}

// DDIMScheduler(num_train_timesteps=1000, beta_schedule="scaled_linear") table, in fp32 and in
// torch's order of operations: linspace(sqrt(b0), sqrt(b1), 1000, f32)**2 -> cumprod(1-beta).
void default_alphas_cumprod(std::vector<float>& ac) {
  const int n = 1000;
  ac.resize(n);
  const float start = (float)sqrt(1e-4), end = (float)sqrt(0.02);
  const float step = (end - start) / (float)(n - 1);
  float prod = 1.0f;
  for (int i = 0; i < n; ++i) {
    const float v = (i < n / 2) ? (start + step * (float)i) : (end - step * (float)(n - 1 - i));
    const float beta = v * v;
    prod = prod * (1.0f - beta);
    ac[i] = prod;
  }
}

// roll_timesteps = (arange(S) * (20 / S)).round()[::-1]  (numpy round-half-even), :585-588
void make_roll(int S, std::vector<int>& roll) {
  roll.resize(S);
  const double ratio = 20.0 / S;
  for (int i = 0; i < S; ++i) roll[S - 1 - i] = (int)nearbyint(i * ratio);
}

// TMA tensor map over a bf16 [N][K] matrix: box {64 k, box_rows}, 128-byte swizzle.
int encode_wmap(ddh_handle* h, CUtensorMap* out, void* w16, int N, int K, int box_rows) {
  if (!h->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(h, DDH_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)N};
  const cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w16, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, DDH_ERR_CUDA, "cuTensorMapEncodeTiled failed, CUresult " + std::to_string((int)r));
  return DDH_OK;
}

// TMA tensor map over an NHWC bf16 feature map [B][H][W][256]: box {64 channels, 64 pixels, 2 rows,
// 1 scene}, 128-byte swizzle; coordinates outside the map read as zeros (the conv's padding).
int encode_nhwc_map(ddh_handle* h, CUtensorMap* out, const void* base, int B, int H, int W) {
  const cuuint64_t gdim[4] = {256, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {512, (cuuint64_t)W * 512, (cuuint64_t)H * W * 512};
  const cuuint32_t box[4] = {64, 64, 2, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = h->encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, DDH_ERR_CUDA, "cuTensorMapEncodeTiled (NHWC map) failed, CUresult " + std::to_string((int)r));
  return DDH_OK;
}

// TMA tensor map over an fp32 [N][K2] matrix: box {32 k (128 bytes), 256 rows}, 128-byte swizzle.
int encode_wmap_f32(ddh_handle* h, CUtensorMap* out, void* w32, int N, int K2) {
  if (!h->encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return fail(h, DDH_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    h->encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)K2, (cuuint64_t)N};
  const cuuint64_t gstride[1] = {(cuuint64_t)K2 * 4};
  const cuuint32_t box[2] = {32, 256};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, w32, gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, DDH_ERR_CUDA, "cuTensorMapEncodeTiled (fp32) failed, CUresult " + std::to_string((int)r));
  return DDH_OK;
}

int make_wmap(ddh_handle* h, PackedLinear& L, int box_rows = 256) {
  return encode_wmap(h, box_rows == 256 ? &L.map : &L.map64, L.w16, L.N, L.K, box_rows);
}

int copy_vec(ddh_handle* h, float** dst, const float* src, size_t n, cudaStream_t st) {
  int rc = dev_alloc(h, h->owned_w, dst, n);
  if (rc) return rc;
  CU_TRY(h, cudaMemcpyAsync(*dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return DDH_OK;
}

// Pack one torch Linear ([N][K] fp32 + bias[N]) for the engine of `precision`.
int pack_linear(ddh_handle* h, PackedLinear& L, const float* w, const float* b, int N, int K,
                cudaStream_t st) {
  L.N = N;
  L.K = K;
  int rc;
  if (b) { rc = copy_vec(h, &L.bias, b, N, st); if (rc) return rc; }
  if (h->precision == DDH_PREC_FP32) {
    rc = dev_alloc(h, h->owned_w, &L.wt32, (size_t)N * K);
    if (rc) return rc;
    launch_transpose_f32(w, L.wt32, N, K, st);
  } else {
    rc = dev_alloc(h, h->owned_w, &L.w16, (size_t)N * K);
    if (rc) return rc;
    launch_cast_f32_bf16(w, L.w16, (size_t)N * K, st);
    rc = make_wmap(h, L);
    if (rc) return rc;
  }
  return DDH_OK;
}

int run_gemm(ddh_handle* h, const PackedLinear& L, const float* a32, const __nv_bfloat16* a16,
             int lda, int M, const RowEpi& epi, cudaStream_t st) {
  GemmParams p;
  p.lda = lda;
  p.M = M;
  p.K = L.K;
  p.epi = epi;
  if (!p.epi.bias) p.epi.bias = L.bias;
  if (h->precision == DDH_PREC_FP32) {
    p.A = a32;
    p.W = L.wt32;
    p.ldw = L.N;
    launch_simt_gemm(p, L.N, st);
  } else {
    p.A = a16;
    p.dbg = (h->tl_gemm >= 0 && h->launches == h->tl_gemm) ? h->dbg : nullptr;
    launch_tc_gemm(p, L.map, L.N, st);
  }
  h->launches++;
  return DDH_OK;
}

size_t ws_bytes_for(const ddh_shape& s, int B, int precision) {
  const size_t M = (size_t)B * s.num_anchors, Dm = 256, F = s.d_ffn;
  const size_t rcap = std::min((size_t)s.num_anchors * s.num_poses * 4, (size_t)s.bev_h * s.bev_w);
  size_t b = 0;
  b += (size_t)B * s.bev_h * s.bev_w * s.bev_channels * (precision == DDH_PREC_BF16 ? 2 : 4);
  b += M * s.num_poses * 2 * 4 * 2;                      // img, pts
  b += M * 512 * 6 + M * Dm * 6 * 9 + M * F * 6;          // emb, activations, ffn hidden
  b += M * Dm * 4 * 2;                                   // qh, r2
  b += (size_t)s.num_layers * B * s.num_agents * 512 * 4 + (size_t)s.num_layers * B * Dm * 4;
  b += (size_t)B * (s.num_agents + 1) * Dm * 2;
  b += (size_t)B * rcap * 4 + B * 4 + M * s.num_poses * 4 * 8;
  b += (size_t)B * rcap * Dm * 4;                        // V
  b += M * s.num_poses * 3 * 4 + M * 4;
  if (precision == DDH_PREC_BF16 && s.num_anchors <= 128) {   // chain engine: bf16 K|V, tiled q0
    const size_t spt = 128 / s.num_anchors, tiles = (B + spt - 1) / spt;
    b += (size_t)s.num_layers * B * s.num_agents * CH_KV_LD * 2 + tiles * 128 * Dm * 4;
    if (s.num_steps >= 2 && B > RES_MAX_B) {   // kept value rows (conv_reuse; skipped when memory is short)
      const size_t HWs = (size_t)s.bev_h * s.bev_w, vcap = std::min(HWs, (size_t)s.num_steps * rcap);
      b += (size_t)s.num_layers * B * (vcap * Dm * 2 + HWs * 2 + 4) + (size_t)B * rcap * 8;
    }
  }
  return b;
}

int ensure_ws(ddh_handle* h, int B) {
  if (B <= h->cap_B) return DDH_OK;
  cudaDeviceSynchronize();
  free_all(h->owned_ws);
  h->cap_B = 0;
  // the engines test these pointers to decide what to write: none may survive a change of precision
  h->emb16 = h->e1_16 = h->q0_16 = h->agents16 = h->ego16 = h->s16 = h->x1_16 = h->o16 = h->x2_16 = h->h16 =
      h->x3_16 = h->c1_16 = h->r1_16 = nullptr;
  h->emb32 = h->e1_32 = h->s32 = h->o32 = h->x2_32 = h->h32 = h->x3_32 = h->c1_32 = h->r1_32 = h->V = nullptr;
  h->vkeep = nullptr;
  const ddh_shape& s = h->shp;
  const size_t M = (size_t)B * s.num_anchors, F = s.d_ffn;
  const int L = s.num_layers;
  h->rcap = (int)std::min((size_t)s.num_anchors * s.num_poses * 4, (size_t)s.bev_h * s.bev_w);
  const bool bf = h->precision == DDH_PREC_BF16;
  auto& o = h->owned_ws;
  int rc = 0;
#define WS(ptr, count) do { rc = dev_alloc(h, o, &(ptr), (size_t)(count)); if (rc) return rc; } while (0)
  {
    unsigned char* p = nullptr;
    WS(p, (size_t)B * s.bev_h * s.bev_w * s.bev_channels * (bf ? 2 : 4));
    h->bev_nhwc = p;
  }
  h->bev_nhwc_lo = nullptr;
  if (!bf && h->tf32_conv_ok) {
    unsigned char* p = nullptr;
    WS(p, (size_t)B * s.bev_h * s.bev_w * s.bev_channels * 4);
    h->bev_nhwc_lo = p;
  }
  WS(h->img, M * s.num_poses * 2);
  WS(h->pts, M * s.num_poses * 2);
  WS(h->q0_32, M * D);
  WS(h->kv32, (size_t)L * B * s.num_agents * 2 * D);
  WS(h->egov, (size_t)L * B * D);
  WS(h->upix, (size_t)B * h->rcap);
  WS(h->nuniq, B);
  WS(h->conv_rows, s.num_layers * s.num_steps);
  WS(h->conv_sched, 4);
  WS(h->need_seg, (size_t)B * 64);
  WS(h->done_seg, (size_t)B * 64);
  WS(h->dbg, 1024);
  WS(h->ent_slot, M * s.num_poses * 4);
  WS(h->ent_w, M * s.num_poses * 4);
  if (!bf) WS(h->V, (size_t)B * h->rcap * D);
  WS(h->x1_32, M * D);
  WS(h->qh32, M * D);
  WS(h->r2_32, M * D);
  WS(h->modes_buf, M * s.num_poses * 3);
  WS(h->scores_buf, M);
  if (bf) {
    WS(h->emb16, M * 512); WS(h->e1_16, M * D); WS(h->q0_16, M * D);
    WS(h->agents16, (size_t)B * s.num_agents * D); WS(h->ego16, (size_t)B * D);
    WS(h->s16, M * D); WS(h->x1_16, M * D); WS(h->o16, M * D); WS(h->x2_16, M * D);
    WS(h->h16, M * F); WS(h->x3_16, M * D); WS(h->c1_16, M * D); WS(h->r1_16, M * D);
    // fp32 taps kept for debugging only where cheap
    WS(h->s32, M * D); WS(h->x2_32, M * D); WS(h->x3_32, M * D);
    if (h->chain_ok) {
      const size_t tiles = ((size_t)B + h->chain_spt - 1) / h->chain_spt;
      WS(h->kv16, (size_t)L * B * s.num_agents * CH_KV_LD);
      WS(h->q0t, tiles * 128 * D);
      WS(h->logit_part, tiles * 2 * 128 * CH_LOGITS);
      rc = encode_wmap(h, &h->smap, h->s16, (int)M, D, 128);
      if (rc) return rc;
      // kept value rows: only when they take a modest share of the free memory (the engine works without)
      h->vkeep = nullptr;
      h->vcap = (int)std::min((size_t)s.bev_h * s.bev_w, (size_t)s.num_steps * h->rcap);
      const size_t keep_bytes = (size_t)L * B * h->vcap * D * 2, HWs = (size_t)s.bev_h * s.bev_w;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      if (h->conv_reuse && s.num_steps >= 2 && B > RES_MAX_B && h->vcap < 0x7fff && (HWs & 7) == 0 &&
          h->persistent_conv >= 2 && s.num_anchors <= 64 && s.num_poses == 8 &&   // (the shapes tc_conv3_kernel serves)
          (size_t)B * HWs < ((size_t)1 << 31) && (size_t)B * h->vcap < ((size_t)1 << 31) &&
          keep_bytes <= free_b / 3) {
        WS(h->vkeep, (size_t)L * B * h->vcap * D);
        WS(h->slot_tab, (size_t)L * B * HWs);
        WS(h->slot_cnt, (size_t)L * B);
        WS(h->new_list, (size_t)B * h->rcap);
        WS(h->new_count, (size_t)s.num_layers * s.num_steps);
      }
    }
  } else {
    WS(h->emb32, M * 512); WS(h->e1_32, M * D);
    WS(h->s32, M * D); WS(h->o32, M * D); WS(h->x2_32, M * D);
    WS(h->h32, M * F); WS(h->x3_32, M * D); WS(h->c1_32, M * D); WS(h->r1_32, M * D);
  }
#undef WS
  h->cap_B = B;
  return DDH_OK;
}

void register_taps(ddh_handle* h, int B) {
  const ddh_shape& s = h->shp;
  const size_t M = (size_t)B * s.num_anchors;
  auto& t = h->taps;
  t.clear();
  t["img"] = {h->img, M * s.num_poses * 2 * 4};
  t["pts"] = {h->pts, M * s.num_poses * 2 * 4};
  t["q0"] = {h->q0_32, M * D * 4};
  t["kv"] = {h->kv32, (size_t)s.num_layers * B * s.num_agents * 2 * D * 4};
  t["egov"] = {h->egov, (size_t)s.num_layers * B * D * 4};
  t["upix"] = {h->upix, (size_t)B * h->rcap * 4};
  t["nuniq"] = {h->nuniq, (size_t)B * 4};
  t["dbg"] = {h->dbg, (size_t)1024 * 8};
  t["done_seg"] = {h->done_seg, (size_t)B * std::max(h->seg_nw32, 1) * 4};   // (of the last forward)
  t["conv_rows"] = {h->conv_rows, (size_t)s.num_layers * s.num_steps * 4};
  t["ent_slot"] = {h->ent_slot, M * s.num_poses * 4 * 4};
  t["ent_w"] = {h->ent_w, M * s.num_poses * 4 * 4};
  t["V"] = {h->V, (size_t)B * h->rcap * D * 4};
  t["s"] = {h->s32, M * D * 4};
  t["x1"] = {h->x1_32, M * D * 4};
  t["qh"] = {h->qh32, M * D * 4};
  t["x2"] = {h->x2_32, M * D * 4};
  t["x3"] = {h->x3_32, M * D * 4};
  t["r2"] = {h->r2_32, M * D * 4};
  t["film"] = {h->film, (size_t)s.num_steps * s.num_layers * 2 * D * 4};
  t["modes"] = {h->modes_buf, M * s.num_poses * 3 * 4};
  t["scores"] = {h->scores_buf, M * 4};
  if (h->precision == DDH_PREC_BF16) t["bev_nhwc"] = {h->bev_nhwc, (size_t)B * s.bev_h * s.bev_w * s.bev_channels * 2};
  else t["bev_nhwc"] = {h->bev_nhwc, (size_t)B * s.bev_h * s.bev_w * s.bev_channels * 4};
}

// Anchor-resident engine (kernels_res2.cu): full-matrix tensor maps (box 64 k x 128 rows), the
// static stage schedule its TMA / MMA threads walk, constants and the L2 exchange buffers.
int build_res2(ddh_handle* h, cudaStream_t st) {
  h->res2_ok = false;
  free_all(h->owned_res2);
  h->res2_consts = nullptr;
  const ddh_shape& s = h->shp;
  const int A = s.num_anchors, P = s.num_poses, Na = s.num_agents, F = s.d_ffn, L = s.num_layers,
            S = s.num_steps, H = s.bev_h, W = s.bev_w;
  if (h->res_mode != 2 || h->precision != DDH_PREC_BF16) return DDH_OK;
  if (A > 28 || A * P > 256 || P != 8 || Na > 30 || F > 1024 || F % 256 || s.num_heads != 8 ||
      H * W > 4096 || H > 64 || W % 32 || L > 2 || S > RES_MAX_S || s.bev_channels != 256)
    return DDH_OK;
  if (const int why = res2_engine_init()) {
    (void)why;
    return DDH_OK;
  }
  auto& o = h->owned_res2;
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  TRY(dev_alloc(h, o, &h->res2_consts, 1));
  R2Consts& C = h->res2_host;
  memset(&C, 0, sizeof C);
  // weights re-packed as pre-swizzled shared-memory images (one bulk copy per 32 KiB fill)
  auto pack = [&](const __nv_bfloat16* w, int N, int K, int rows, const void** out) -> int {
    __nv_bfloat16* pk;
    int r = dev_alloc(h, o, &pk, (size_t)N * K);
    if (r) return r;
    launch_pack_sw128(w, pk, N, K, rows, st);
    *out = pk;
    return DDH_OK;
  };
  const void *w_enc0, *w_enc3;
  struct LayerW { const void *kvego, *bev_out, *q, *attn_out, *ffn0, *ffn2, *reg0, *reg2, *cls0, *cls3, *conv; };
  std::vector<LayerW> lw(L);
  TRY(pack(h->enc0.w16, D, 64 * P, 64, &w_enc0));
  TRY(pack(h->enc3.w16, D, D, 64, &w_enc3));
  for (int l = 0; l < L; ++l) {
    PackedLayer& pl = h->layers[l];
    __nv_bfloat16* kvego;
    float* b_kvego;
    TRY(dev_alloc(h, o, &kvego, (size_t)3 * D * D));
    TRY(dev_alloc(h, o, &b_kvego, (size_t)3 * D));
    CU_TRY(h, cudaMemcpyAsync(kvego, pl.kv.w16, (size_t)2 * D * D * 2, cudaMemcpyDeviceToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(kvego + (size_t)2 * D * D, pl.ego.w16, (size_t)D * D * 2, cudaMemcpyDeviceToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(b_kvego, pl.kv.bias, (size_t)2 * D * 4, cudaMemcpyDeviceToDevice, st));
    CU_TRY(h, cudaMemcpyAsync(b_kvego + 2 * D, pl.ego.bias, (size_t)D * 4, cudaMemcpyDeviceToDevice, st));
    TRY(pack(kvego, 3 * D, D, 3 * D / RES_CL, &lw[l].kvego));
    TRY(pack(pl.bev_out.w16, D, D, 64, &lw[l].bev_out));
    TRY(pack(pl.q.w16, D, D, 64, &lw[l].q));
    TRY(pack(pl.attn_out.w16, D, D, 64, &lw[l].attn_out));
    TRY(pack(pl.ffn0.w16, F, D, 64, &lw[l].ffn0));
    TRY(pack(pl.ffn2.w16, D, F, 64, &lw[l].ffn2));
    TRY(pack(pl.reg0.w16, D, D, 64, &lw[l].reg0));
    TRY(pack(pl.reg2.w16, D, D, 64, &lw[l].reg2));
    TRY(pack(pl.cls0.w16, D, D, 64, &lw[l].cls0));
    TRY(pack(pl.cls3.w16, D, D, 64, &lw[l].cls3));
    TRY(pack(pl.conv.w16, D, pl.conv.K, 64, &lw[l].conv));
    ResLayerC& lc = C.layer[l];
    lc.b_kvego = b_kvego; lc.b_bev_out = pl.bev_out.bias; lc.b_q = pl.q.bias;
    lc.b_attn_out = pl.attn_out.bias; lc.b_ffn0 = pl.ffn0.bias; lc.b_ffn2 = pl.ffn2.bias;
    lc.b_reg0 = pl.reg0.bias; lc.b_reg2 = pl.reg2.bias; lc.b_cls0 = pl.cls0.bias;
    lc.b_cls3 = pl.cls3.bias; lc.b_conv = pl.conv.bias;
    lc.attw_w = pl.attw_w; lc.attw_b = pl.attw_b;
    lc.norm1_g = pl.norm1_g; lc.norm1_b = pl.norm1_b; lc.norm2_g = pl.norm2_g; lc.norm2_b = pl.norm2_b;
    lc.norm3_g = pl.norm3_g; lc.norm3_b = pl.norm3_b;
    lc.cls_ln2_g = pl.cls_ln2_g; lc.cls_ln2_b = pl.cls_ln2_b;
    lc.cls_ln5_g = pl.cls_ln5_g; lc.cls_ln5_b = pl.cls_ln5_b;
    lc.cls6_w = pl.cls6_w; lc.cls6_b = pl.cls6_b; lc.reg4_w = pl.reg4_w; lc.reg4_b = pl.reg4_b;
    lc.conv_map = nullptr;
  }
  C.b_enc0 = h->enc0.bias; C.b_enc3 = h->enc3.bias; C.enc_ln_g = h->enc_ln_g; C.enc_ln_b = h->enc_ln_b;
  C.anchors = h->anchors; C.dim_t = h->dim_t; C.film = h->film;
  C.A = A; C.P = P; C.Na = Na; C.F = F; C.L = L; C.S = S; C.H = H; C.W = W; C.heads = s.num_heads;
  C.rcap = (int)std::min((size_t)A * P * 4, (size_t)H * W);
  C.oc = OdoConsts{s.lidar_max_x, s.lidar_max_y};
  const float ac_tr = h->ac[s.trunc_timestep];
  C.sa_tr = sqrtf(ac_tr); C.sb_tr = sqrtf(1.0f - ac_tr);
  for (int si = 0; si < S; ++si) {
    const int t = h->roll[si], prev = t - 1;
    const float ac_t = h->ac[t], ac_p = prev >= 0 ? h->ac[prev] : 1.0f;
    C.dc[si] = DdimCoef{sqrtf(ac_t), sqrtf(1.0f - ac_t), sqrtf(ac_p), sqrtf(1.0f - ac_p)};
  }
  // the schedule, in the program order of the compute warps (kernels_res2.cu); chain stage k
  // reads B operand buffer k & 1
  constexpr unsigned short ACC_LIN = 256;   // linear tile t, issuer j at ACC_LIN + 32 t + 8 j; cls branch at ACC_LIN + 128
  int n = 0, k = 0;
  auto stage = [&](const void* m, int rows, int mtiles, int K, int acc_col, int flags, int bsel) {
    R2Stage& g = C.stages[n++];
    g.w = m; g.rows = (unsigned short)rows; g.mtiles = (unsigned char)mtiles;
    g.kchunks = (unsigned char)(K / 64); g.acc_col = (unsigned short)acc_col;
    g.flags = (unsigned char)flags; g.bsel = (unsigned char)bsel;
  };
  const int both = R2F_WAITB | R2F_COMMIT;
  for (int l = 0; l < L; ++l)
    stage(lw[l].kvego, 3 * D / RES_CL, 1, D, ACC_LIN + 128 * l,
          R2F_N32 | R2F_RANK16 | (l == 0 ? R2F_WAITB : 0) | (l == L - 1 ? R2F_COMMIT : 0), 0);
  const int fq = F / 4;   // hidden features per CTA
  for (int si = 0; si < S; ++si) {
    stage(w_enc0, 64, 1, 64 * P, ACC_LIN, both, k++ & 1);
    stage(w_enc3, 64, 1, D, ACC_LIN, both, k++ & 1);
    for (int l = 0; l < L; ++l) {
      const bool want_cls = (si == S - 1) && (l == L - 1);
      stage(lw[l].conv, 0, 0, 0, 0, R2F_CONV, 0);
      stage(lw[l].bev_out, 64, 1, D, ACC_LIN, both, k++ & 1);
      stage(lw[l].q, 64, 1, D, ACC_LIN, both, k++ & 1);
      stage(lw[l].attn_out, 64, 1, D, ACC_LIN, both, k++ & 1);
      stage(lw[l].ffn0, 64, fq / 64, D, ACC_LIN, both, k++ & 1);
      stage(lw[l].ffn2, 64, 1, F, ACC_LIN, both, k++ & 1);
      if (!want_cls) {
        stage(lw[l].reg0, 64, 1, D, ACC_LIN, both, k++ & 1);
        stage(lw[l].reg2, 64, 1, D, ACC_LIN, both, k++ & 1);
      } else {
        stage(lw[l].reg0, 64, 1, D, ACC_LIN, R2F_WAITB, k & 1);
        stage(lw[l].cls0, 64, 1, D, ACC_LIN + 128, R2F_COMMIT, k++ & 1);
        stage(lw[l].reg2, 64, 1, D, ACC_LIN, R2F_WAITB, k++ & 1);
        stage(lw[l].cls3, 64, 1, D, ACC_LIN + 128, R2F_COMMIT, 2);
      }
    }
  }
  C.n_stages = n;
  TRY(dev_alloc(h, o, &C.kv, (size_t)RES_MAX_B * L * Na * 2 * D));
  TRY(dev_alloc(h, o, &C.egov, (size_t)RES_MAX_B * L * D));
  TRY(dev_alloc(h, o, &C.bev_nhwc, (size_t)RES_MAX_B * H * W * D));
  TRY(dev_alloc(h, o, &C.tap_q0, (size_t)RES_MAX_B * A * D));
  TRY(dev_alloc(h, o, &C.tap_x1, (size_t)RES_MAX_B * A * D));
  TRY(dev_alloc(h, o, &C.tap_regraw, (size_t)RES_MAX_B * A * 3 * P));
  // dense mode: whole-map value_proj on helper clusters of the same launch
  h->dense_ok = false;
  h->dense_amap_base = nullptr;
  if (W == 64 && H % 2 == 0 && res2_max_clusters() >= 2 && h->encode) {
    DenseArgs& da = h->dense_host;
    memset(&da, 0, sizeof da);
    TRY(dev_alloc(h, o, &da.V, (size_t)DENSE_MAX_B * L * H * W * D));
    TRY(dev_alloc(h, o, &da.ctrl, (size_t)DC_WORDS));
    CU_TRY(h, cudaMemsetAsync(da.ctrl, 0, (size_t)DC_WORDS * 4, st));
    for (int l = 0; l < L; ++l) {   // weight tiles of a job: [128 channels x 64 k]
      TRY(encode_wmap(h, &da.wmap[l], h->layers[l].conv.w16, 256, h->layers[l].conv.K, 128));
      da.bias[l] = h->layers[l].conv.bias;
    }
    da.L = L; da.H = H; da.W = W;
    da.enabled = 1;
    h->dense_ok = true;
  }
#undef TRY
  CU_TRY(h, cudaMemcpyAsync(h->res2_consts, &C, sizeof(R2Consts), cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  h->res2_ok = true;
  return DDH_OK;
}

}  // namespace

// =====================================================================================
extern "C" {

int ddh_abi_version(void) { return DDH_ABI_VERSION; }

const char* ddh_build_info(void) {
  return "ddh sm_100a: engines=simt_f32,tcgen05_bf16(chain,conv3,resident); tma=weights,features; "
#ifdef DDH_CHECKED
         "checked build (bounded mbarrier waits, index asserts); "
#endif
         "cuda nvcc";
}

const char* ddh_last_error(const ddh_handle* h) {
  return h ? h->err.c_str() : g_create_error.c_str();
}

int ddh_create(const ddh_shape* s, ddh_handle** out) {
  if (!s || !out) return fail(nullptr, DDH_ERR_BAD_ARG, "ddh_create: null argument");
  *out = nullptr;
  char msg[256];
#define REQUIRE(cond, text)                                              \
  do {                                                                   \
    if (!(cond)) {                                                       \
      snprintf(msg, sizeof msg, "ddh_create: unsupported shape: %s", text); \
      return fail(nullptr, DDH_ERR_UNSUPPORTED, msg);                    \
    }                                                                    \
  } while (0)
  REQUIRE(s->d_model == 256, "d_model must be 256");
  REQUIRE(s->bev_channels == 256, "bev_channels must be 256");
  REQUIRE(s->num_poses == 8, "num_poses must be 8");
  REQUIRE(s->num_heads == 8, "num_heads must be 8 (head_dim 32)");
  REQUIRE(s->d_ffn > 0 && s->d_ffn % 256 == 0, "d_ffn must be a positive multiple of 256");
  REQUIRE(s->num_agents >= 1 && s->num_agents <= 32, "num_agents must be in [1, 32]");
  REQUIRE(s->num_anchors >= 1 && s->num_anchors <= 4096, "num_anchors must be in [1, 4096]");
  REQUIRE(s->bev_h >= 1 && s->bev_w >= 1 && (s->bev_h * s->bev_w) % 64 == 0 &&
              s->bev_h * s->bev_w <= 65535 && s->bev_w < 32768 && s->bev_h < 32768,
          "bev_h*bev_w must be a multiple of 64 and <= 65535");
  REQUIRE(s->num_layers >= 1 && s->num_layers <= 16, "num_layers must be in [1, 16]");
  REQUIRE(s->num_steps >= 1 && s->num_steps <= 20, "num_steps must be in [1, 20]");
  REQUIRE(s->trunc_timestep >= 0 && s->trunc_timestep < 1000, "trunc_timestep must be in [0, 1000)");
  REQUIRE(s->lidar_max_x > 0.f && s->lidar_max_y > 0.f, "lidar_max_x/y must be positive");
#undef REQUIRE
  ddh_handle* h = new ddh_handle();
  h->shp = *s;
  default_alphas_cumprod(h->ac);
  make_roll(s->num_steps, h->roll);
  *out = h;
  return DDH_OK;
}

void ddh_destroy(ddh_handle* h) {
  if (!h) return;
  free_all(h->owned_w);
  free_all(h->owned_ws);
  free_all(h->owned_host);
  free_all(h->owned_res2);
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : h->sync_events) cudaEventDestroy(e);
  if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
  delete h;
}

int ddh_set_alphas_cumprod(ddh_handle* h, const float* table, int n) {
  if (!h || !table || n < 21) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_alphas_cumprod: bad argument");
  if (n > 1000) n = 1000;
  for (int i = 0; i < n; ++i) h->ac[i] = table[i];
  return DDH_OK;
}

int ddh_get_alphas_cumprod(const ddh_handle* h, float* table, int n) {
  if (!h || !table || n < 1 || n > 1000) return DDH_ERR_BAD_ARG;
  for (int i = 0; i < n; ++i) table[i] = h->ac[i];
  return DDH_OK;
}

size_t ddh_workspace_bytes(const ddh_handle* h, int B) {
  if (!h || B <= 0) return 0;
  return ws_bytes_for(h->shp, B, h->precision < 0 ? DDH_PREC_FP32 : h->precision);
}

int ddh_reserve(ddh_handle* h, int B) {
  if (!h || B <= 0) return fail(h, DDH_ERR_BAD_ARG, "ddh_reserve: bad argument");
  if (!h->packed) return fail(h, DDH_ERR_NOT_PACKED, "ddh_reserve: call ddh_pack_weights first");
  return ensure_ws(h, B);
}

int ddh_pack_weights(ddh_handle* h, const ddh_weight_ptrs* w, int precision, void* stream) {
  if (!h || !w || !w->layers) return fail(h, DDH_ERR_BAD_ARG, "ddh_pack_weights: null argument");
  if (precision != DDH_PREC_FP32 && precision != DDH_PREC_BF16)
    return fail(h, DDH_ERR_BAD_ARG, "ddh_pack_weights: unknown precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DDH_ERR_CUDA, "ddh_pack_weights: no CUDA device (there is no CPU fallback)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const ddh_shape& s = h->shp;
  // a change of precision changes the workspace layout
  cudaDeviceSynchronize();
  free_all(h->owned_w);
  free_all(h->owned_ws);   // the workspace layout depends on precision and engine selection
  h->cap_B = 0;
  h->packed = false;
  h->precision = precision;
  if (precision == DDH_PREC_BF16 && tc_conv_smem_bytes(s.num_anchors, s.num_poses * 4) > 227 * 1024)
    return fail(h, DDH_ERR_UNSUPPORTED, "ddh_pack_weights: too many anchors for the bf16 conv kernel's shared memory");
  if (precision == DDH_PREC_BF16 && !h->tc_ready) {
    int e = tc_engine_init();
    if (e) return fail(h, DDH_ERR_CUDA, std::string("tc_engine_init: ") + cudaGetErrorString((cudaError_t)e));
    h->tc_ready = true;
  }
  // fp32 engine: value_proj as 3xTF32 on the tensor core (tc_conv2_kernel<true>) when its shared
  // memory fits; everything else of the fp32 engine stays on the CUDA cores
  h->tf32_conv_ok = false;
  if (precision == DDH_PREC_FP32 && h->fp32_tensor_conv && s.bev_channels == 256 &&
      tc_conv_smem_bytes(s.num_anchors, s.num_poses * 4) <= 227 * 1024) {
    if (!h->tc_ready) {
      int e = tc_engine_init();
      if (e) return fail(h, DDH_ERR_CUDA, std::string("tc_engine_init: ") + cudaGetErrorString((cudaError_t)e));
      h->tc_ready = true;
    }
    h->tf32_conv_ok = true;
  }
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  const int A = s.num_anchors, P = s.num_poses, F = s.d_ffn, L = s.num_layers, S = s.num_steps;
  TRY(copy_vec(h, &h->anchors, w->plan_anchor, (size_t)A * P * 2, st));
  TRY(pack_linear(h, h->enc0, w->enc0_w, w->enc0_b, D, 64 * P, st));
  TRY(copy_vec(h, &h->enc_ln_g, w->enc_ln_w, D, st));
  TRY(copy_vec(h, &h->enc_ln_b, w->enc_ln_b, D, st));
  TRY(pack_linear(h, h->enc3, w->enc3_w, w->enc3_b, D, D, st));
  {  // dim_t of gen_sineembed_for_position (modules/blocks.py:27-30), fp32 like torch
    float dt[32];
    for (int i = 0; i < 32; ++i) dt[i] = powf(10000.0f, (float)(2 * (i / 2)) / 32.0f);
    TRY(dev_alloc(h, h->owned_w, &h->dim_t, 32));
    CU_TRY(h, cudaMemcpyAsync(h->dim_t, dt, sizeof dt, cudaMemcpyHostToDevice, st));
    CU_TRY(h, cudaStreamSynchronize(st));   // dt is a stack buffer
  }
  // ---- time path: temb_s = time_mlp(SinusoidalPosEmb(t_s)) (:463-468), identical for all scenes
  float *semb, *hid, *temb;
  TRY(dev_alloc(h, h->owned_w, &semb, D));
  TRY(dev_alloc(h, h->owned_w, &hid, 4 * D));
  TRY(dev_alloc(h, h->owned_w, &temb, (size_t)S * D));
  TRY(dev_alloc(h, h->owned_w, &h->film, (size_t)S * L * 2 * D));
  for (int si = 0; si < S; ++si) {
    launch_time_sinemb(semb, D, h->roll[si], st);
    launch_matvec(w->time1_w, semb, w->time1_b, hid, 4 * D, D, 0, st);
    launch_matvec(w->time3_w, hid, w->time3_b, temb + (size_t)si * D, D, 4 * D, 1, st);
  }
  h->layers.assign(L, PackedLayer());
  for (int l = 0; l < L; ++l) {
    const ddh_layer_weights& lw = w->layers[l];
    PackedLayer& pl = h->layers[l];
    // conv: K = 9*C ordered (tap, channel)
    pl.conv.N = 256; pl.conv.K = 9 * s.bev_channels;
    TRY(copy_vec(h, &pl.conv.bias, lw.bev_conv_b, 256, st));
    if (precision == DDH_PREC_FP32) {
      TRY(dev_alloc(h, h->owned_w, &pl.conv.wt32, (size_t)256 * pl.conv.K));
      launch_pack_conv_f32(lw.bev_conv_w, pl.conv.wt32, 256, s.bev_channels, st);
      if (h->tf32_conv_ok) {
        TRY(dev_alloc(h, h->owned_w, &pl.conv.w2_32, (size_t)256 * 2 * pl.conv.K));
        launch_pack_conv_tf32x2(lw.bev_conv_w, pl.conv.w2_32, 256, s.bev_channels, st);
        TRY(encode_wmap_f32(h, &pl.conv.map_tf32, pl.conv.w2_32, 256, 2 * pl.conv.K));
      }
    } else {
      TRY(dev_alloc(h, h->owned_w, &pl.conv.w16, (size_t)256 * pl.conv.K));
      launch_pack_conv_bf16(lw.bev_conv_w, pl.conv.w16, 256, s.bev_channels, st);
      TRY(make_wmap(h, pl.conv));
      TRY(make_wmap(h, pl.conv, 64));
    }
    TRY(copy_vec(h, &pl.attw_w, lw.bev_attw_w, (size_t)P * D, st));
    TRY(copy_vec(h, &pl.attw_b, lw.bev_attw_b, P, st));
    if (l == 0) TRY(dev_alloc(h, h->owned_w, &h->attw_all, (size_t)L * P * D));
    CU_TRY(h, cudaMemcpyAsync(h->attw_all + (size_t)l * P * D, lw.bev_attw_w, (size_t)P * D * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
    TRY(pack_linear(h, pl.bev_out, lw.bev_out_w, lw.bev_out_b, D, D, st));
    TRY(pack_linear(h, pl.q, lw.agent_in_w, lw.agent_in_b, D, D, st));
    TRY(pack_linear(h, pl.kv, lw.agent_in_w + (size_t)D * D, lw.agent_in_b + D, 2 * D, D, st));
    TRY(pack_linear(h, pl.attn_out, lw.agent_out_w, lw.agent_out_b, D, D, st));
    {  // ego collapse: W_ego = Wo . Wv, b_ego = Wo . bv + bo   (single key => softmax == 1)
      float *wego, *bego;
      TRY(dev_alloc(h, h->owned_w, &wego, (size_t)D * D));
      TRY(dev_alloc(h, h->owned_w, &bego, D));
      GemmParams gp;
      gp.A = lw.ego_out_w; gp.lda = D; gp.M = D; gp.K = D;
      gp.W = lw.ego_in_w + (size_t)2 * D * D; gp.ldw = D;   // Wv as Wt[k][n]
      gp.epi.out_f32 = wego; gp.epi.ldo32 = D;
      launch_simt_gemm(gp, D, st);
      launch_matvec(lw.ego_out_w, lw.ego_in_b + 2 * D, lw.ego_out_b, bego, D, D, 0, st);
      TRY(pack_linear(h, pl.ego, wego, bego, D, D, st));
    }
    TRY(pack_linear(h, pl.ffn0, lw.ffn0_w, lw.ffn0_b, F, D, st));
    TRY(pack_linear(h, pl.ffn2, lw.ffn2_w, lw.ffn2_b, D, F, st));
    TRY(copy_vec(h, &pl.norm1_g, lw.norm1_w, D, st)); TRY(copy_vec(h, &pl.norm1_b, lw.norm1_b, D, st));
    TRY(copy_vec(h, &pl.norm2_g, lw.norm2_w, D, st)); TRY(copy_vec(h, &pl.norm2_b, lw.norm2_b, D, st));
    TRY(copy_vec(h, &pl.norm3_g, lw.norm3_w, D, st)); TRY(copy_vec(h, &pl.norm3_b, lw.norm3_b, D, st));
    // FiLM vectors per (step, layer): Linear(D->2D)(Mish(temb_s))   (:276-294)
    for (int si = 0; si < S; ++si)
      launch_matvec(lw.film_w, temb + (size_t)si * D, lw.film_b,
                    h->film + ((size_t)si * L + l) * 2 * D, 2 * D, D, 1, st);
    TRY(pack_linear(h, pl.cls0, lw.cls0_w, lw.cls0_b, D, D, st));
    TRY(copy_vec(h, &pl.cls_ln2_g, lw.cls_ln2_w, D, st)); TRY(copy_vec(h, &pl.cls_ln2_b, lw.cls_ln2_b, D, st));
    TRY(pack_linear(h, pl.cls3, lw.cls3_w, lw.cls3_b, D, D, st));
    TRY(copy_vec(h, &pl.cls_ln5_g, lw.cls_ln5_w, D, st)); TRY(copy_vec(h, &pl.cls_ln5_b, lw.cls_ln5_b, D, st));
    TRY(copy_vec(h, &pl.cls6_w, lw.cls6_w, D, st)); TRY(copy_vec(h, &pl.cls6_b, lw.cls6_b, 1, st));
    TRY(pack_linear(h, pl.reg0, lw.reg0_w, lw.reg0_b, D, D, st));
    TRY(pack_linear(h, pl.reg2, lw.reg2_w, lw.reg2_b, D, D, st));
    TRY(copy_vec(h, &pl.reg4_w, lw.reg4_w, (size_t)3 * P * D, st));
    TRY(copy_vec(h, &pl.reg4_b, lw.reg4_b, 3 * P, st));
    if (precision == DDH_PREC_BF16) {
      TRY(dev_alloc(h, h->owned_w, &pl.reg4_hl, (size_t)64 * D));
      launch_pack_hilo(lw.reg4_w, pl.reg4_hl, 3 * P, D, st);
      TRY(encode_wmap(h, &pl.reg4_map, pl.reg4_hl, 64, D, 64));
    }
  }
  // scene-tile chain engine: shape limits of kernels_chain.cu
  h->chain_ok = false;
  h->chain_prog.clear();
  if (precision == DDH_PREC_BF16 && h->chain_enabled && A <= 128 && F <= 1024 && P == 8 &&
      s.num_agents <= 32 && chain_engine_init() == 0) {
    h->chain_ok = true;
    h->chain_spt = 128 / A;
  }
#undef TRY
  CU_TRY(h, cudaGetLastError());
  rc = build_res2(h, st);
  if (rc) return rc;
  h->packed = true;
  return DDH_OK;
}

}  // extern "C" (reopened below)

namespace {

// Per-chunk view of the workspace and of the caller's buffers: every pointer advanced to
// scene `s0`.  Layer-major buffers (kv32, egov) keep the full-batch layer stride `Btot`.
struct View {
  float *img, *pts, *emb32, *e1_32, *q0_32, *kv32, *egov, *ent_w, *V, *s32, *x1_32, *qh32, *o32,
      *x2_32, *h32, *x3_32, *c1_32, *r1_32, *r2_32, *modes_buf, *scores_buf;
  __nv_bfloat16 *emb16, *e1_16, *q0_16, *agents16, *ego16, *s16, *x1_16, *o16, *x2_16, *h16,
      *x3_16, *c1_16, *r1_16;
  int *upix, *nuniq, *ent_slot;
  void* bev_nhwc;
};
template <typename T>
T* adv(T* p, size_t n) { return p ? p + n : nullptr; }

View make_view(const ddh_handle* h, int s0) {
  const ddh_shape& s = h->shp;
  const size_t A = s.num_anchors, P = s.num_poses, Na = s.num_agents, F = s.d_ffn, z = s0;
  const size_t rows = z * A;
  View v;
  v.img = adv(h->img, rows * P * 2); v.pts = adv(h->pts, rows * P * 2);
  v.emb32 = adv(h->emb32, rows * 64 * P); v.emb16 = adv(h->emb16, rows * 64 * P);
  v.e1_32 = adv(h->e1_32, rows * D); v.e1_16 = adv(h->e1_16, rows * D);
  v.q0_32 = adv(h->q0_32, rows * D); v.q0_16 = adv(h->q0_16, rows * D);
  v.kv32 = adv(h->kv32, z * Na * 2 * D); v.egov = adv(h->egov, z * D);
  v.agents16 = adv(h->agents16, z * Na * D); v.ego16 = adv(h->ego16, z * D);
  v.upix = adv(h->upix, z * h->rcap); v.nuniq = adv(h->nuniq, z);
  v.ent_slot = adv(h->ent_slot, rows * P * 4); v.ent_w = adv(h->ent_w, rows * P * 4);
  v.V = adv(h->V, z * h->rcap * D);
  v.s32 = adv(h->s32, rows * D); v.s16 = adv(h->s16, rows * D);
  v.x1_32 = adv(h->x1_32, rows * D); v.x1_16 = adv(h->x1_16, rows * D);
  v.qh32 = adv(h->qh32, rows * D);
  v.o32 = adv(h->o32, rows * D); v.o16 = adv(h->o16, rows * D);
  v.x2_32 = adv(h->x2_32, rows * D); v.x2_16 = adv(h->x2_16, rows * D);
  v.h32 = adv(h->h32, rows * F); v.h16 = adv(h->h16, rows * F);
  v.x3_32 = adv(h->x3_32, rows * D); v.x3_16 = adv(h->x3_16, rows * D);
  v.c1_32 = adv(h->c1_32, rows * D); v.c1_16 = adv(h->c1_16, rows * D);
  v.r1_32 = adv(h->r1_32, rows * D); v.r1_16 = adv(h->r1_16, rows * D);
  v.r2_32 = adv(h->r2_32, rows * D);
  v.modes_buf = adv(h->modes_buf, rows * P * 3); v.scores_buf = adv(h->scores_buf, rows);
  const size_t bev_elt = h->precision == DDH_PREC_BF16 ? 2 : 4;
  v.bev_nhwc = h->bev_nhwc ? (unsigned char*)h->bev_nhwc + z * s.bev_h * s.bev_w * s.bev_channels * bev_elt : nullptr;
  return v;
}


// On-demand layout granularity of this call: segments of seg_px pixels (32 when the map is read in
// place from pinned host memory), mask words per scene; 0 words = map too large for the masks.
void pick_segments(ddh_handle* h) {
  const ddh_shape& s = h->shp;
  h->seg_px_call = h->host_map_call ? h->host_seg_px : h->seg_px;
  // (a host map is read in whole 64-pixel runs when the width allows: 256-byte PCIe reads)
  while (h->seg_px_call > h->seg_px && (s.bev_w % h->seg_px_call || s.bev_h * (s.bev_w / h->seg_px_call) > 2048))
    h->seg_px_call >>= 1;
  if (s.bev_w % h->seg_px_call) h->seg_px_call = h->seg_px;
  const int bits = s.bev_h * (s.bev_w / h->seg_px_call);
  h->seg_nw32 = (s.bev_w % h->seg_px_call == 0 && bits <= 2048) ? (bits + 31) / 32 : 0;
}

// ---- scene-tile chain engine (kernels_chain.cu): the program tables
struct ProgBuilder {
  ChainArgs& a;
  int nops = 0, nsteps = 0, npar = 0, parpos = 0, nmaps = 0;
  explicit ProgBuilder(ChainArgs& a_) : a(a_) {}
  int map(const CUtensorMap& m) { a.maps[nmaps] = m; return nmaps++; }
  int par(const float* src, int n) {
    const int off = parpos;
    a.par[npar++] = ChainParSrc{src, n, off};
    parpos += (n + 3) / 4 * 4;
    return off;
  }
  int op(int map, int a_chunk, int nk, int k0, int n0, int acc_col, int flags) {
    ChainOp& o = a.ops[nops];
    o.map = (uint8_t)map; o.a_chunk = (uint8_t)a_chunk; o.nk = (uint8_t)nk; o.k0 = (uint8_t)k0;
    o.n0 = (uint16_t)n0; o.acc_col = (uint16_t)acc_col; o.flags = (uint8_t)flags;
    return nops++;
  }
  void step(int op0, int nops_, int epi, int flags, int dst_chunk, int acc_col,
            std::initializer_list<int> pars) {
    ChainStep& st = a.steps[nsteps++];
    st.op0 = (uint8_t)op0; st.nops = (uint8_t)nops_; st.epi = (uint8_t)epi; st.flags = (uint8_t)flags;
    st.dst_chunk = (uint8_t)dst_chunk; st.acc_col = (uint16_t)acc_col;
    int i = 0;
    for (int v : pars) st.par[i++] = (uint16_t)v;
  }
  bool finish() {
    a.n_steps = nsteps; a.n_par = npar;
    return nops <= CH_MAX_OPS && nsteps <= CH_MAX_STEPS && npar <= CH_MAX_PAR && nmaps <= CH_MAX_MAPS &&
           parpos <= CH_PAR_FLOATS;
  }
};

// embedding + plan_anchor_encoder (:459-462, 601-609) of denoise step si
bool build_enc_program(ddh_handle* h, int si, ChainArgs& a) {
  memset(&a, 0, sizeof a);
  ProgBuilder b(a);
  const int m0 = b.map(h->enc0.map), m3 = b.map(h->enc3.map);
  const int p_b0 = b.par(h->enc0.bias, D), p_g = b.par(h->enc_ln_g, D), p_b = b.par(h->enc_ln_b, D);
  const int p_b3 = b.par(h->enc3.bias, D), p_dt = b.par(h->dim_t, 32);
  // attention-weight logits of every layer ride on the q0 epilogue when they fit (<= CH_LOGITS per row)
  const int nlog = h->shp.num_layers * h->shp.num_poses;
  h->logits_hoisted = h->shp.num_poses == 8 && nlog <= CH_LOGITS && h->attw_all != nullptr;
  const int p_lw = h->logits_hoisted ? b.par(h->attw_all, nlog * D) : 0;
  int o = b.op(m0, 0, h->enc0.K / 64, 0, 0, 0, 0);
  b.step(o, 1, CE_RELU_LN, 0, 0, 0, {p_b0, p_g, p_b, 0, 0, p_dt});
  o = b.op(m3, 0, 4, 0, 0, 256, 0);
  b.step(o, 1, CE_Q0, 0, 0, 256, {p_b3, p_lw, h->logits_hoisted ? nlog : 0});
  a.mode = 1;
  a.first_step = si == 0 ? 1 : 0;
  const float ac_tr = h->ac[h->shp.trunc_timestep];
  a.sa = sqrtf(ac_tr); a.sb = sqrtf(1.0f - ac_tr);
  return b.finish();
}

// one decoder-layer call after the BEV sampling (:355-380), layer l of denoise step si
bool build_layer_program(ddh_handle* h, int si, int l, ChainArgs& a) {
  memset(&a, 0, sizeof a);
  const ddh_shape& s = h->shp;
  const PackedLayer& pl = h->layers[l];
  const int L = s.num_layers, S = s.num_steps, nb = s.d_ffn / D;
  const bool last = (si == S - 1) && (l == L - 1);
  ProgBuilder b(a);
  const int m_bo = b.map(pl.bev_out.map), m_q = b.map(pl.q.map), m_ao = b.map(pl.attn_out.map);
  const int m_f0 = b.map(pl.ffn0.map), m_f2 = b.map(pl.ffn2.map);
  const int m_r0 = b.map(pl.reg0.map), m_r2 = b.map(pl.reg2.map), m_r4 = b.map(pl.reg4_map);
  const int p_bbo = b.par(pl.bev_out.bias, D), p_bao = b.par(pl.attn_out.bias, D), p_bq = b.par(pl.q.bias, D);
  const int p_n1g = b.par(pl.norm1_g, D), p_n1b = b.par(pl.norm1_b, D);
  const int p_n2g = b.par(pl.norm2_g, D), p_n2b = b.par(pl.norm2_b, D);
  const int p_bf0 = b.par(pl.ffn0.bias, s.d_ffn), p_bf2 = b.par(pl.ffn2.bias, D);
  const int p_n3g = b.par(pl.norm3_g, D), p_n3b = b.par(pl.norm3_b, D);
  const int p_film = b.par(h->film + ((size_t)si * L + l) * 2 * D, 2 * D);
  const int p_br0 = b.par(pl.reg0.bias, D), p_br2 = b.par(pl.reg2.bias, D);
  const int p_br4 = b.par(pl.reg4_b, 3 * s.num_poses);
  constexpr int R0 = 0, R1 = 4, T0 = 0, T1 = 256;
  int o;
  // cross_bev_attention output_proj + residual (modules/blocks.py:127-129)
  o = b.op(m_bo, R0, 4, 0, 0, T0, 0);
  b.step(o, 1, CE_X1, CS_WAIT_S, R0, T0, {p_bbo, p_bao});
  // cross_agent_attention: q projection, attention core (:355-357)
  o = b.op(m_q, R0, 4, 0, 0, T1, 0);
  b.step(o, 1, CE_ATTN, CS_KVGO, R1, T1, {p_bq});
  // out_proj accumulated onto x1 (+ bias) already in TMEM; norm1, + ego, norm2 (:358-365)
  o = b.op(m_ao, R1, 4, 0, 0, T0, CO_ACCUM);
  b.step(o, 1, CE_LN2EGO, 0, R0, T0, {p_n1g, p_n1b, p_n2g, p_n2b});
  // ffn (:368): hidden blocks of 256 streamed through region 1 as k-blocks of ffn.2
  o = b.op(m_f0, R0, 4, 0, 0, T1, 0);
  b.step(o, 1, CE_RELU, 0, R1, T1, {p_bf0});
  for (int j = 1; j < nb; ++j) {
    o = b.op(m_f2, R1, 4, 4 * (j - 1), 0, T0, j > 1 ? CO_ACCUM : 0);
    b.op(m_f0, R0, 4, 0, D * j, T1, 0);
    b.step(o, 2, CE_RELU, 0, R1, T1, {p_bf0 + D * j});
  }
  o = b.op(m_f2, R1, 4, 4 * (nb - 1), 0, T0, nb > 1 ? CO_ACCUM : 0);
  b.step(o, 1, CE_LN_FILM, 0, R0, T0, {p_bf2, p_n3g, p_n3b, p_film});   // norm3 + time FiLM (:368-373)
  // task_decoder (:244-256, 376-380)
  o = b.op(m_r0, R0, 4, 0, 0, T1, 0);
  b.step(o, 1, CE_RELU, last ? 0 : CS_SAFREE, R1, T1, {p_br0});
  o = b.op(m_r2, R1, 4, 0, 0, T1, 0);
  b.step(o, 1, CE_RELU, 0, R1, T1, {p_br2});
  if (!last) {
    o = b.op(m_r4, R1, 4, 0, 0, T1, CO_N64);
    b.step(o, 1, CE_TAIL, 0, R1, T1, {p_br4});
  } else {   // cls branch only where it is read (:631)
    const int m_c0 = b.map(pl.cls0.map), m_c3 = b.map(pl.cls3.map);
    const int p_bc0 = b.par(pl.cls0.bias, D), p_l2g = b.par(pl.cls_ln2_g, D), p_l2b = b.par(pl.cls_ln2_b, D);
    const int p_bc3 = b.par(pl.cls3.bias, D), p_l5g = b.par(pl.cls_ln5_g, D), p_l5b = b.par(pl.cls_ln5_b, D);
    const int p_w6 = b.par(pl.cls6_w, D), p_b6 = b.par(pl.cls6_b, 1);
    o = b.op(m_r4, R1, 4, 0, 0, T1, CO_N64);
    b.op(m_c0, R0, 4, 0, 0, T0, 0);
    b.step(o, 2, CE_TAIL, CS_SAFREE, R1, T1, {p_br4});
    b.step(0, 0, CE_RELU_LN, 0, R1, T0, {p_bc0, p_l2g, p_l2b});
    o = b.op(m_c3, R1, 4, 0, 0, T0, 0);
    b.step(o, 1, CE_SCORE, 0, R1, T0, {p_bc3, p_l5g, p_l5b, p_w6, p_b6});
  }
  a.mode = 0;
  a.do_ddim = (l == L - 1 && si != S - 1) ? 1 : 0;
  a.dc = DdimCoef{0.f, 1.f, 1.f, 0.f};
  if (a.do_ddim) {
    const int t = h->roll[si], prev = t - 1;   // set_timesteps(1000) => step ratio 1 (:584)
    const float ac_t = h->ac[t], ac_p = prev >= 0 ? h->ac[prev] : 1.0f;
    a.dc = DdimCoef{sqrtf(ac_t), sqrtf(1.0f - ac_t), sqrtf(ac_p), sqrtf(1.0f - ac_p)};
  }
  return b.finish();
}

int build_chain_programs(ddh_handle* h) {
  const ddh_shape& s = h->shp;
  const int L = s.num_layers, S = s.num_steps;
  h->chain_prog.assign((size_t)S + (size_t)S * L, ChainArgs());
  for (int si = 0; si < S; ++si) {
    if (!build_enc_program(h, si, h->chain_prog[si]))
      return fail(h, DDH_ERR_UNSUPPORTED, "chain engine: encoder program exceeds the table sizes");
    for (int l = 0; l < L; ++l)
      if (!build_layer_program(h, si, l, h->chain_prog[S + (size_t)si * L + l]))
        return fail(h, DDH_ERR_UNSUPPORTED, "chain engine: layer program exceeds the table sizes");
  }
  return DDH_OK;
}

// forward_test (:578-641) on the scene-tile chain engine: per denoise step one encoder launch, per
// decoder layer plan -> on-demand layout -> conv (+ fused combine) -> chain.
int forward_fused(ddh_handle* h, const float* ego, const float* agents, const void* bev,
                  int bev_dtype, int bev_layout, const float* noise, float* out_traj,
                  float* out_modes, float* out_scores, int64_t* out_mode_idx, int B,
                  cudaStream_t st) {
  const ddh_shape& s = h->shp;
  const int A = s.num_anchors, P = s.num_poses, Na = s.num_agents, L = s.num_layers, S = s.num_steps;
  const int HW = s.bev_h * s.bev_w;
  int rc;
  if (h->chain_prog.empty()) { rc = build_chain_programs(h); if (rc) return rc; }
  const int spt = h->chain_spt, n_tiles = (B + spt - 1) / spt;
  const void* bevn = bev;
  pick_segments(h);
  const bool lazy = bev_layout == DDH_NCHW && h->lazy_layout && h->seg_nw32 > 0 && !h->profiling_eager;
  const int seg_shift = h->seg_px_call == 64 ? 6 : h->seg_px_call == 32 ? 5 : (h->seg_px_call == 16 ? 4 : 3);
  { ProfSpan ps(h, ST_BEV, st);
  if (bev_layout == DDH_NCHW) {
    if (lazy) {
      CU_TRY(h, cudaMemsetAsync(h->done_seg, 0, (size_t)B * h->seg_nw32 * 4, st));
    } else {
      launch_bev_to_nhwc(bev, bev_dtype, h->bev_nhwc, DDH_BF16, B, s.bev_channels, HW, st);
      h->launches++;
    }
    bevn = h->bev_nhwc;
  } else if (bev_dtype != DDH_BF16) {
    launch_cast_f32_bf16(reinterpret_cast<const float*>(bev), reinterpret_cast<__nv_bfloat16*>(h->bev_nhwc),
                         (size_t)B * HW * s.bev_channels, st);
    h->launches++;
    bevn = h->bev_nhwc;
  }
  }
  // hoisted per (scene, layer): agent K|V (bf16, padded rows: the attention's ldmatrix operand) and
  // the collapsed ego vector
  { ProfSpan ps(h, ST_HOIST, st);
  launch_cast_f32_bf16(agents, h->agents16, (size_t)B * Na * D, st);
  launch_cast_f32_bf16(ego, h->ego16, (size_t)B * D, st);
  h->launches += 2;
  for (int l = 0; l < L; ++l) {
    RowEpi e;
    e.out_bf16 = h->kv16 + (size_t)l * B * Na * CH_KV_LD;
    e.ldo16 = CH_KV_LD;
    run_gemm(h, h->layers[l].kv, nullptr, h->agents16, D, B * Na, e, st);
    RowEpi e2;
    e2.out_f32 = h->egov + (size_t)l * B * D;
    e2.ldo32 = D;
    run_gemm(h, h->layers[l].ego, nullptr, h->ego16, D, B, e2, st);
  }
  }
  float* modes = out_modes ? out_modes : h->modes_buf;
  float* scores = out_scores ? out_scores : h->scores_buf;
  OdoConsts oc{s.lidar_max_x, s.lidar_max_y};
  const int conv_mode = (h->persistent_conv >= 2 && A <= 64 && P == 8) ? 2 : (h->persistent_conv ? 1 : 0);
  auto fill = [&](ChainArgs& a) {
    a.B = B; a.A = A; a.Na = Na; a.P = P; a.spt = spt; a.n_tiles = n_tiles;
    a.anchors = h->anchors; a.noise = noise; a.img = h->img; a.pts = h->pts; a.q0t = h->q0t;
    a.modes = modes; a.scores = scores; a.smap = h->smap; a.dbg = nullptr;
    a.logit_part = h->logits_hoisted ? h->logit_part : nullptr;
  };
  const bool reuse = h->conv_reuse && h->vkeep && conv_mode == 2 && S >= 2;
  if (reuse) CU_TRY(h, cudaMemsetAsync(h->new_count, 0, (size_t)S * L * 4, st));
  for (int si = 0; si < S; ++si) {
    { ProfSpan ps(h, ST_EMBED, st);
    ChainArgs& a = h->chain_prog[si];
    fill(a);
    launch_chain(a, st);
    h->launches++; }
    for (int l = 0; l < L; ++l) {
      const PackedLayer& pl = h->layers[l];
      PlanReuse ru;
      __nv_bfloat16* vkeep_l = nullptr;
      if (reuse) {
        ru.mode = si == 0 ? 1 : (h->conv_reuse == 2 ? 3 : 2);
        ru.keep = si + 1 < S;
        ru.slot_tab = h->slot_tab + (size_t)l * B * HW;
        ru.slot_cnt = h->slot_cnt + (size_t)l * B;
        ru.new_list = h->new_list;
        ru.new_count = h->new_count + si * L + l;
        ru.vcap = h->vcap;
        vkeep_l = h->vkeep + (size_t)l * B * h->vcap * D;
      }
      if (h->logits_hoisted) { ru.logit_part = h->logit_part; ru.logit_ld = CH_LOGITS; ru.logit_off = l * P; }
      { ProfSpan ps(h, ST_PLAN, st);
      launch_plan(h->q0t, pl.attw_w, pl.attw_b, h->pts, h->upix, h->nuniq, h->ent_slot, h->ent_w,
                  h->conv_rows + si * L + l, lazy ? h->need_seg : nullptr, h->done_seg, seg_shift,
                  h->seg_nw32, B, A, P, s.bev_h, s.bev_w, h->rcap, oc, st, spt, ru);
      h->launches++; }
      if (lazy) {
        ProfSpan ps(h, ST_BEV, st);
        launch_bev_segs_to_nhwc(bev, bev_dtype, h->bev_nhwc, DDH_BF16, h->need_seg, h->seg_nw32,
                                (h->host_map_call && h->seg_px_call == 16) ? -16 : h->seg_px_call, B,
                                s.bev_channels, s.bev_h, s.bev_w, st);
        h->launches++;
      }
      GemmParams gp;
      gp.K = pl.conv.K;
      gp.bev = bevn; gp.upix = h->upix; gp.nuniq = h->nuniq; gp.rcap = h->rcap;
      gp.H = s.bev_h; gp.W_ = s.bev_w; gp.C = s.bev_channels;
      gp.epi.bias = pl.conv.bias; gp.epi.relu = 1;
      gp.vout = vkeep_l; gp.vcap = h->vcap;
      if (ru.mode < 2) {
        ProfSpan ps(h, ST_CONV, st);
        gp.dbg = (h->conv_timeline == si * L + l) ? h->dbg + 256 : nullptr;
        gp.ent_slot = h->ent_slot; gp.ent_w = h->ent_w; gp.n_anchor = A; gp.ent_per_anchor = P * 4;
        gp.epi.out_f32 = h->s32; gp.epi.ldo32 = D; gp.epi.out_bf16 = h->s16; gp.epi.ldo16 = D;
        if (conv_mode == 2 && h->conv_dynamic) {
          CU_TRY(h, cudaMemsetAsync(h->conv_sched, 0, 4, st));
          gp.sched = h->conv_sched;
        }
        launch_tc_conv(gp, pl.conv.map, B, st, conv_mode);
        h->launches++;
      } else {
        // a later denoise step: value rows of the few pixels no earlier step sampled, then the combine
        // over the kept rows (modules/blocks.py:114 is step-invariant; :117-126 is not)
        { ProfSpan ps(h, ST_CONV_NEW, st);
        gp.vrows = h->new_list; gp.n_vrows = ru.new_count;
        launch_tc_convv(gp, pl.conv.map, st);
        h->launches++; }
        { ProfSpan ps(h, ST_COMBINE, st);
        launch_combine_rows(vkeep_l, h->ent_slot, h->ent_w, h->s16, B, A, P * 4, h->vcap, st);
        h->launches++; }
      }
      { ProfSpan ps(h, ST_GEMM, st);
      ChainArgs& a = h->chain_prog[S + (size_t)si * L + l];
      fill(a);
      if (h->chain_timeline == si * L + l) a.dbg = h->dbg;
      a.kv16 = h->kv16 + (size_t)l * B * Na * CH_KV_LD;
      a.egov = h->egov + (size_t)l * B * D;
      launch_chain(a, st);
      h->launches++; }
    }
  }
  { ProfSpan ps(h, ST_SELECT, st);
  launch_select(scores, modes, out_traj, reinterpret_cast<long long*>(out_mode_idx), B, A, P, st);
  h->launches++; }
  CU_TRY(h, cudaGetLastError());
  return DDH_OK;
}

int forward_range(ddh_handle* h, const float* ego, const float* agents, const void* bev,
                  int bev_dtype, int bev_layout, const float* noise, float* out_traj,
                  float* out_modes, float* out_scores, int64_t* out_mode_idx, int s0, int B,
                  int Btot, cudaStream_t st, cudaEvent_t layout_done) {
  const ddh_shape& s = h->shp;
  const bool bf = h->precision == DDH_PREC_BF16;
  const int A = s.num_anchors, P = s.num_poses, Na = s.num_agents, L = s.num_layers, S = s.num_steps;
  const int M = B * A, F = s.d_ffn, HW = s.bev_h * s.bev_w;
  const int want_dtype = bf ? DDH_BF16 : DDH_F32;
  const View v = make_view(h, s0);
  {
    const size_t z = s0, bev_in_elt = bev_dtype == DDH_BF16 ? 2 : 4;
    ego += z * D; agents += z * Na * D; noise += z * A * P * 2;
    bev = (const unsigned char*)bev + z * HW * s.bev_channels * bev_in_elt;
    out_traj = adv(out_traj, z * P * 3); out_modes = adv(out_modes, z * A * P * 3);
    out_scores = adv(out_scores, z * A); out_mode_idx = adv(out_mode_idx, z);
  }
  // ---- BEV map -> NHWC in the engine's operand type.  NCHW input with H <= 64: rows are
  // converted on demand before each conv call (only rows a conv will read); otherwise up front.
  const void* bevn = bev;
  if (s0 == 0) pick_segments(h);
  const bool lazy = bev_layout == DDH_NCHW && h->lazy_layout && h->seg_nw32 > 0 && !h->profiling_eager;
  const int seg_shift = h->seg_px_call == 64 ? 6 : h->seg_px_call == 32 ? 5 : (h->seg_px_call == 16 ? 4 : 3);
  unsigned int* need_seg = lazy ? h->need_seg + (size_t)s0 * h->seg_nw32 : nullptr;
  unsigned int* done_seg = lazy ? h->done_seg + (size_t)s0 * h->seg_nw32 : nullptr;
  // fp32 engine: value_proj as 3xTF32 on the tensor core.  Its operands are a high / low plane pair of
  // the NHWC fp32 map: written directly by the on-demand layout pass (8- / 16-pixel device segments),
  // or by a split pass behind the eager layout / in front for NHWC input.  A pinned host map read in
  // place (other segment kernels) keeps the CUDA-core conv.
  const bool lazy_split = lazy && (h->seg_px_call == 8 || h->seg_px_call == 16);
  const bool tf32_conv = !bf && h->tf32_conv_ok && h->bev_nhwc_lo && !h->host_map_call && (!lazy || lazy_split);
  void* bev_lo = tf32_conv ? (unsigned char*)h->bev_nhwc_lo + (size_t)s0 * HW * s.bev_channels * 4 : nullptr;
  { ProfSpan ps(h, ST_BEV, st);
  const size_t n_map = (size_t)B * HW * s.bev_channels;
  if (bev_layout == DDH_NCHW) {
    if (lazy) {
      CU_TRY(h, cudaMemsetAsync(done_seg, 0, (size_t)B * h->seg_nw32 * 4, st));
    } else {
      launch_bev_to_nhwc(bev, bev_dtype, v.bev_nhwc, want_dtype, B, s.bev_channels, HW, st);
      h->launches++;
      if (tf32_conv) {
        launch_split_tf32(reinterpret_cast<const float*>(v.bev_nhwc), reinterpret_cast<float*>(v.bev_nhwc),
                          reinterpret_cast<float*>(bev_lo), n_map, st);
        h->launches++;
      }
    }
    bevn = v.bev_nhwc;
  } else if (bev_dtype != want_dtype) {
    if (bf) launch_cast_f32_bf16(reinterpret_cast<const float*>(bev),
                                 reinterpret_cast<__nv_bfloat16*>(v.bev_nhwc), n_map, st);
    else launch_cast_bf16_f32(reinterpret_cast<const __nv_bfloat16*>(bev),
                              reinterpret_cast<float*>(v.bev_nhwc), n_map, st);
    h->launches++;
    bevn = v.bev_nhwc;
    if (tf32_conv) {   // (bf16 values: the low plane is zero)
      launch_split_tf32(reinterpret_cast<const float*>(v.bev_nhwc), reinterpret_cast<float*>(v.bev_nhwc),
                        reinterpret_cast<float*>(bev_lo), n_map, st);
      h->launches++;
    }
  } else if (tf32_conv) {   // NHWC fp32 input: planes of the caller's map into the working copy
    launch_split_tf32(reinterpret_cast<const float*>(bev), reinterpret_cast<float*>(v.bev_nhwc),
                      reinterpret_cast<float*>(bev_lo), n_map, st);
    h->launches++;
    bevn = v.bev_nhwc;
  }
  }

  bool layout_event_pending = layout_done != nullptr;
  if (layout_event_pending && !lazy) {
    CU_TRY(h, cudaEventRecord(layout_done, st));
    layout_event_pending = false;
  }

  // ---- hoisted per (scene, layer): agent K|V and the collapsed ego vector
  { ProfSpan ps(h, ST_HOIST, st);
  if (bf) {
    launch_cast_f32_bf16(agents, v.agents16, (size_t)B * Na * D, st);
    launch_cast_f32_bf16(ego, v.ego16, (size_t)B * D, st);
    h->launches += 2;
  }
  for (int l = 0; l < L; ++l) {
    RowEpi e;
    e.out_f32 = v.kv32 + (size_t)l * Btot * Na * 2 * D;
    e.ldo32 = 2 * D;
    run_gemm(h, h->layers[l].kv, agents, v.agents16, D, B * Na, e, st);
    RowEpi e2;
    e2.out_f32 = v.egov + (size_t)l * Btot * D;
    e2.ldo32 = D;
    run_gemm(h, h->layers[l].ego, ego, v.ego16, D, B, e2, st);
  }
  }

  // ---- truncated noising of the anchors (:591-597)
  const float ac_tr = h->ac[s.trunc_timestep];
  { ProfSpan ps(h, ST_INIT, st);
  launch_init_img(h->anchors, noise, v.img, B, A * P, sqrtf(ac_tr), sqrtf(1.0f - ac_tr), st);
  h->launches++; }

  float* modes = out_modes ? out_modes : v.modes_buf;
  float* scores = out_scores ? out_scores : v.scores_buf;
  OdoConsts oc{s.lidar_max_x, s.lidar_max_y};

  for (int si = 0; si < S; ++si) {
    // clamp, denorm, sine embedding, plan_anchor_encoder (:601-609)
    {
      ProfSpan ps(h, ST_EMBED, st);
      launch_embed(v.img, v.pts, v.emb32, v.emb16, M, P, h->dim_t, st);
      h->launches++;
      RowEpi e;
      e.relu = 1; e.ln1_g = h->enc_ln_g; e.ln1_b = h->enc_ln_b;
      e.out_f32 = v.e1_32; e.ldo32 = D; e.out_bf16 = v.e1_16; e.ldo16 = D;
      run_gemm(h, h->enc0, v.emb32, v.emb16, 64 * P, M, e, st);
      RowEpi e3;
      e3.out_f32 = v.q0_32; e3.ldo32 = D; e3.out_bf16 = v.q0_16; e3.ldo16 = D;
      run_gemm(h, h->enc3, v.e1_32, v.e1_16, D, M, e3, st);
    }
    for (int l = 0; l < L; ++l) {
      const PackedLayer& pl = h->layers[l];
      const bool last_layer = (l == L - 1), last_step = (si == S - 1);
      // -- cross_bev_attention (modules/blocks.py:88-129)
      { ProfSpan ps(h, ST_PLAN, st);
      launch_plan(v.q0_32, pl.attw_w, pl.attw_b, v.pts, v.upix, v.nuniq, v.ent_slot,
                  v.ent_w, h->conv_rows + si * L + l, need_seg, done_seg, seg_shift, h->seg_nw32, B, A,
                  P, s.bev_h, s.bev_w, h->rcap, oc, st); }
      if (lazy) {
        ProfSpan ps(h, ST_BEV, st);
        launch_bev_segs_to_nhwc(bev, bev_dtype, v.bev_nhwc, want_dtype, need_seg, h->seg_nw32, h->seg_px_call,
                                B, s.bev_channels, s.bev_h, s.bev_w, st, tf32_conv ? bev_lo : nullptr);
        h->launches++;
        if (layout_event_pending) {   // the next chunk may start: its layout runs under our conv
          CU_TRY(h, cudaEventRecord(layout_done, st));
          layout_event_pending = false;
        }
      }
      {
        ProfSpan ps(h, ST_CONV, st);
        GemmParams gp;
        gp.K = pl.conv.K;
        gp.bev = bevn; gp.upix = v.upix; gp.nuniq = v.nuniq; gp.rcap = h->rcap;
        gp.H = s.bev_h; gp.W_ = s.bev_w; gp.C = s.bev_channels;
        gp.epi.bias = pl.conv.bias; gp.epi.relu = 1;
        if (bf) {   // combine fused into the conv epilogue: V never leaves the SM
          gp.dbg = h->tl_gemm < 0 ? h->dbg : nullptr;
          gp.ent_slot = v.ent_slot; gp.ent_w = v.ent_w; gp.n_anchor = A; gp.ent_per_anchor = P * 4;
          gp.epi.out_f32 = v.s32; gp.epi.ldo32 = D; gp.epi.out_bf16 = v.s16; gp.epi.ldo16 = D;
          launch_tc_conv(gp, pl.conv.map, B, st, h->persistent_conv != 0 ? 1 : 0);
          h->launches += 1;
        } else if (tf32_conv) {   // fp32 engine, 3xTF32 on the tensor core, fp32 combine fused: S in fp32
          gp.bev_lo = bev_lo;
          gp.ent_slot = v.ent_slot; gp.ent_w = v.ent_w; gp.n_anchor = A; gp.ent_per_anchor = P * 4;
          gp.epi.out_f32 = v.s32; gp.epi.ldo32 = D;
          launch_tc_conv_tf32(gp, pl.conv.map_tf32, B, st);
          h->launches += 1;
        } else {
          gp.epi.out_f32 = v.V; gp.epi.ldo32 = D;
          gp.W = pl.conv.wt32; gp.ldw = D;
          launch_simt_conv(gp, B, st);
        }
      }
      if (!bf && !tf32_conv) {
        ProfSpan ps(h, ST_COMBINE, st);
        launch_combine(v.V, v.ent_slot, v.ent_w, v.s32, v.s16, B, A, P, h->rcap, st);
        h->launches += 3;
      }
      {
        ProfSpan ps(h, ST_GEMM, st);
        RowEpi e;   // output_proj + residual (:127-129)
        e.res = v.q0_32; e.ldres = D;
        e.out_f32 = v.x1_32; e.ldo32 = D; e.out_bf16 = v.x1_16; e.ldo16 = D;
        run_gemm(h, pl.bev_out, v.s32, v.s16, D, M, e, st);
      }
      // -- cross_agent_attention + norm1, cross_ego_attention + norm2 (:355-365)
      {
        ProfSpan ps(h, ST_GEMM, st);
        RowEpi e;
        e.out_f32 = v.qh32; e.ldo32 = D;
        run_gemm(h, pl.q, v.x1_32, v.x1_16, D, M, e, st);
      }
      { ProfSpan ps(h, ST_ATTN, st);
      launch_attn_core(v.qh32, v.kv32 + (size_t)l * Btot * Na * 2 * D, v.o32, v.o16, B, A, Na,
                       s.num_heads, st); }
      h->launches++;
      {
      ProfSpan chain(h, ST_GEMM, st);
      {
        RowEpi e;
        e.res = v.x1_32; e.ldres = D;
        e.ln1_g = pl.norm1_g; e.ln1_b = pl.norm1_b;
        e.rowvec = v.egov + (size_t)l * Btot * D; e.rows_per_group = A;
        e.ln2_g = pl.norm2_g; e.ln2_b = pl.norm2_b;
        e.out_f32 = (bf && !h->debug_taps) ? nullptr : v.x2_32; e.ldo32 = D; e.out_bf16 = v.x2_16; e.ldo16 = D;
        run_gemm(h, pl.attn_out, v.o32, v.o16, D, M, e, st);
      }
      // -- FFN (no residual) + norm3 + time FiLM (:368-373)
      {
        RowEpi e;
        e.relu = 1;
        e.out_f32 = v.h32; e.ldo32 = F; e.out_bf16 = v.h16; e.ldo16 = F;
        run_gemm(h, pl.ffn0, v.x2_32, v.x2_16, D, M, e, st);
        RowEpi e2;
        e2.ln1_g = pl.norm3_g; e2.ln1_b = pl.norm3_b;
        e2.film = h->film + ((size_t)si * L + l) * 2 * D;
        e2.out_f32 = (bf && !h->debug_taps) ? nullptr : v.x3_32; e2.ldo32 = D; e2.out_bf16 = v.x3_16; e2.ldo16 = D;
        run_gemm(h, pl.ffn2, v.h32, v.h16, F, M, e2, st);
      }
      // -- task_decoder (:244-256, 376-380); cls only where it is read (:631)
      if (last_layer && last_step) {
        RowEpi e;
        e.relu = 1; e.ln1_g = pl.cls_ln2_g; e.ln1_b = pl.cls_ln2_b;
        e.out_f32 = v.c1_32; e.ldo32 = D; e.out_bf16 = v.c1_16; e.ldo16 = D;
        run_gemm(h, pl.cls0, v.x3_32, v.x3_16, D, M, e, st);
        RowEpi e2;
        e2.relu = 1; e2.ln1_g = pl.cls_ln5_g; e2.ln1_b = pl.cls_ln5_b;
        e2.dot_w = pl.cls6_w; e2.dot_b = pl.cls6_b; e2.dot_out = scores;
        run_gemm(h, pl.cls3, v.c1_32, v.c1_16, D, M, e2, st);
      }
      {
        RowEpi e;
        e.relu = 1;
        e.out_f32 = v.r1_32; e.ldo32 = D; e.out_bf16 = v.r1_16; e.ldo16 = D;
        run_gemm(h, pl.reg0, v.x3_32, v.x3_16, D, M, e, st);
        RowEpi e2;
        e2.relu = 1;
        e2.out_f32 = v.r2_32; e2.ldo32 = D;
        run_gemm(h, pl.reg2, v.r1_32, v.r1_16, D, M, e2, st);
      }
      }
      DdimCoef dc{0.f, 1.f, 1.f, 0.f};
      const int do_ddim = (last_layer && !last_step) ? 1 : 0;
      if (do_ddim) {
        const int t = h->roll[si], prev = t - 1;   // set_timesteps(1000) => step ratio 1 (:584)
        const float ac_t = h->ac[t], ac_p = prev >= 0 ? h->ac[prev] : 1.0f;
        dc.sqrt_ac_t = sqrtf(ac_t); dc.sqrt_1m_ac_t = sqrtf(1.0f - ac_t);
        dc.sqrt_ac_prev = sqrtf(ac_p); dc.sqrt_1m_ac_prev = sqrtf(1.0f - ac_p);
      }
      { ProfSpan ps(h, ST_REG, st);
      launch_reg_finish(v.r2_32, pl.reg4_w, pl.reg4_b, v.pts, v.img, modes, M, P, do_ddim, dc, st); }
      h->launches++;
    }
  }
  { ProfSpan ps(h, ST_SELECT, st);
  launch_select(scores, modes, out_traj, reinterpret_cast<long long*>(out_mode_idx), B, A, P, st); }
  h->launches++;
  CU_TRY(h, cudaGetLastError());
  return DDH_OK;
}

}  // namespace

extern "C" {

int ddh_forward(ddh_handle* h, const float* ego, const float* agents, const void* bev,
                int bev_dtype, int bev_layout, const float* noise, float* out_traj,
                float* out_modes, float* out_scores, int64_t* out_mode_idx, int B, void* stream) {
  if (!h) return DDH_ERR_BAD_ARG;
  if (!h->packed) return fail(h, DDH_ERR_NOT_PACKED, "ddh_forward: weights not packed");
  if (!ego || !agents || !bev || !noise || B <= 0)
    return fail(h, DDH_ERR_BAD_ARG, "ddh_forward: null input or B <= 0");
  if ((bev_dtype != DDH_F32 && bev_dtype != DDH_BF16) || (bev_layout != DDH_NCHW && bev_layout != DDH_NHWC))
    return fail(h, DDH_ERR_BAD_ARG, "ddh_forward: bad bev dtype/layout");
  if ((reinterpret_cast<uintptr_t>(bev) | reinterpret_cast<uintptr_t>(ego) |
       reinterpret_cast<uintptr_t>(agents) | reinterpret_cast<uintptr_t>(noise) |
       reinterpret_cast<uintptr_t>(out_modes)) & 15)
    return fail(h, DDH_ERR_ALIGNMENT, "ddh_forward: ego/agents/bev/noise/out_modes must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int rc = ensure_ws(h, B);
  if (rc) return rc;
  h->last_B = B;
  h->launches = 0;
  const ddh_shape& s = h->shp;
  const bool resident = h->res2_ok && !h->profiling && h->precision == DDH_PREC_BF16 && B <= RES_MAX_B;
  if (!resident || h->debug_taps) register_taps(h, B);   // (string-keyed map: kept off the one-launch path)
  h->ev_used = 0;
  h->ev_spans.clear();

  // <= RES_MAX_B scenes: the whole forward as ONE launch on one 16-CTA cluster per scene,
  // activations resident on the SM that owns the anchor (kernels_res2.cu)
  if (resident) {
    ResCall call;
    call.ego = ego; call.agents = agents; call.bev = bev; call.bev_dtype = bev_dtype == DDH_BF16 ? 1 : 0;
    call.bev_nhwc_bf16 = (bev_layout == DDH_NHWC) ? 1 : 0;
    h->launches = 0;
    if (bev_layout == DDH_NHWC && bev_dtype != DDH_BF16) {   // NHWC fp32: one cast pass in front
      launch_cast_f32_bf16(reinterpret_cast<const float*>(bev), h->res2_host.bev_nhwc,
                           (size_t)B * s.bev_h * s.bev_w * s.bev_channels, st);
      call.bev = h->res2_host.bev_nhwc;
      call.bev_dtype = 1;
      h->launches++;
    }
    call.noise = noise; call.out_traj = out_traj;
    call.out_modes = out_modes ? out_modes : h->modes_buf;
    call.out_scores = out_scores ? out_scores : h->scores_buf;
    call.out_mode_idx = reinterpret_cast<long long*>(out_mode_idx);
    call.dbg = h->debug_taps ? h->dbg : nullptr;
    // dense mode (tiny batches): the launch carries helper clusters that run value_proj over the whole
    // map while the scene clusters do the embedding / encoder; not when the map is read in place
    // from pinned host memory (the helpers would pull all of it across PCIe)
    const DenseArgs* dense = nullptr;
    if (h->dense_ok && B <= h->dense_max_b && B <= DENSE_MAX_B && !h->host_map_call) {
      DenseArgs& da = h->dense_host;
      const void* base = call.bev_nhwc_bf16 ? call.bev : h->res2_host.bev_nhwc;
      if (base != h->dense_amap_base) {
        rc = encode_nhwc_map(h, &da.amap, base, call.bev_nhwc_bf16 ? B : RES_MAX_B, s.bev_h, s.bev_w);
        if (rc) return rc;
        h->dense_amap_base = call.bev_nhwc_bf16 ? nullptr : base;   // a caller's map is re-described every call (its extent is B)
      }
      da.nhwc = call.bev_nhwc_bf16 ? nullptr : h->res2_host.bev_nhwc;
      da.B = B;
      da.n_helper_ctas = RES_CL * std::max(1, std::min(8, res2_max_clusters() - B));
      dense = &da;
    }
    const int e = launch_res2_forward(h->res2_consts, call, B, st, dense);
    if (e) return fail(h, DDH_ERR_CUDA, std::string("res2_forward launch: ") + cudaGetErrorString((cudaError_t)e));
    h->launches++;
    if (h->debug_taps) {
      const R2Consts& R = h->res2_host;
      const size_t MA = (size_t)B * s.num_anchors;
      h->taps["res_q0"] = {R.tap_q0, MA * D * 4}; h->taps["res_x1"] = {R.tap_x1, MA * D * 4};
      h->taps["res_regraw"] = {R.tap_regraw, MA * 3 * s.num_poses * 4};
      h->taps["res_kv"] = {R.kv, (size_t)B * s.num_layers * s.num_agents * 2 * D * 4};
      h->taps["res_egov"] = {R.egov, (size_t)B * s.num_layers * D * 4};
      if (dense) h->taps["dense_v"] = {dense->V, (size_t)B * s.num_layers * s.bev_h * s.bev_w * D * 2};
      else h->taps.erase("dense_v");
    }
    CU_TRY(h, cudaGetLastError());
    return DDH_OK;
  }
  CU_TRY(h, cudaMemsetAsync(h->conv_rows, 0, (size_t)s.num_layers * s.num_steps * 4, st));
  if (h->chain_ok && h->precision == DDH_PREC_BF16) {
    rc = forward_fused(h, ego, agents, bev, bev_dtype, bev_layout, noise, out_traj, out_modes, out_scores,
                       out_mode_idx, B, st);
    if (rc) return rc;
    CU_TRY(h, cudaGetLastError());
    return DDH_OK;
  }
  // Scene chunks on two streams: scenes are independent, so chunk c+1's HBM-bound layout pass
  // runs under chunk c's tensor-bound conv/GEMMs.  Chunk c starts once layout(c-1) is done.
  int nchunk = 1;
  if (!h->profiling && h->chunks > 1 && B >= h->chunks * h->min_chunk_scenes) nchunk = h->chunks;
  if (nchunk == 1) {
    rc = forward_range(h, ego, agents, bev, bev_dtype, bev_layout, noise, out_traj, out_modes,
                       out_scores, out_mode_idx, 0, B, B, st, nullptr);
    if (rc) return rc;
  } else {
    if (!h->aux_stream) CU_TRY(h, cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    while ((int)h->sync_events.size() < nchunk + 2) {
      cudaEvent_t e;
      CU_TRY(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->sync_events.push_back(e);
    }
    cudaEvent_t ev_fork = h->sync_events[nchunk], ev_join = h->sync_events[nchunk + 1];
    CU_TRY(h, cudaEventRecord(ev_fork, st));
    CU_TRY(h, cudaStreamWaitEvent(h->aux_stream, ev_fork, 0));
    const int base = B / nchunk, rem = B % nchunk;
    int lo = 0, launches = 0;
    for (int c = 0; c < nchunk; ++c) {
      const int n = base + (c < rem ? 1 : 0);
      cudaStream_t sc = (c & 1) ? h->aux_stream : st;
      if (c > 0) CU_TRY(h, cudaStreamWaitEvent(sc, h->sync_events[c - 1], 0));
      h->launches = 0;
      rc = forward_range(h, ego, agents, bev, bev_dtype, bev_layout, noise, out_traj, out_modes,
                         out_scores, out_mode_idx, lo, n, B, sc, h->sync_events[c]);
      if (rc) return rc;
      launches += h->launches;
      lo += n;
    }
    h->launches = launches;
    CU_TRY(h, cudaEventRecord(ev_join, h->aux_stream));
    CU_TRY(h, cudaStreamWaitEvent(st, ev_join, 0));
  }
  CU_TRY(h, cudaGetLastError());
  return DDH_OK;
}

int ddh_forward_host(ddh_handle* h, const float* ego, const float* agents, const void* bev,
                     int bev_dtype, int bev_layout, const float* noise, float* out_traj,
                     float* out_modes, float* out_scores, int64_t* out_mode_idx, int B,
                     void* stream) {
  if (!h) return DDH_ERR_BAD_ARG;
  if (!h->packed) return fail(h, DDH_ERR_NOT_PACKED, "ddh_forward_host: weights not packed");
  if (!ego || !agents || !bev || !noise || B <= 0)
    return fail(h, DDH_ERR_BAD_ARG, "ddh_forward_host: null input or B <= 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const ddh_shape& s = h->shp;
  const int A = s.num_anchors, P = s.num_poses, Na = s.num_agents;
  const size_t bev_bytes = (size_t)B * s.bev_channels * s.bev_h * s.bev_w * (bev_dtype == DDH_BF16 ? 2 : 4);
  // Pinned (device-mapped) NCHW maps are not copied: the on-demand layout pass (or the resident
  // engine's in-kernel one) reads only the segments the conv calls need straight across PCIe,
  // ~20 % of the map instead of all of it.
  const void* bev_dev = nullptr;
  h->host_map_call = false;
  const bool resident_b = h->res2_ok && h->precision == DDH_PREC_BF16 && B <= RES_MAX_B;
  if (h->host_zero_copy && bev_layout == DDH_NCHW && h->lazy_layout) {
    const bool resident = resident_b;
    const bool lazy_ok = (s.bev_w % 8 == 0 && s.bev_h * (s.bev_w / std::min(64, (int)s.bev_w)) <= 2048);
    cudaPointerAttributes at;
    if ((resident || lazy_ok) && cudaPointerGetAttributes(&at, bev) == cudaSuccess &&
        at.type == cudaMemoryTypeHost && at.devicePointer != nullptr)
      bev_dev = at.devicePointer;
    else
      cudaGetLastError();   // pageable memory: clear the sticky "invalid value"
    h->host_map_call = bev_dev != nullptr;
  }
  const size_t bev_stage = bev_dev ? 0 : bev_bytes;
  if (B > h->host_cap_B || bev_stage > h->host_bev_bytes) {
    cudaDeviceSynchronize();
    free_all(h->owned_host);
    h->host_cap_B = 0;
    int rc;
    auto& o = h->owned_host;
    unsigned char* pb = nullptr;
#define HS(ptr, count) do { rc = dev_alloc(h, o, &(ptr), (size_t)(count)); if (rc) return rc; } while (0)
    HS(h->hs_ego, (size_t)B * D); HS(h->hs_agents, (size_t)B * Na * D);
    HS(pb, bev_stage); h->hs_bev = pb;
    HS(h->hs_noise, (size_t)B * A * P * 2); HS(h->hs_traj, (size_t)B * P * 3);
    HS(h->hs_modes, (size_t)B * A * P * 3); HS(h->hs_scores, (size_t)B * A); HS(h->hs_idx, B);
#undef HS
    h->host_cap_B = B;
    h->host_bev_bytes = bev_stage;
  }
  CU_TRY(h, cudaMemcpyAsync(h->hs_ego, ego, (size_t)B * D * 4, cudaMemcpyHostToDevice, st));
  CU_TRY(h, cudaMemcpyAsync(h->hs_agents, agents, (size_t)B * Na * D * 4, cudaMemcpyHostToDevice, st));
  if (!bev_dev) {
    CU_TRY(h, cudaMemcpyAsync(h->hs_bev, bev, bev_bytes, cudaMemcpyHostToDevice, st));
    bev_dev = h->hs_bev;
  }
  CU_TRY(h, cudaMemcpyAsync(h->hs_noise, noise, (size_t)B * A * P * 2 * 4, cudaMemcpyHostToDevice, st));
  int rc = ddh_forward(h, h->hs_ego, h->hs_agents, bev_dev, bev_dtype, bev_layout, h->hs_noise,
                       h->hs_traj, h->hs_modes, h->hs_scores, reinterpret_cast<int64_t*>(h->hs_idx),
                       B, stream);
  h->host_map_call = false;
  if (rc) return rc;
  if (out_traj) CU_TRY(h, cudaMemcpyAsync(out_traj, h->hs_traj, (size_t)B * P * 3 * 4, cudaMemcpyDeviceToHost, st));
  if (out_modes) CU_TRY(h, cudaMemcpyAsync(out_modes, h->hs_modes, (size_t)B * A * P * 3 * 4, cudaMemcpyDeviceToHost, st));
  if (out_scores) CU_TRY(h, cudaMemcpyAsync(out_scores, h->hs_scores, (size_t)B * A * 4, cudaMemcpyDeviceToHost, st));
  if (out_mode_idx) CU_TRY(h, cudaMemcpyAsync(out_mode_idx, h->hs_idx, (size_t)B * 8, cudaMemcpyDeviceToHost, st));
  CU_TRY(h, cudaStreamSynchronize(st));
  return DDH_OK;
}

size_t ddh_bev_producer_scratch_bytes(int B, int grid, int bev_channels) {
  if (B <= 0 || grid <= 0 || bev_channels <= 0) return 0;
  return bev_producer_scratch_bytes(B, grid, bev_channels);
}

int ddh_bev_producer(const float* keyval_tokens, const float* bev_map, const float* weight,
                     const float* bias, const float* ln_weight, const float* ln_bias, void* out,
                     int out_dtype, int B, int H, int W, int grid, int bev_channels, void* scratch,
                     void* stream) {
  if (!keyval_tokens || !bev_map || !weight || !bias || !ln_weight || !ln_bias || !out || !scratch || B <= 0)
    return fail(nullptr, DDH_ERR_BAD_ARG, "ddh_bev_producer: null argument or B <= 0");
  if (out_dtype != DDH_F32 && out_dtype != DDH_BF16)
    return fail(nullptr, DDH_ERR_BAD_ARG, "ddh_bev_producer: bad output dtype");
  if (H <= 0 || W <= 0 || W % 32 || (H * W) % 32 || grid < 1 || grid > 64 || bev_channels < 1 ||
      bev_channels > 128 || bev_channels % 4)
    return fail(nullptr, DDH_ERR_UNSUPPORTED, "ddh_bev_producer: need W % 32 == 0, grid in [1, 64], bev_channels % 4 == 0 and <= 128");
  if ((reinterpret_cast<uintptr_t>(keyval_tokens) | reinterpret_cast<uintptr_t>(bev_map) |
       reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(scratch) |
       reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(ln_weight) | reinterpret_cast<uintptr_t>(ln_bias)) & 15)
    return fail(nullptr, DDH_ERR_ALIGNMENT, "ddh_bev_producer: pointers must be 16-byte aligned");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, DDH_ERR_CUDA, "ddh_bev_producer: no CUDA device (there is no CPU fallback)");
  const int e = launch_bev_producer(keyval_tokens, bev_map, weight, bias, ln_weight, ln_bias, out,
                                    out_dtype == DDH_BF16 ? 1 : 0, B, H, W, grid, bev_channels,
                                    reinterpret_cast<float*>(scratch), reinterpret_cast<cudaStream_t>(stream));
  if (e) return fail(nullptr, DDH_ERR_CUDA, std::string("ddh_bev_producer: ") + cudaGetErrorString((cudaError_t)e));
  return DDH_OK;
}

// ---- query decoder + AgentHead (row N3): the head's GEMM engines + a small attention kernel ----
struct QdecLayer {
  PackedLinear self_in, self_out, cross_q, cross_kv, cross_out, lin1, lin2;
  float *n1g = nullptr, *n1b = nullptr, *n2g = nullptr, *n2b = nullptr, *n3g = nullptr, *n3b = nullptr;
};
struct ddh_qdec {
  ddh_handle ctx;            // precision, packed-weight ownership, TMA encoder, error text
  ddh_qdec_shape shp;
  std::vector<QdecLayer> layers;
  PackedLinear states0;
  float *qemb = nullptr, *states2_w = nullptr, *states2_b = nullptr, *label_w = nullptr, *label_b = nullptr;
  int cap_B = 0;
  std::vector<void*> ws;
  float *x32 = nullptr, *y32 = nullptr, *qkv32 = nullptr, *o32 = nullptr, *qc32 = nullptr, *kvc32 = nullptr, *h32 = nullptr;
  __nv_bfloat16 *x16 = nullptr, *y16 = nullptr, *o16 = nullptr, *mem16 = nullptr, *h16 = nullptr;
};

int ddh_qdec_create(const ddh_qdec_shape* s, ddh_qdec** out) {
  if (!s || !out) return fail(nullptr, DDH_ERR_BAD_ARG, "ddh_qdec_create: null argument");
  *out = nullptr;
  if (s->d_model != 256 || s->num_heads != 8 || s->num_queries < 1 || s->num_queries > 32 ||
      s->num_keys < 1 || s->num_keys > 96 || s->d_ffn <= 0 || s->d_ffn % 256 || s->num_layers < 1 ||
      s->num_layers > 16)
    return fail(nullptr, DDH_ERR_UNSUPPORTED,
                "ddh_qdec_create: need d_model 256, 8 heads, <= 32 queries, <= 96 keys, d_ffn % 256 == 0");
  ddh_qdec* q = new ddh_qdec();
  q->shp = *s;
  memset(&q->ctx.shp, 0, sizeof q->ctx.shp);
  *out = q;
  return DDH_OK;
}

void ddh_qdec_destroy(ddh_qdec* q) {
  if (!q) return;
  free_all(q->ctx.owned_w);
  free_all(q->ws);
  delete q;
}

const char* ddh_qdec_last_error(const ddh_qdec* q) { return q ? q->ctx.err.c_str() : g_create_error.c_str(); }

int ddh_qdec_pack_weights(ddh_qdec* q, const ddh_qdec_weight_ptrs* w, int precision, void* stream) {
  if (!q || !w || !w->layers) return fail(q ? &q->ctx : nullptr, DDH_ERR_BAD_ARG, "ddh_qdec_pack_weights: null argument");
  ddh_handle* h = &q->ctx;
  if (precision != DDH_PREC_FP32 && precision != DDH_PREC_BF16)
    return fail(h, DDH_ERR_BAD_ARG, "ddh_qdec_pack_weights: unknown precision");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(h, DDH_ERR_CUDA, "ddh_qdec_pack_weights: no CUDA device (there is no CPU fallback)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  cudaDeviceSynchronize();
  free_all(h->owned_w);
  free_all(q->ws);
  q->cap_B = 0;
  h->packed = false;
  h->precision = precision;
  if (precision == DDH_PREC_BF16 && !h->tc_ready) {
    int e = tc_engine_init();
    if (e) return fail(h, DDH_ERR_CUDA, std::string("tc_engine_init: ") + cudaGetErrorString((cudaError_t)e));
    h->tc_ready = true;
  }
  const ddh_qdec_shape& s = q->shp;
  const int F = s.d_ffn, Q = s.num_queries;
  int rc;
#define TRY(x) do { rc = (x); if (rc) return rc; } while (0)
  TRY(copy_vec(h, &q->qemb, w->query_embedding, (size_t)Q * D, st));
  q->layers.assign(s.num_layers, QdecLayer());
  for (int l = 0; l < s.num_layers; ++l) {
    const ddh_qdec_layer_weights& lw = w->layers[l];
    QdecLayer& pl = q->layers[l];
    TRY(pack_linear(h, pl.self_in, lw.self_in_w, lw.self_in_b, 3 * D, D, st));
    TRY(pack_linear(h, pl.self_out, lw.self_out_w, lw.self_out_b, D, D, st));
    TRY(pack_linear(h, pl.cross_q, lw.cross_in_w, lw.cross_in_b, D, D, st));
    TRY(pack_linear(h, pl.cross_kv, lw.cross_in_w + (size_t)D * D, lw.cross_in_b + D, 2 * D, D, st));
    TRY(pack_linear(h, pl.cross_out, lw.cross_out_w, lw.cross_out_b, D, D, st));
    TRY(pack_linear(h, pl.lin1, lw.lin1_w, lw.lin1_b, F, D, st));
    TRY(pack_linear(h, pl.lin2, lw.lin2_w, lw.lin2_b, D, F, st));
    TRY(copy_vec(h, &pl.n1g, lw.norm1_w, D, st)); TRY(copy_vec(h, &pl.n1b, lw.norm1_b, D, st));
    TRY(copy_vec(h, &pl.n2g, lw.norm2_w, D, st)); TRY(copy_vec(h, &pl.n2b, lw.norm2_b, D, st));
    TRY(copy_vec(h, &pl.n3g, lw.norm3_w, D, st)); TRY(copy_vec(h, &pl.n3b, lw.norm3_b, D, st));
  }
  TRY(pack_linear(h, q->states0, w->states0_w, w->states0_b, F, D, st));
  TRY(copy_vec(h, &q->states2_w, w->states2_w, (size_t)5 * F, st));
  TRY(copy_vec(h, &q->states2_b, w->states2_b, 5, st));
  TRY(copy_vec(h, &q->label_w, w->label_w, D, st));
  TRY(copy_vec(h, &q->label_b, w->label_b, 1, st));
#undef TRY
  CU_TRY(h, cudaGetLastError());
  h->packed = true;
  return DDH_OK;
}

int ddh_qdec_forward(ddh_qdec* q, const float* keyval, float* query_out, float* agent_states,
                     float* agent_labels, int B, void* stream) {
  if (!q) return DDH_ERR_BAD_ARG;
  ddh_handle* h = &q->ctx;
  if (!h->packed) return fail(h, DDH_ERR_NOT_PACKED, "ddh_qdec_forward: weights not packed");
  if (!keyval || B <= 0) return fail(h, DDH_ERR_BAD_ARG, "ddh_qdec_forward: null input or B <= 0");
  if (reinterpret_cast<uintptr_t>(keyval) & 15) return fail(h, DDH_ERR_ALIGNMENT, "ddh_qdec_forward: keyval must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const ddh_qdec_shape& s = q->shp;
  const int Q = s.num_queries, Nk = s.num_keys, F = s.d_ffn, M = B * Q, MK = B * Nk;
  const bool bf = h->precision == DDH_PREC_BF16;
  int rc;
  if (B > q->cap_B) {
    cudaDeviceSynchronize();
    free_all(q->ws);
    q->cap_B = 0;
#define WS(ptr, count) do { rc = dev_alloc(h, q->ws, &(ptr), (size_t)(count)); if (rc) return rc; } while (0)
    WS(q->x32, (size_t)M * D); WS(q->y32, (size_t)M * D); WS(q->qkv32, (size_t)M * 3 * D); WS(q->o32, (size_t)M * D);
    WS(q->qc32, (size_t)M * D); WS(q->kvc32, (size_t)MK * 2 * D); WS(q->h32, (size_t)M * F);
    if (bf) {
      WS(q->x16, (size_t)M * D); WS(q->y16, (size_t)M * D); WS(q->o16, (size_t)M * D);
      WS(q->mem16, (size_t)MK * D); WS(q->h16, (size_t)M * F);
    } else {
      q->x16 = q->y16 = q->o16 = q->mem16 = q->h16 = nullptr;
    }
#undef WS
    q->cap_B = B;
  }
  h->launches = 0;
  launch_broadcast_rows(q->qemb, q->x32, q->x16, Q, (size_t)M * D, st);
  h->launches++;
  if (bf) { launch_cast_f32_bf16(keyval, q->mem16, (size_t)MK * D, st); h->launches++; }
  float *x32 = q->x32, *y32 = q->y32;
  __nv_bfloat16 *x16 = q->x16, *y16 = q->y16;
  for (int l = 0; l < s.num_layers; ++l) {
    const QdecLayer& pl = q->layers[l];
    // x = norm1(x + self_attn(x, x, x))
    { RowEpi e; e.out_f32 = q->qkv32; e.ldo32 = 3 * D;
      run_gemm(h, pl.self_in, x32, x16, D, M, e, st); }
    rc = launch_mha_small(q->qkv32, 3 * D, q->qkv32 + D, 3 * D, q->o32, q->o16, B, Q, Q, st);
    if (rc) return fail(h, DDH_ERR_CUDA, "ddh_qdec_forward: attention launch failed");
    h->launches++;
    { RowEpi e; e.res = x32; e.ldres = D; e.ln1_g = pl.n1g; e.ln1_b = pl.n1b;
      e.out_f32 = y32; e.ldo32 = D; e.out_bf16 = y16; e.ldo16 = D;
      run_gemm(h, pl.self_out, q->o32, q->o16, D, M, e, st); }
    // y = norm2(y + multihead_attn(y, memory, memory))
    { RowEpi e; e.out_f32 = q->qc32; e.ldo32 = D;
      run_gemm(h, pl.cross_q, y32, y16, D, M, e, st); }
    { RowEpi e; e.out_f32 = q->kvc32; e.ldo32 = 2 * D;
      run_gemm(h, pl.cross_kv, keyval, q->mem16, D, MK, e, st); }
    rc = launch_mha_small(q->qc32, D, q->kvc32, 2 * D, q->o32, q->o16, B, Q, Nk, st);
    if (rc) return fail(h, DDH_ERR_CUDA, "ddh_qdec_forward: attention launch failed");
    h->launches++;
    { RowEpi e; e.res = y32; e.ldres = D; e.ln1_g = pl.n2g; e.ln1_b = pl.n2b;
      e.out_f32 = x32; e.ldo32 = D; e.out_bf16 = x16; e.ldo16 = D;
      run_gemm(h, pl.cross_out, q->o32, q->o16, D, M, e, st); }
    // x = norm3(x + linear2(relu(linear1(x))))
    { RowEpi e; e.relu = 1; e.out_f32 = bf ? nullptr : q->h32; e.ldo32 = F; e.out_bf16 = q->h16; e.ldo16 = F;
      run_gemm(h, pl.lin1, x32, x16, D, M, e, st); }
    { RowEpi e; e.res = x32; e.ldres = D; e.ln1_g = pl.n3g; e.ln1_b = pl.n3b;
      e.out_f32 = y32; e.ldo32 = D; e.out_bf16 = y16; e.ldo16 = D;
      run_gemm(h, pl.lin2, q->h32, q->h16, F, M, e, st); }
    std::swap(x32, y32);
    std::swap(x16, y16);
  }
  if (query_out) CU_TRY(h, cudaMemcpyAsync(query_out, x32, (size_t)M * D * 4, cudaMemcpyDeviceToDevice, st));
  if (agent_states) {
    RowEpi e; e.relu = 1; e.out_f32 = q->h32; e.ldo32 = F;
    run_gemm(h, q->states0, x32, x16, D, M, e, st);
    launch_rowdot(q->h32, F, q->states2_w, q->states2_b, agent_states, M, F, 5, Q, 1, 1, st);
    h->launches++;
  }
  if (agent_labels) {
    launch_rowdot(x32, D, q->label_w, q->label_b, agent_labels, M, D, 1, Q, 1, 0, st);
    h->launches++;
  }
  CU_TRY(h, cudaGetLastError());
  return DDH_OK;
}

int ddh_last_launch_count(const ddh_handle* h) { return h ? h->launches : 0; }

int ddh_set_concurrency(ddh_handle* h, int chunks, int min_chunk_scenes) {
  if (!h || chunks < 1 || chunks > 64 || min_chunk_scenes < 1)
    return fail(h, DDH_ERR_BAD_ARG, "ddh_set_concurrency: bad argument");
  h->chunks = chunks;
  h->min_chunk_scenes = min_chunk_scenes;
  return DDH_OK;
}

int ddh_set_option(ddh_handle* h, const char* name, int value) {
  if (!h || !name) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: null argument");
  const std::string n(name);
  bool repack = false;
  if (n == "lazy_layout") h->lazy_layout = value;
  else if (n == "chain_engine") { repack = h->chain_enabled != value; h->chain_enabled = value; }
  else if (n == "resident_engine") { repack = h->res_mode != (value ? 2 : 0); h->res_mode = value ? 2 : 0; }
  else if (n == "debug_taps") h->debug_taps = value != 0;
  else if (n == "dense_conv") {
    if (value < 0 || value > DENSE_MAX_B) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: dense_conv must be 0.." + std::to_string(DENSE_MAX_B));
    h->dense_max_b = value;
  }
  else if (n == "chain_timeline") h->chain_timeline = value;
  else if (n == "persistent_conv") {
    if ((h->persistent_conv >= 2) != (value >= 2)) { cudaDeviceSynchronize(); free_all(h->owned_ws); h->cap_B = 0; h->vkeep = nullptr; h->chain_prog.clear(); }
    h->persistent_conv = value;
  }
  else if (n == "conv_timeline") h->conv_timeline = value;
  else if (n == "conv_dynamic") h->conv_dynamic = value;
  else if (n == "conv_reuse") {
    if (value < 0 || value > 2) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: conv_reuse must be 0, 1 or 2");
    if ((h->conv_reuse != 0) != (value != 0)) { cudaDeviceSynchronize(); free_all(h->owned_ws); h->cap_B = 0; h->vkeep = nullptr; h->chain_prog.clear(); }
    h->conv_reuse = value;
  }
  else if (n == "fp32_tensor_conv") { repack = h->fp32_tensor_conv != value; h->fp32_tensor_conv = value; }
  else if (n == "host_zero_copy") h->host_zero_copy = value;
  else if (n == "host_segment") {
    if (value != 16 && value != 32 && value != 64) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: host_segment must be 16, 32 or 64");
    h->host_seg_px = value;
  }
  else if (n == "layout_segment") {
    if (value != 8 && value != 16) return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: layout_segment must be 8 or 16");
    if (h->seg_px != value) { h->seg_px = value; cudaDeviceSynchronize(); free_all(h->owned_ws); h->cap_B = 0; h->chain_prog.clear(); }
  }
  else if (n == "timeline_gemm") h->tl_gemm = value;
  else return fail(h, DDH_ERR_BAD_ARG, "ddh_set_option: unknown option " + n);
  if (repack) h->packed = false;   // engine selection is fixed at pack time: the caller packs again
  return DDH_OK;
}

int ddh_set_profiling(ddh_handle* h, int on) {
  if (!h) return DDH_ERR_BAD_ARG;
  h->profiling = on != 0;
  return DDH_OK;
}

int ddh_get_profile(ddh_handle* h, const char* stage, float* total_ms, int* spans) {
  if (!h || !stage || !total_ms || !spans) return DDH_ERR_BAD_ARG;
  int id = -1;
  for (int i = 0; i < kNumStages; ++i) if (!strcmp(stage, kStageNames[i])) id = i;
  if (id < 0) return fail(h, DDH_ERR_BAD_ARG, std::string("ddh_get_profile: unknown stage ") + stage);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(h, DDH_ERR_CUDA, std::string("ddh_get_profile: ") + cudaGetErrorString(e));
  float tot = 0.f;
  int n = 0;
  for (auto& sp : h->ev_spans) {
    if (sp.first != id) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_pool[sp.second], h->ev_pool[sp.second + 1]) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  *total_ms = tot;
  *spans = n;
  return DDH_OK;
}

long long ddh_debug_copy(ddh_handle* h, const char* name, void* host_dst, size_t max_bytes) {
  if (!h || !name || !host_dst) return DDH_ERR_BAD_ARG;
  auto it = h->taps.find(name);
  if (it == h->taps.end() || !it->second.first)
    return fail(h, DDH_ERR_BAD_ARG, std::string("ddh_debug_copy: unknown tap ") + name);
  const size_t n = std::min(max_bytes, it->second.second);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(host_dst, it->second.first, n, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return fail(h, DDH_ERR_CUDA, std::string("ddh_debug_copy: ") + cudaGetErrorString(e));
  return (long long)n;
}

int ddh_test_gemm(ddh_handle* h, const float* A, const float* W, const float* bias, float* C,
                  int M, int N, int K, int precision, void* stream) {
  if (!h || !A || !W || !C) return fail(h, DDH_ERR_BAD_ARG, "ddh_test_gemm: null argument");
  if (N % 256 || K % 64 || M <= 0) return fail(h, DDH_ERR_UNSUPPORTED, "ddh_test_gemm: need N%256==0, K%64==0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::vector<void*> tmp;
  int rc = DDH_OK;
  GemmParams p;
  p.lda = K; p.M = M; p.K = K;
  p.epi.bias = bias; p.epi.out_f32 = C; p.epi.ldo32 = N;
  if (precision == DDH_PREC_FP32) {
    float* wt;
    rc = dev_alloc(h, tmp, &wt, (size_t)N * K);
    if (!rc) {
      launch_transpose_f32(W, wt, N, K, st);
      p.A = A; p.W = wt; p.ldw = N;
      launch_simt_gemm(p, N, st);
    }
  } else {
    if (!h->tc_ready) {
      int e = tc_engine_init();
      if (e) return fail(h, DDH_ERR_CUDA, std::string("tc_engine_init: ") + cudaGetErrorString((cudaError_t)e));
      h->tc_ready = true;
    }
    PackedLinear L;
    L.N = N; L.K = K;
    __nv_bfloat16* a16;
    rc = dev_alloc(h, tmp, &a16, (size_t)M * K);
    if (!rc) rc = dev_alloc(h, tmp, &L.w16, (size_t)N * K);
    if (!rc) {
      launch_cast_f32_bf16(A, a16, (size_t)M * K, st);
      launch_cast_f32_bf16(W, L.w16, (size_t)N * K, st);
      rc = make_wmap(h, L);
    }
    if (!rc) {
      p.A = a16;
      launch_tc_gemm(p, L.map, N, st);
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  if (!rc && e == cudaSuccess) e = cudaGetLastError();
  free_all(tmp);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, DDH_ERR_CUDA, std::string("ddh_test_gemm: ") + cudaGetErrorString(e));
  return DDH_OK;
}

}  // extern "C"
