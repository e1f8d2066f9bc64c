// Scene-tile chain engine (kernels_chain.cu): one persistent kernel per decoder-layer call that
// keeps a 128-row tile of anchor queries on the SM through the whole post-conv chain of
// CustomTransformerDecoderLayer.forward (transfuser_model_v2.py:343-382).
#pragma once
#include <cuda.h>
#include <stdint.h>

#include "kernels.h"

namespace ddh {

constexpr int CH_MAX_MAPS = 14;
constexpr int CH_MAX_OPS = 24;
constexpr int CH_MAX_STEPS = 16;
constexpr int CH_MAX_PAR = 28;
constexpr int CH_LOGITS = 16;        // attention-weight logits per row the encoder program can hoist (layers x poses)
constexpr int CH_PAR_FLOATS = 7168;   // staged per-column vectors (bias, LayerNorm, FiLM, ...)
constexpr int CH_KV_LD = 520;         // bf16 elements per hoisted K|V row (512 + 8 pad: conflict-free ldmatrix)

// epilogue kinds, in the order of the layer (see the program built by ddh_api.cu)
enum ChainEpi : uint8_t {
  CE_X1 = 0,      // x1 = acc + b + q0 -> bf16 operand; (x1 + b_attn_out) back to TMEM (residual of attn_out)
  CE_ATTN,        // q = (acc + b) * scale -> bf16; agent attention (mma.sync) -> bf16 operand
  CE_LN2EGO,      // LN1 -> + ego -> LN2 -> bf16 operand
  CE_RELU,        // ReLU(acc + b) -> bf16 operand
  CE_LN_FILM,     // LN3(acc + b) * (1 + scale) + shift -> bf16 operand
  CE_RELU_LN,     // LN(ReLU(acc + b)) -> bf16 operand
  CE_SCORE,       // LN(ReLU(acc + b)) . w6 + b6 -> scores
  CE_TAIL,        // reg head outputs: + points, tanh * pi, DDIM update
  CE_Q0           // acc + b -> q0 (tiled fp32, global)
};
enum : uint8_t { CO_ACCUM = 1, CO_N64 = 2 };
enum : uint8_t { CS_WAIT_S = 1, CS_KVGO = 2, CS_SAFREE = 4 };

struct ChainOp {        // acc[128 x N] (+)= A[128 x 64 nk] . W[n0 : n0 + N, 64 k0 : 64 (k0 + nk)]^T
  uint8_t map;          // index into ChainArgs.maps
  uint8_t a_chunk;      // first 16 KiB chunk of the A operand in the operand region (0..7)
  uint8_t nk;           // k-chunks of 64
  uint8_t k0;           // first k-chunk of W
  uint16_t n0;          // first weight row
  uint16_t acc_col;     // TMEM column of the accumulator
  uint8_t flags;        // CO_*
  uint8_t pad[3];
};
struct ChainStep {
  uint8_t op0, nops;    // MMA ops issued before the commit
  uint8_t epi;          // ChainEpi run by the compute warps once the accumulators are ready
  uint8_t flags;        // CS_*
  uint8_t dst_chunk;    // operand region chunk the epilogue writes (0 or 4)
  uint8_t pad;
  uint16_t acc_col;     // TMEM column the epilogue reads
  uint16_t par[6];      // offsets (floats) into the staged parameter block
};
struct ChainParSrc { const float* src; int n; int dst; };

struct ChainArgs {
  alignas(64) CUtensorMap maps[CH_MAX_MAPS];   // weights: bf16 [N][K], box {64 k, 256 rows} (reg4: 64 rows)
  alignas(64) CUtensorMap smap;                // sampled features S [rows][256] bf16, box {64, 128}
  ChainOp ops[CH_MAX_OPS];
  ChainStep steps[CH_MAX_STEPS];
  ChainParSrc par[CH_MAX_PAR];
  int n_steps, n_par, mode;                    // mode 0: decoder-layer chain, 1: embedding + plan_anchor_encoder
  int B, A, Na, P, spt, n_tiles;               // spt: scenes per 128-row tile
  int first_step;                              // encoder mode: img = sa * norm(anchor) + sb * noise first
  float sa, sb;
  const float* anchors;
  const float* noise;
  float* img;
  float* pts;
  float* q0t;                                  // [tile][64][128] float4: q0 tiled so that a row per lane is coalesced
  // encoder mode, optional: the attention-weight logits of every decoder layer (blocks.py:110: Linear(D -> P)
  // of the layer-invariant queries) as two partial dot products per row, one per column half of the CE_Q0
  // epilogue: [tile][2][128][CH_LOGITS]; steps[].par[1] = staged weights [n][256], par[2] = n (0: off)
  float* logit_part;
  const __nv_bfloat16* kv16;                   // [B * Na][CH_KV_LD] hoisted K|V of this layer
  const float* egov;                           // [B][256] collapsed ego attention of this layer
  float* modes;
  float* scores;
  int do_ddim;
  DdimCoef dc;
  long long* dbg;
};

int chain_engine_init();   // cudaError_t as int
int chain_smem_bytes();
void launch_chain(const ChainArgs& a, cudaStream_t st);

}  // namespace ddh
