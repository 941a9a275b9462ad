// Interface of the persistent cluster decode kernel (decode_persistent.cu).
#pragma once
#include "kernels.cuh"

namespace hmocr {

// Packed weight layout (bf16, rows padded to K+8 elements so ldmatrix is bank-conflict free).
// wblob[layer][cta c][chunk]:  chunk 0..2 = rows of self_attn.in_proj {q,k,v} of head c  [32][264]
//                              chunk 3    = rows c*32.. of self_attn.out_proj            [32][264]
//                              chunk 4    = rows of multihead_attn.in_proj q of head c   [32][264]
//                              chunk 5    = rows c*32.. of multihead_attn.out_proj       [32][264]
//                              chunk 6    = rows c*64.. of linear1                       [64][264]
//                              chunk 7    = rows c*32.. of linear2                       [32][520]
// fcblob[cta c][chunk j] = rows c*(64*fc_chunks) + j*64 .. of fc_out (zero rows past V)   [64][264]
constexpr int DP_CH_ATT = 32 * 264 * 2;
constexpr int DP_CH_F1 = 64 * 264 * 2;
constexpr int DP_CH_F2 = 32 * 520 * 2;
constexpr int DP_CH_FC = 64 * 264 * 2;
constexpr int DP_LAYER_CTA_BYTES = 6 * DP_CH_ATT + DP_CH_F1 + DP_CH_F2;

// fp32 parameters of one layer, one contiguous block of DP_FP_LAYER floats
constexpr int DP_FP_BIN = 0;        // self_attn.in_proj_bias      [768]
constexpr int DP_FP_BO = 768;       // self_attn.out_proj.bias     [256]
constexpr int DP_FP_BCQ = 1024;     // multihead_attn.in_proj_bias[:256]
constexpr int DP_FP_BCO = 1280;     // multihead_attn.out_proj.bias
constexpr int DP_FP_B1 = 1536;      // linear1.bias [512]
constexpr int DP_FP_B2 = 2048;      // linear2.bias
constexpr int DP_FP_LN1G = 2304, DP_FP_LN1B = 2560, DP_FP_LN2G = 2816, DP_FP_LN2B = 3072, DP_FP_LN3G = 3328,
              DP_FP_LN3B = 3584;
constexpr int DP_FP_LAYER = 3840;

struct DecPersistParams {
  const uint8_t* wblob;
  const uint8_t* fcblob;
  const float* fparams;       // [L][DP_FP_LAYER]
  const float* fc_bias;       // [>= vocab]
  const float* emb;           // [vocab][256]
  const float* pos;           // [max_pos][256]
  __nv_bfloat16* kcache;      // [L][rows][8][tmax][32]
  __nv_bfloat16* vcache;
  const __nv_bfloat16* memk;  // [L][images][8][30][32]
  const __nv_bfloat16* memv;
  int64_t* tokens;            // [rows][ld_tok]
  float* logprob;             // [rows][max_len] or nullptr
  uint8_t* finished;          // [rows]
  DecodeState* state;
  int rows, images, beam;     // rows = images * beam; row r reads the memory of image r / beam
  int num_layers, fc_chunks, vocab;
  int tmax, max_pos, max_len, ld_tok, eos;
};

int decode_persistent_init();
int decode_persistent_launch(cudaStream_t st, const DecPersistParams& p, int t_begin, int t_end);
int repack_memkv(cudaStream_t st, const __nv_bfloat16* memkv, int images, int L, __nv_bfloat16* memk,
                 __nv_bfloat16* memv);

}  // namespace hmocr
