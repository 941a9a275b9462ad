// Interface of the persistent cluster decode kernel (decode_persistent.cu).
#pragma once
#include "kernels.cuh"

namespace hmocr {

// Packed weight layout (bf16, rows padded to K+8 elements so ldmatrix is bank-conflict free).
// Every chunk is one cp.async.bulk copy of <= DP_CHUNK bytes.
// wblob[layer][cta c][chunk]:
//   0..2  rows of self_attn.in_proj {q,k,v} of head c                 [32][264]
//   3     rows c*32.. of self_attn.out_proj                           [32][264]
//   4     rows of multihead_attn.in_proj q of head c                  [32][264]
//   5     rows c*32.. of multihead_attn.out_proj                      [32][264]
//   6,7   rows c*64.. / c*64+32.. of linear1                          [32][264]
//   8,9   rows c*32.. / c*32+16.. of linear2                          [16][520]
// fcblob[cta c][chunk j] = rows c*(32*fc_chunks) + j*32 .. of fc_out (zero rows past V)  [32][264]
constexpr int DP_CHUNK = 32 * 264 * 2;        // 16896 B
constexpr int DP_CHUNK_F2 = 16 * 520 * 2;     // 16640 B
constexpr int DP_LAYER_CHUNKS = 10;
constexpr int DP_LAYER_CTA_BYTES = 8 * DP_CHUNK + 2 * DP_CHUNK_F2;

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

// The kernel reads a per-(layer, CTA) repack of those: only the slices CTA c needs, DP_FPC floats,
// fetched into shared memory with one bulk copy per layer.
constexpr int DPC_BQKV = 0;     // in_proj_bias q|k|v of head c            [3][32]
constexpr int DPC_BO = 96;      // self out_proj.bias[c*32 ..]             [32]
constexpr int DPC_BCQ = 128;    // cross in_proj_bias q of head c          [32]
constexpr int DPC_BCO = 160;    // cross out_proj.bias[c*32 ..]            [32]
constexpr int DPC_B1 = 192;     // linear1.bias[c*64 ..]                   [64]
constexpr int DPC_B2 = 256;     // linear2.bias[c*32 ..]                   [32]
constexpr int DP_FPC = 288;
constexpr int DP_FCB_MAX = 1024;    // max vocabulary columns per CTA (vocab <= 8192)

struct DecPersistParams {
  const uint8_t* wblob;
  const uint8_t* fcblob;
  const float* fparams;       // [L][8][DP_FPC] per-(layer, CTA) bias slices
  const float* lnparams;      // [L][6][256]: norm1.weight, norm1.bias, norm2.weight, ... norm3.bias
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
  int rows_per_cluster;       // sequences owned by one cluster (1..16); <= 0 = automatic
  int num_layers, fc_chunks, vocab;
  int tmax, max_pos, max_len, ld_tok, eos;
  long long* trace;           // optional: clock64() of cluster 0 / CTA 0 / thread 0 at every phase boundary
  int trace_step;             //           of decode step `trace_step`
};

int decode_persistent_init();
// rows_per_cluster <= 0: chosen so that all clusters are co-resident (single wave) when possible
int decode_persistent_launch(cudaStream_t st, DecPersistParams p, int t_begin, int t_end);
int decode_persistent_max_clusters(int* out);   // co-resident clusters on this device
int repack_memkv(cudaStream_t st, const __nv_bfloat16* memkv, int images, int L, __nv_bfloat16* memk,
                 __nv_bfloat16* memv);

}  // namespace hmocr
