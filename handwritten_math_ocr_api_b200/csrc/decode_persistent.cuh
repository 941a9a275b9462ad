// Interface of the persistent cluster decode kernel (decode_persistent.cu).
#pragma once
#include <cuda_fp16.h>

#include "kernels.cuh"

namespace hmocr {

// Packed weight stream (fp16, values saturated to +-65504).  Every warp of every CTA has its OWN stream of "pieces":
// a piece = 16 weight rows (output features, one mma m-tile) x a K-slice of the input columns, rows padded by 8 halves so
// ldmatrix is bank-conflict free, and it is exactly what ONE warp consumes in ONE projection phase.  The fp32 bias of a
// row travels in that padding (the first two pad halves) of the piece that holds the row's FIRST K-slice; the other
// slices carry zero.  Every piece is one cp.async.bulk copy; pieces sit at a fixed stride of DP_PIECE bytes.
//
// wstream[cta c][warp w][piece]: per layer, in consumption order
//    warps 0..5:  in_proj rows of head c - tile j = w: part = j / 2 (q, k, v), rows [c*32 + 16 (j % 2) ..]:
//                 two K = 128 pieces (columns 0..127, 128..255) accumulated by the same warp
//    all warps:   self_attn.out_proj   rows [c*32 + 16 (w / 4) ..], columns [64 (w % 4) ..]      K = 64 piece
//                 multihead_attn q     rows [c*32 + 16 (w / 4) ..], columns [64 (w % 4) ..]      K = 64 piece
//                 multihead_attn.out_proj                      (same split)                     K = 64 piece
//                 linear1              rows [c*64 + 16 (w / 2) ..], columns [128 (w % 2) ..]     K = 128 piece
//                 linear2              rows [c*32 + 16 (w / 4) ..], columns [128 (w % 4) ..]     K = 128 piece (of 512)
// then per step fc_tiles / 8 tiles of fc_out: rows [c*16*fc_tiles + 16 (w + 8 i) ..], two K = 128 pieces each.
// So ALL 8 warps work in every projection phase; the partial tiles of a K-split are summed through shared memory.
constexpr int DP_CH_ROWS = 16;
constexpr int DP_PIECE = DP_CH_ROWS * 136 * 2;     // 4352 B: 16 rows x 128 columns (+ 8 pad)
constexpr int DP_PIECE_Q = DP_CH_ROWS * 72 * 2;    // 2304 B: 16 rows x 64 columns (+ 8 pad)
constexpr int DP_QKV_WARPS = 6;                    // warps that own an in_proj tile
// pieces per layer of a warp, and per step
constexpr int dp_layer_pieces(int warp) { return warp < DP_QKV_WARPS ? 7 : 5; }
constexpr int dp_step_pieces(int warp, int layers, int fc_tiles) { return dp_layer_pieces(warp) * layers + 2 * (fc_tiles / 8); }
// bytes between the streams of two warps / two CTAs
inline size_t dp_warp_stride(int layers, int fc_tiles) { return (size_t)dp_step_pieces(0, layers, fc_tiles) * DP_PIECE; }
inline size_t dp_cta_stride(int layers, int fc_tiles) { return 8 * dp_warp_stride(layers, fc_tiles); }

constexpr int DP_ROWS = 8;          // sequences owned by one cluster (= the N of mma.m16n8k16)

struct DecPersistParams {
  const uint8_t* wstream;     // [8 CTAs][8 warps][pieces per step][DP_PIECE] (layout above)
  const float* lnparams;      // [L][6][256]: norm1.weight, norm1.bias, norm2.weight, ... norm3.bias
  const float* emb;           // [vocab][256]
  const float* pos;           // [max_pos][256]
  // fp16 caches in mma-fragment-major blocks of 32 keys x 32 dims (2048 bytes, layout in decode_persistent.cu)
  __half* kcache;             // [L][rows][8][cache_blocks][1024]
  __half* vcache;             // [L][rows][8][cache_blocks][1024]
  const __half* memk;         // [L][images][8][1024]   (mem_len <= 32 memory tokens, the other slots zero)
  const __half* memv;         // [L][images][8][1024]
  int cache_blocks;           // ceil(max_seq_len / 32)
  int mem_len;                // memory tokens per image: 30 (Swin-T) or 10 (ResNet-18 variant)
  int64_t* tokens;            // [rows][ld_tok]
  float* logprob;             // [rows][max_len] or nullptr
  uint8_t* finished;          // [rows]
  DecodeState* state;
  int rows, images, beam;     // rows = images * beam; row r reads the memory of image r / beam
  // ---- beam search (beam > 1; the reference has none - definition in oracle/decode.py::beam_search) ----
  // A cluster owns floor(8 / beam) images = rows_per_cluster rows.  The K/V caches exist twice: step t reads the
  // history of hypothesis j from set (t & 1), row bm_src[j] (its parent's row), and writes history + this step's
  // key/value to set ((t + 1) & 1), row j - the parent gather rides on the attention loads, nothing is copied
  // separately.  Per step the chosen (parent, token) of every row is recorded for the final back-track.
  int rows_per_cluster;       // 8 (greedy) or beam * (8 / beam)
  int num_clusters;
  size_t cache_set_stride;    // halves between the two cache sets (0 in greedy mode)
  float* bm_score;            // [rows] summed log-probability of each live hypothesis
  int* bm_fin;                // [rows] hypothesis has emitted eos
  int* bm_src;                // [rows] cluster-local row that holds the hypothesis' history in the current set
  int* bm_tok;                // [rows] token fed at the next step
  int* bp_parent;             // [max_len][rows] beam index (within the image) of the parent chosen at step t
  int* bp_token;              // [max_len][rows] token chosen at step t
  int num_layers, fc_tiles, vocab;
  int tmax, max_pos, max_len, ld_tok, eos, pad;
  long long* trace;           // optional: clock64() of cluster 0 / CTA 0 / thread 0 at every phase boundary
  int trace_step;             //           of decode step `trace_step`
  int flags;                  // developer switches: 1 = no L2 prefetch of the next layer's cache, 2 = none of the memory K/V
};

int decode_persistent_init();
int decode_persistent_launch(cudaStream_t st, DecPersistParams p, int t_begin, int t_end);
int decode_persistent_max_clusters(int* out);   // co-resident clusters on this device
// memkv f32 [img*mem_len+s][l*512 + kv*256 + h*32 + d] -> the fp16 memk / memv layouts above
int repack_memkv(cudaStream_t st, const float* memkv, int images, int L, int mem_len, void* memk, void* memv);
constexpr int DP_MAX_BEAM = 5;
// beam search bookkeeping around the persistent kernel
int beam_init(cudaStream_t st, DecodeState* state, float* bm_score, int* bm_fin, int* bm_src, int* bm_tok, int rows,
              int beam, int rows_per_cluster, int sos);
// best hypothesis per image (highest score, ties -> lowest beam index), back-tracked through bp_parent / bp_token
int beam_finalize(cudaStream_t st, const DecodeState* state, const float* bm_score, const int* bp_parent,
                  const int* bp_token, int images, int beam, int rows, int max_len, int sos, int pad, int64_t* tokens,
                  int ld_tok, float* score_out, int32_t* steps_out);

}  // namespace hmocr
