// Fused Swin MLP for the wide-token stages (C = 96, 192):   x <- x + fc2( GELU( fc1(xn) ) )
//
// torchvision swin_transformer.py:444 (MLP(dim, [4 dim, dim], GELU)) + :455 (x = x + stochastic_depth(mlp(norm2(x)))),
// eval mode.  As two GEMM launches the [tokens, 4C] hidden tensor is written to HBM and read back: 2 x 377 MB per
// stage-1 block at B = 256 - 1.5 GB of the encoder's 12 GB - and fc2 runs at the HBM roof moving it.  Here the hidden
// tile never leaves the SM:
//
//   per 128-token tile, the 4C hidden columns in chunks of HC = 128:
//     fc1_j : acc1[j & 1] (TMEM, 128 columns)  =  A[128 x C] . W1[j]^T                     tcgen05.mma, K = C
//     GELU  : 16 epilogue warps read acc1 (tcgen05.ld), add b1, GELU, pack to fp16 and write the chunk into shared
//             memory in the canonical K-major 128B-swizzled layout - exactly what TMA would have written - so it is
//             the A operand of
//     fc2_j : acc2 (TMEM, C columns)  +=  H_j[128 x 128] . W2[:, j]^T                      tcgen05.mma, K = 128
//   then acc2 + b2 + x -> x (fp32, in place), coalesced through a swizzled shared tile.
//
// Warp roles (576 threads, one CTA per SM, persistent over tiles): warp 0 = TMA producer (the A tile once per tile,
// the weight boxes [128 x 64] of W1 / [C x 64] of W2 through a ring, in exactly the order the MMA warp consumes them),
// warp 1 = MMA issuer (fc1 of chunk j+1 is issued BEFORE fc2 of chunk j, so the tensor pipe works while the epilogue
// warps are in the GELU of chunk j), warps 2..17 = epilogue (a warp owns TMEM lane quarter warp % 4 and one 32-column
// group of every chunk).  TMEM: acc1[2] (2 x 128 columns) + acc2 (C columns) <= 448 of 512 columns.
// The weights (2 x 72 KB at C = 96) stream from L2 per tile: 168 KB per 128 tokens against 120 KB of HBM traffic.
#include "gemm.cuh"
#include "kernels.cuh"

namespace hmocr {
namespace {

constexpr int BM = 128, BK = 64, HC = 128;
constexpr int EPI_WARPS = 16;
constexpr int NUM_THREADS = 32 * (2 + EPI_WARPS);

template <int C>
struct MlpCfg {
  static constexpr int KB_A = (C + BK - 1) / BK;               // k-blocks of the A tile and of a W1 box row (K = C)
  static constexpr int A_BYTES = KB_A * BM * BK * 2;           // 32 KB (C = 96: the second k-block is half zero fill) / 48 KB
  static constexpr int NCH = 4 * C / HC;                       // hidden chunks per tile: 3 / 6
  static constexpr int KB_H = HC / BK;                         // k-blocks of a hidden chunk: 2
  static constexpr int H_BYTES = BM * HC * 2;                  // 32 KB per hidden buffer
  static constexpr int W1_BOX = HC * BK * 2;                   // [128 hidden rows x 64 k]  16 KB
  static constexpr int W2_BOX = C * BK * 2;                    // [C output rows x 64 k]    12 / 24 KB
  static constexpr int STAGE_BYTES = W1_BOX > W2_BOX ? W1_BOX : W2_BOX;
  // The ring must hold about two chunks' worth of boxes (a chunk consumes KB_A + KB_H of them) or the L2 latency of
  // a box shows up in every chunk: 8 x 16 KB at C = 96 (two chunks).  C = 192 has room for 4 x 24 KB only (less than
  // one chunk: its launch is bound by exactly that, 133 us against 139 us for the two GEMM launches).
  static constexpr int STAGES = C == 96 ? 8 : (96 * 1024) / STAGE_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int BIAS_BYTES = 4 * C * 4;                 // b1 [4C], read by every chunk of every tile
  static constexpr int SMEM_BYTES = A_BYTES + 2 * H_BYTES + STAGES * STAGE_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
  static constexpr uint32_t ACC2_COL = 2 * HC;
  static_assert(C % 32 == 0 && C <= 256 && (4 * C) % HC == 0, "fused MLP: unsupported width");
  static_assert(ACC2_COL + C <= 512, "TMEM columns");
  static_assert(EPI_WARPS * 4096 <= 2 * H_BYTES, "the final epilogue's staging tiles alias the hidden buffers");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory");
};

struct MlpParams {
  int M, num_tiles;
  const float* b1;      // [4C]
  const float* b2;      // [C]
  float* x;             // [M, C] fp32 residual stream, updated in place
};

// same descriptors as gemm.cu: K-major, 128B swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__device__ __forceinline__ constexpr uint32_t idesc_f16(int n) {      // fp16 x fp16 -> fp32, M = 128, N = n
  return (1u << 4) | (uint32_t(n >> 3) << 17) | (uint32_t(BM >> 4) << 24);
}

template <int C>
__global__ void __launch_bounds__(NUM_THREADS, 1)
swin_mlp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const MlpParams p) {
  using G = MlpCfg<C>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sH = sA + G::A_BYTES;                          // two hidden buffers; the final epilogue's staging aliases them
  uint8_t* sW = sH + 2 * G::H_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + G::STAGES * G::STAGE_BYTES);
  uint64_t* w_full = bars;                                // [STAGES] TMA bytes of a weight box
  uint64_t* w_empty = w_full + G::STAGES;                 // [STAGES] tcgen05.commit: the box has been consumed
  uint64_t* a_full = w_empty + G::STAGES;                 // A tile landed
  uint64_t* a_empty = a_full + 1;                         // commit after the last fc1 of the tile
  uint64_t* acc1_full = a_empty + 1;                      // [2] commit after fc1_j
  uint64_t* acc1_empty = acc1_full + 2;                   // [2] 16 epilogue warps have read the chunk out of TMEM
  uint64_t* h_ready = acc1_empty + 2;                     // [2] 16 epilogue warps have written the fp16 chunk
  uint64_t* h_free = h_ready + 2;                         // [2] commit after fc2_j
  uint64_t* acc2_full = h_free + 2;                       // commit after the last fc2 of the tile
  uint64_t* acc2_empty = acc2_full + 1;                   // 16 epilogue warps are done with acc2
  uint64_t* stg_free = acc2_empty + 1;                    // 16 epilogue warps are done with their staging tiles (inside sH)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(stg_free + 1);
  float* s_b1 = reinterpret_cast<float*>(sW + G::STAGES * G::STAGE_BYTES + G::BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2);
    for (int s = 0; s < G::STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(a_full, 1);
    mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc1_full[b], 1);
      mbar_init(&acc1_empty[b], EPI_WARPS);
      mbar_init(&h_ready[b], EPI_WARPS);
      mbar_init(&h_free[b], 1);
    }
    mbar_init(acc2_full, 1);
    mbar_init(acc2_empty, EPI_WARPS);
    mbar_init(stg_free, EPI_WARPS);
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < 4 * C; i += NUM_THREADS) s_b1[i] = __ldg(p.b1 + i);       // weights: not PDL-dependent
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                             // xn and x belong to earlier kernels up to here

  if (warp == 0) {
    // ---- TMA producer -------------------------------------------------------------------------------
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      auto push_w1 = [&](int j) {                          // KB_A boxes [HC x 64] of W1 rows [j*HC, +HC)
        for (int kb = 0; kb < G::KB_A; ++kb) {
          mbar_wait(&w_empty[s], ph ^ 1u);
          mbar_expect_tx(&w_full[s], G::W1_BOX);
          tma_load_2d(sW + s * G::STAGE_BYTES, &tmW1, &w_full[s], kb * BK, j * HC);
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
      };
      auto push_w2 = [&](int j) {                          // KB_H boxes [C x 64] of W2 columns [j*HC, +HC)
        for (int kb = 0; kb < G::KB_H; ++kb) {
          mbar_wait(&w_empty[s], ph ^ 1u);
          mbar_expect_tx(&w_full[s], G::W2_BOX);
          tma_load_2d(sW + s * G::STAGE_BYTES, &tmW2, &w_full[s], j * HC + kb * BK, 0);
          if (++s == G::STAGES) { s = 0; ph ^= 1u; }
        }
      };
      uint32_t it = 0;
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        {   // the tile's fp32 residual rows (contiguous) -> L2, so the final epilogue's loads are L2 hits
          const int rows = p.M - t * BM < BM ? p.M - t * BM : BM;
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(p.x + (size_t)t * BM * C)),
                       "r"(rows * C * 4)
                       : "memory");
        }
        mbar_wait(a_empty, (it & 1u) ^ 1u);
        mbar_expect_tx(a_full, G::A_BYTES);
        for (int kb = 0; kb < G::KB_A; ++kb) tma_load_2d(sA + kb * (BM * BK * 2), &tmA, a_full, kb * BK, t * BM);
        for (int j = 0; j <= G::NCH; ++j) {                // the MMA warp's order: fc1_j, then fc2_{j-1}
          if (j < G::NCH) push_w1(j);
          if (j >= 1) push_w2(j - 1);
        }
      }
    }
  } else if (warp == 1) {
    // ---- MMA issuer ---------------------------------------------------------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc1 = idesc_f16(HC), idesc2 = idesc_f16(C);
      int s = 0;
      uint32_t ph = 0;
      uint32_t n_acc1[2] = {0u, 0u}, n_h[2] = {0u, 0u};
      uint32_t it = 0;
      const uint32_t a_base = smem_u32(sA), h_base = smem_u32(sH);
      for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
        mbar_wait_backoff(a_full, it & 1u);
        for (int j = 0; j <= G::NCH; ++j) {
          if (j < G::NCH) {                                // fc1_j -> acc1[j & 1]
            const int b = j & 1;
            mbar_wait_backoff(&acc1_empty[b], (n_acc1[b] & 1u) ^ 1u);
            ++n_acc1[b];
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + b * HC;
            for (int kb = 0; kb < G::KB_A; ++kb) {
              mbar_wait(&w_full[s], ph);
              tc_fence_after();
              const uint32_t wb = smem_u32(sW + s * G::STAGE_BYTES);
              const int nks = (C - kb * BK) >= BK ? BK / 16 : (C - kb * BK) / 16;
              for (int k = 0; k < nks; ++k)
                umma_f16(d_tmem, smem_desc(a_base + kb * (BM * BK * 2) + k * 32), smem_desc(wb + k * 32), idesc1,
                         (kb | k) != 0 ? 1u : 0u);
              umma_commit(&w_empty[s]);
              if (++s == G::STAGES) { s = 0; ph ^= 1u; }
            }
            umma_commit(&acc1_full[b]);
            if (j == G::NCH - 1) umma_commit(a_empty);      // the A tile may be overwritten by the next tile's
          }
          if (j >= 1) {                                    // fc2_{j-1}: acc2 += H[(j-1) & 1] . W2[:, j-1]^T
            const int jj = j - 1, b = jj & 1;
            if (jj == 0) mbar_wait_backoff(acc2_empty, (it & 1u) ^ 1u);
            mbar_wait_backoff(&h_ready[b], n_h[b] & 1u);
            ++n_h[b];
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + G::ACC2_COL;
            for (int kb = 0; kb < G::KB_H; ++kb) {
              mbar_wait(&w_full[s], ph);
              tc_fence_after();
              const uint32_t wb = smem_u32(sW + s * G::STAGE_BYTES);
              const uint32_t hb = h_base + b * G::H_BYTES + kb * (BM * BK * 2);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)
                umma_f16(d_tmem, smem_desc(hb + k * 32), smem_desc(wb + k * 32), idesc2, (jj | kb | k) != 0 ? 1u : 0u);
              umma_commit(&w_empty[s]);
              if (++s == G::STAGES) { s = 0; ph ^= 1u; }
            }
            umma_commit(&h_free[b]);
            if (jj == G::NCH - 1) umma_commit(acc2_full);
          }
        }
      }
    }
  } else {
    // ---- epilogue warps -----------------------------------------------------------------------------
    const int ew = warp - 2;
    const int quarter = warp & 3;                          // TMEM lanes [32 quarter, +32) belong to this warp
    const int cgrp = ew >> 2;                              // its 32-column group of every chunk
    const uint32_t lane_base = tmem_base + (uint32_t(quarter * 32) << 16);
    const int row_in_tile = quarter * 32 + lane;
    uint32_t n_acc1[2] = {0u, 0u}, n_h[2] = {0u, 0u};
    uint32_t it = 0;
    // hidden chunk -> shared memory: k-block (cgrp / 2) of the buffer, 16-byte pieces 4 (cgrp % 2) .. +3 of this row
    const uint32_t h_row = smem_u32(sH) + (cgrp >> 1) * (BM * BK * 2) + row_in_tile * 128;
    const uint32_t h_sw = row_in_tile & 7, h_c0 = (cgrp & 1) * 4;
    // final epilogue (second layout): lane = (row % 4, 4-column group); staging tile of this warp inside sH
    const uint32_t st_addr = smem_u32(sH) + ew * 4096;
    const int rr = lane >> 3, cg4 = lane & 7;
    for (int t = blockIdx.x; t < p.num_tiles; t += gridDim.x, ++it) {
      const int m0 = t * BM;
      for (int j = 0; j < G::NCH; ++j) {
        const int b = j & 1;
        mbar_wait_backoff(&acc1_full[b], n_acc1[b] & 1u);
        ++n_acc1[b];
        tc_fence_after();
        uint32_t r[32];
        tmem_ld32(lane_base + b * HC + cgrp * 32, r);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc1_empty[b]);        // the MMA warp may start fc1_{j+2} into these columns
        const float4* bp = reinterpret_cast<const float4*>(s_b1 + j * HC + cgrp * 32);     // warp-uniform: broadcast
        uint32_t h[16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = bp[q];
          float v0, v1, v2, v3;
          upk2(add2(pk2(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1])), pk2(bb.x, bb.y)), v0, v1);
          upk2(add2(pk2(__uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])), pk2(bb.z, bb.w)), v2, v3);
          gelu_erf2(v0, v1);
          gelu_erf2(v2, v3);
          h[2 * q] = pack16(v0, v1);
          h[2 * q + 1] = pack16(v2, v3);
        }
        mbar_wait_backoff(&h_free[b], (n_h[b] & 1u) ^ 1u);  // fc2 of the chunk that used this buffer last has completed
        ++n_h[b];
        if (j < 2) mbar_wait_backoff(stg_free, (it & 1u) ^ 1u);   // ... and the previous tile's staging tiles (same memory) are idle
        const uint32_t dst = h_row + b * G::H_BYTES;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (((h_c0 + q) ^ h_sw) << 4)), "r"(h[4 * q]),
                       "r"(h[4 * q + 1]), "r"(h[4 * q + 2]), "r"(h[4 * q + 3])
                       : "memory");
        fence_proxy_async();                               // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) mbar_arrive(&h_ready[b]);
      }
      // ---- x <- acc2 + b2 + x ---------------------------------------------------------------------------
      const int mb = m0 + quarter * 32;
      const int rows_left = p.M - mb - rr;                 // row i4*4 + rr exists iff i4*4 < rows_left
      bool first = true;
      for (int c = cgrp; c < C / 32; c += 4) {
        const int col = c * 32 + cg4 * 4;
        float* xp = p.x + (size_t)(mb + rr) * C + col;
        float4 res[8];
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4)                      // issued BEFORE the accumulator wait: L2 latency under the last fc2
          res[i4] = (i4 * 4 < rows_left) ? __ldcs(reinterpret_cast<const float4*>(xp + (size_t)i4 * 4 * C))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 bias = __ldg(reinterpret_cast<const float4*>(p.b2 + col));
        if (first) {
          mbar_wait_backoff(acc2_full, it & 1u);
          tc_fence_after();
          first = false;
        }
        uint32_t r[32];
        tmem_ld32(lane_base + G::ACC2_COL + c * 32, r);
#pragma unroll
        for (int q = 0; q < 8; ++q)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(st_addr + lane * 128 + ((q ^ (lane & 7)) << 4)),
                       "r"(r[4 * q]), "r"(r[4 * q + 1]), "r"(r[4 * q + 2]), "r"(r[4 * q + 3])
                       : "memory");
        __syncwarp();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          float4 v;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                       : "r"(st_addr + rr * 128 + i4 * 512 + ((cg4 ^ ((i4 * 4 + rr) & 7)) << 4)));
          v.x += bias.x + res[i4].x; v.y += bias.y + res[i4].y; v.z += bias.z + res[i4].z; v.w += bias.w + res[i4].w;
          if (i4 * 4 < rows_left) *reinterpret_cast<float4*>(xp + (size_t)i4 * 4 * C) = v;
        }
        __syncwarp();                                      // the staging tile is reused by this warp's next chunk
      }
      if (first) {                                         // a warp without a column chunk still follows the phases
        mbar_wait_backoff(acc2_full, it & 1u);
        tc_fence_after();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(acc2_empty);
        mbar_arrive(stg_free);     // the next tile's first two hidden chunks land where the staging tiles are
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int C>
int launch_mlp(cudaStream_t st, const h16* xn, int M, const h16* w1, const float* b1, const h16* w2, const float* b2,
               float* x) {
  using G = MlpCfg<C>;
  HM_DEVICE_ONCE(HM_CUDA(cudaFuncSetAttribute(swin_mlp_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM_BYTES)));
  alignas(64) CUtensorMap tmA, tmW1, tmW2;
  HM_TRY(gemm_tensor_map(xn, M, C, C, BM, &tmA));
  HM_TRY(gemm_tensor_map(w1, 4 * C, C, C, HC, &tmW1));
  HM_TRY(gemm_tensor_map(w2, C, 4 * C, 4 * C, C, &tmW2));
  MlpParams p;
  p.M = M;
  p.num_tiles = ceil_div(M, BM);
  p.b1 = b1; p.b2 = b2; p.x = x;
  const int sms = gemm_num_sms();
  const int grid = p.num_tiles < sms ? p.num_tiles : sms;
  HM_CUDA(launch_pdl(swin_mlp_kernel<C>, dim3(grid), dim3(NUM_THREADS), G::SMEM_BYTES, st, tmA, tmW1, tmW2, p));
  HM_LAUNCHED();
  return 0;
}

}  // namespace

bool swin_mlp_supported(int C) { return C == 96 || C == 192; }

int swin_mlp(cudaStream_t st, const h16* xn, int M, int C, const h16* w1, const float* b1, const h16* w2,
             const float* b2, float* x) {
  HM_TRY(gemm_init());
  HM_CHECK(M > 0, "swin_mlp: empty input");
  HM_CHECK(b1 != nullptr && b2 != nullptr, "swin_mlp: biases missing");
  HM_CHECK((reinterpret_cast<uintptr_t>(xn) & 15) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "swin_mlp: unaligned operands");
  switch (C) {
    case 96: return launch_mlp<96>(st, xn, M, w1, b1, w2, b2, x);
    case 192: return launch_mlp<192>(st, xn, M, w1, b1, w2, b2, x);
    default: break;
  }
  HM_CHECK(false, "swin_mlp: C=%d unsupported (96, 192)", C);
  return -2;
}

}  // namespace hmocr
