// Non-GEMM kernels of the Swin-T encoder (sm_100a): patch embedding + LN, LayerNorm,
// PatchMerging gather + LN, fused shifted-window attention.
#include "kernels.cuh"

namespace hmocr {
namespace {

constexpr float LN_EPS = 1e-5f;

// ------------------------------------------------------------------------------------------
// LayerNorm (one warp per row, row held in registers, two-pass variance)
// swin_transformer.py:433,443 (norm1/norm2), :74 (PatchMerging.norm); torch LayerNorm eps 1e-5
// ------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 12;   // float4 per lane: C <= 1536

struct RowSrc {
  const float* x;
  int C;
  __device__ __forceinline__ const float4* vec(int row, int idx) const {
    return reinterpret_cast<const float4*>(x + (size_t)row * C) + idx;
  }
};

// PatchMerging: output row (b,i,j) = concat of x[b,2i+dr,2j+dc,:] for (dr,dc) in (0,0),(1,0),(0,1),(1,1)
struct MergeSrc {
  const float* x;
  int H, W, Cin;   // input grid and channels; output C = 4*Cin
  __device__ __forceinline__ const float4* vec(int row, int idx) const {
    const int Ho = H >> 1, Wo = W >> 1;
    const int j = row % Wo, i = (row / Wo) % Ho, b = row / (Wo * Ho);
    const int per = Cin >> 2;
    const int seg = idx / per, within = idx - seg * per;
    const int r = 2 * i + (seg & 1), c = 2 * j + (seg >> 1);
    return reinterpret_cast<const float4*>(x + ((size_t)(b * H + r) * W + c) * Cin) + within;
  }
};

template <class Src>
__global__ void __launch_bounds__(256) layernorm_kernel(Src src, int rows, int C, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        h16* __restrict__ out16, float* __restrict__ out32) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = C >> 2;
  float4 v[LN_MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      v[i] = *src.vec(row, idx);
      sum += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  }
  const float mean = warp_sum(sum) / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += a * a + b * b + c * c + d * d;
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / C + LN_EPS);
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int idx = lane + 32 * i;
    if (idx < nvec) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
      const float4 b = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      float4 y;
      y.x = (v[i].x - mean) * rstd * g.x + b.x;
      y.y = (v[i].y - mean) * rstd * g.y + b.y;
      y.z = (v[i].z - mean) * rstd * g.z + b.z;
      y.w = (v[i].w - mean) * rstd * g.w + b.w;
      if (out16 != nullptr)
        reinterpret_cast<uint2*>(out16 + (size_t)row * C)[idx] = make_uint2(pack16(y.x, y.y), pack16(y.z, y.w));
      if (out32 != nullptr) reinterpret_cast<float4*>(out32 + (size_t)row * C)[idx] = y;
    }
  }
}

// Specialised LayerNorm for the Swin channel counts: a row is shared by G = C / (4 V) lanes (V float4 per
// lane), so a warp normalises 32 / G rows at once and every lane moves data (the generic kernel above
// keeps 8 of 32 lanes idle at C = 96 and spends most of its instructions on predicated-off iterations).
// Warps walk the rows grid-stride; gamma / beta stay in registers when they fit.
template <int C, int V, bool MERGE>
__global__ void __launch_bounds__(256) layernorm_c_kernel(const float* __restrict__ x, int rows, int H, int W,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          h16* __restrict__ out16, float* __restrict__ out32) {
  constexpr int G = C / (4 * V), RPW = 32 / G;
  static_assert(G >= 1 && G <= 32 && (G & (G - 1)) == 0 && G * 4 * V == C, "bad LayerNorm geometry");
  constexpr bool PARAMS_IN_REGS = V <= 3;
  const int lane = threadIdx.x & 31, g = lane % G, sub = lane / G;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  pdl_launch_dependents();
  float4 gr[PARAMS_IN_REGS ? V : 1], br[PARAMS_IN_REGS ? V : 1];
  if (PARAMS_IN_REGS) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      gr[i] = __ldg(reinterpret_cast<const float4*>(gamma) + g + G * i);
      br[i] = __ldg(reinterpret_cast<const float4*>(beta) + g + G * i);
    }
  }
  pdl_wait();
  for (int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW; row0 < rows; row0 += warps_total * RPW) {
    const int row = row0 + sub;
    const bool ok = row < rows;
    float4 v[V];
    float sum = 0.f;
    size_t base00 = 0;
    if (MERGE) {   // output row (b,i,j) = concat of x[b,2i+dr,2j+dc,:] for (dr,dc) in (0,0),(1,0),(0,1),(1,1)
      const int Ho = H >> 1, Wo = W >> 1, rr = ok ? row : 0;
      const int j = rr % Wo, i = (rr / Wo) % Ho, b = rr / (Wo * Ho);
      base00 = ((size_t)(b * H + 2 * i) * W + 2 * j) * (C / 4);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int idx = g + G * i;                       // float4 index within the row
      const float4* p;
      if (MERGE) {
        constexpr int per = C / 16;                    // float4 per source segment
        const int sgm = idx / per;
        p = reinterpret_cast<const float4*>(x + base00 + (size_t)((sgm & 1) * W + (sgm >> 1)) * (C / 4)) + (idx - sgm * per);
      } else {
        p = reinterpret_cast<const float4*>(x + (size_t)row * C) + idx;
      }
      v[i] = ok ? __ldcs(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * (1.0f / C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
      sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
    }
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * (1.0f / C) + LN_EPS);
    if (!ok) continue;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int idx = g + G * i;
      float4 gg, bb;
      if (PARAMS_IN_REGS) { gg = gr[i]; bb = br[i]; }
      else {
        gg = __ldg(reinterpret_cast<const float4*>(gamma) + idx);
        bb = __ldg(reinterpret_cast<const float4*>(beta) + idx);
      }
      float4 y;
      y.x = v[i].x * rstd * gg.x + bb.x; y.y = v[i].y * rstd * gg.y + bb.y;
      y.z = v[i].z * rstd * gg.z + bb.z; y.w = v[i].w * rstd * gg.w + bb.w;
      if (out16 != nullptr)
        reinterpret_cast<uint2*>(out16 + (size_t)row * C)[idx] = make_uint2(pack16(y.x, y.y), pack16(y.z, y.w));
      if (out32 != nullptr) reinterpret_cast<float4*>(out32 + (size_t)row * C)[idx] = y;
    }
  }
}

template <int C, int V, bool MERGE>
int launch_ln_c(cudaStream_t st, const float* x, int rows, int H, int W, const float* gamma, const float* beta,
                h16* out16, float* out32) {
  constexpr int RPW = 32 / (C / (4 * V));
  int blocks = ceil_div(rows, 8 * RPW);
  if (blocks > 148 * 8) blocks = 148 * 8;
  HM_CUDA(launch_pdl(layernorm_c_kernel<C, V, MERGE>, dim3(blocks), dim3(256), 0, st, x, rows, H, W, gamma, beta, out16, out32));
  HM_LAUNCHED();
  return 0;
}

// ------------------------------------------------------------------------------------------
// Patch embedding: Conv2d(1,96,k4,s4) + NHWC + LayerNorm(96)     swin_transformer.py:556-562
// One THREAD per token: its 16 pixels in registers, 96 accumulators as 48 fp32 pairs (FFMA2), the
// transposed conv weight [tap][channel] broadcast from shared memory; the LayerNorm statistics need no
// shuffles.  A warp's 32 tokens are one contiguous 12 KB block of the output, so the normalised rows go
// through a padded per-warp shared tile and leave as 24 fully coalesced 512-byte stores.
// ------------------------------------------------------------------------------------------
constexpr int PE_WARPS = 4, PE_PITCH = 100;
struct PatchSmem {
  float wT[16][96];
  float bias[96], g[96], beta[96];
  float tile[PE_WARPS][32][PE_PITCH];
};

__global__ void __launch_bounds__(PE_WARPS * 32, 3) patch_embed_kernel(const float* __restrict__ img, int ntok,
                                                                      const float* __restrict__ w,
                                                                      const float* __restrict__ bias,
                                                                      const float* __restrict__ g,
                                                                      const float* __restrict__ beta,
                                                                      float* __restrict__ x) {
  extern __shared__ __align__(16) uint8_t pe_raw[];
  PatchSmem& s = *reinterpret_cast<PatchSmem*>(pe_raw);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_launch_dependents();
  for (int i = threadIdx.x; i < 96 * 16; i += PE_WARPS * 32) s.wT[i & 15][i >> 4] = __ldg(w + i);   // w[ch][tap]
  for (int i = threadIdx.x; i < 96; i += PE_WARPS * 32) {
    s.bias[i] = __ldg(bias + i);
    s.g[i] = __ldg(g + i);
    s.beta[i] = __ldg(beta + i);
  }
  __syncthreads();
  pdl_wait();
  const uint32_t w_addr = smem_u32(&s.wT[0][0]), b_addr = smem_u32(&s.bias[0]);
  float(*tile)[PE_PITCH] = s.tile[warp];
  const int groups = ntok >> 5;                                   // ntok = B * 1920 is a multiple of 32
  for (int grp = blockIdx.x * PE_WARPS + warp; grp < groups; grp += gridDim.x * PE_WARPS) {
    const int tok = grp * 32 + lane;
    const int b = tok / 1920, rem = tok - b * 1920, ty = rem / 80, tx = rem - ty * 80;
    const float* base = img + ((size_t)b * 96 + ty * 4) * 320 + tx * 4;
    float px[16];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(base + r * 320));
      px[4 * r] = v.x; px[4 * r + 1] = v.y; px[4 * r + 2] = v.z; px[4 * r + 3] = v.w;
    }
    uint64_t acc[48];
#pragma unroll
    for (int c = 0; c < 24; ++c)
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(acc[2 * c]), "=l"(acc[2 * c + 1]) : "r"(b_addr + c * 16));
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint64_t pk = pk2(px[k], px[k]);
#pragma unroll
      for (int c = 0; c < 24; ++c) {
        uint64_t w0, w1;
        asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(w_addr + (k * 96 + c * 4) * 4));
        acc[2 * c] = fma2(pk, w0, acc[2 * c]);
        acc[2 * c + 1] = fma2(pk, w1, acc[2 * c + 1]);
      }
    }
    float o[96];
#pragma unroll
    for (int c = 0; c < 48; ++c) upk2(acc[c], o[2 * c], o[2 * c + 1]);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 96; ++c) sum += o[c];
    const float mean = sum * (1.0f / 96.0f);
    float sq = 0.f;
#pragma unroll
    for (int c = 0; c < 96; ++c) { o[c] -= mean; sq = fmaf(o[c], o[c], sq); }
    const float rstd = rsqrtf(sq * (1.0f / 96.0f) + LN_EPS);
    __syncwarp();                                                   // the previous group has left the tile
#pragma unroll
    for (int c = 0; c < 24; ++c) {
      const float4 gg = *reinterpret_cast<const float4*>(&s.g[4 * c]);
      const float4 bb = *reinterpret_cast<const float4*>(&s.beta[4 * c]);
      *reinterpret_cast<float4*>(&tile[lane][4 * c]) =
          make_float4(fmaf(o[4 * c] * rstd, gg.x, bb.x), fmaf(o[4 * c + 1] * rstd, gg.y, bb.y),
                      fmaf(o[4 * c + 2] * rstd, gg.z, bb.z), fmaf(o[4 * c + 3] * rstd, gg.w, bb.w));
    }
    __syncwarp();
    float* xo = x + (size_t)grp * 32 * 96;
#pragma unroll
    for (int it = 0; it < 24; ++it) {
      const int f = it * 128 + lane * 4, tk = f / 96, ch = f - tk * 96;
      *reinterpret_cast<float4*>(xo + f) = *reinterpret_cast<const float4*>(&tile[tk][ch]);
    }
  }
}

}  // namespace

int layernorm(cudaStream_t st, const float* x, int rows, int C, const float* gamma, const float* beta,
              h16* out16, float* out32) {
  HM_CHECK(C % 4 == 0 && C <= LN_MAXV * 128, "layernorm: C=%d unsupported (multiple of 4, <= %d)", C, LN_MAXV * 128);
  HM_CHECK(rows > 0, "layernorm: empty input");
  switch (C) {      // the Swin-T widths get the lane-exact kernel
    case 96: return launch_ln_c<96, 3, false>(st, x, rows, 0, 0, gamma, beta, out16, out32);
    case 192: return launch_ln_c<192, 3, false>(st, x, rows, 0, 0, gamma, beta, out16, out32);
    case 384: return launch_ln_c<384, 3, false>(st, x, rows, 0, 0, gamma, beta, out16, out32);
    case 768: return launch_ln_c<768, 6, false>(st, x, rows, 0, 0, gamma, beta, out16, out32);
    default: break;
  }
  RowSrc src{x, C};
  layernorm_kernel<RowSrc><<<ceil_div(rows, 8), 256, 0, st>>>(src, rows, C, gamma, beta, out16, out32);
  HM_LAUNCHED();
  return 0;
}

int patch_merge_ln(cudaStream_t st, const float* x, int B, int H, int W, int Cin, const float* gamma,
                   const float* beta, h16* out16) {
  HM_CHECK(H % 2 == 0 && W % 2 == 0, "patch_merge: odd grid %dx%d (the reference pads; never hit at 96x320)", H, W);
  HM_CHECK(Cin % 4 == 0 && 4 * Cin <= LN_MAXV * 128, "patch_merge: C=%d unsupported", Cin);
  const int rows = B * (H / 2) * (W / 2);
  switch (Cin) {
    case 96: return launch_ln_c<384, 3, true>(st, x, rows, H, W, gamma, beta, out16, nullptr);
    case 192: return launch_ln_c<768, 6, true>(st, x, rows, H, W, gamma, beta, out16, nullptr);
    case 384: return launch_ln_c<1536, 12, true>(st, x, rows, H, W, gamma, beta, out16, nullptr);
    default: break;
  }
  MergeSrc src{x, H, W, Cin};
  layernorm_kernel<MergeSrc><<<ceil_div(rows, 8), 256, 0, st>>>(src, rows, 4 * Cin, gamma, beta, out16, nullptr);
  HM_LAUNCHED();
  return 0;
}

int patch_embed(cudaStream_t st, const float* images, int B, const float* w, const float* b, const float* g,
                const float* beta, float* x) {
  const int ntok = B * 24 * 80;
  HM_DEVICE_ONCE(HM_CUDA(cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PatchSmem))));
  int blocks = ceil_div(ntok / 32, PE_WARPS);
  if (blocks > 148 * 3) blocks = 148 * 3;
  HM_CUDA(launch_pdl(patch_embed_kernel, dim3(blocks), dim3(PE_WARPS * 32), sizeof(PatchSmem), st, images, ntok, w, b, g, beta, x));
  HM_LAUNCHED();
  return 0;
}

}  // namespace hmocr
