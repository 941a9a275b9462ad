// Non-GEMM kernels of the ResNet-18 + TransformerEncoder encoder (BASELINE.json config 4,
// /root/reference/src/model_res18trans.py:13-64; torchvision resnet.py BasicBlock).
//
// The 3x3 / 1x1 convolutions of layer1..4 run on the tcgen05 GEMM as im2col x [Cout, k*k*Cin] with the
// eval-mode BatchNorm folded into the weights and the bias on the host (engine.cu), ReLU / residual in the
// GEMM epilogue.  Activations are NHWC: fp16 for the next convolution, fp32 for the identity path.
#include "kernels.cuh"

namespace hmocr {
namespace {

// conv1: 7x7, stride 2, pad 3, 1 -> 64 channels (+ folded BN + ReLU).  images f32 [B,1,96,320] -> fp16 NHWC [B,48,160,64]
// as an implicit GEMM on mma.sync: M = 16 consecutive output pixels of one row (a warp's work item), K = 49
// taps padded to 64, N = 64 channels.  The folded weights sit in shared memory as the fp16 B operand
// ([channel][tap], 144-byte rows: conflict-free ldmatrix); a warp copies the 7 x 37 input patch of its item into
// a private fp32 buffer (zero padded at the image border) and builds the A fragments straight from it: tap k of
// pixel p is patch[k / 7][2 p + k % 7], each lane's 16 tap offsets are fixed and live in registers.
constexpr int C1_WARPS = 8, C1_WP = 72, C1_PP = 40;
struct Conv1Smem {
  h16 w[64][C1_WP];
  float bias[64];
  float patch[C1_WARPS][8][C1_PP];      // row 7 stays zero: where the padding taps 49..63 point
};

__global__ void __launch_bounds__(C1_WARPS * 32, 2) conv7x7_kernel(const float* __restrict__ img, int B,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias, h16* __restrict__ out) {
  __shared__ __align__(16) Conv1Smem s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 2, t4 = lane & 3;
  for (int i = threadIdx.x; i < 64 * 64; i += C1_WARPS * 32) {
    const int ch = i >> 6, k = i & 63;
    s.w[ch][k] = to_h16(k < 49 ? __ldg(w + ch * 49 + k) : 0.f);       // w is [64][49]
  }
  if (threadIdx.x < 64) s.bias[threadIdx.x] = __ldg(bias + threadIdx.x);
  for (int i = lane; i < 8 * C1_PP; i += 32) (&s.patch[warp][0][0])[i] = 0.f;
  __syncthreads();
  float(*patch)[C1_PP] = s.patch[warp];
  // fixed per lane: patch offsets (floats) of taps 16 ks + 2 t4 + {0, 1, 8, 9}, relative to column 2 p
  int toff[4][4];
#pragma unroll
  for (int ks = 0; ks < 4; ++ks)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 16 * ks + 2 * t4 + (j & 1) + 8 * (j >> 1);
      toff[ks][j] = k < 49 ? (k / 7) * C1_PP + (k % 7) : 7 * C1_PP;
    }
  const uint32_t w_lane = smem_u32(&s.w[lane & 7][(lane >> 3) * 8]);
  const int items = B * 48 * 10;
  for (int it = blockIdx.x * C1_WARPS + warp; it < items; it += gridDim.x * C1_WARPS) {
    const int xt = it % 10, oy = (it / 10) % 48, b = it / 480;
    const float* ib = img + (size_t)b * 96 * 320;
    const int ix0 = 32 * xt - 3, iy0 = 2 * oy - 3;
    __syncwarp();                                   // the previous item's fragments have been read
#pragma unroll
    for (int r = 0; r < 7; ++r) {
      const int iy = iy0 + r;
      const bool rok = iy >= 0 && iy < 96;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = lane + 32 * half, ix = ix0 + c;
        if (c < 37) patch[r][c] = (rok && ix >= 0 && ix < 320) ? __ldg(ib + iy * 320 + ix) : 0.f;
      }
    }
    __syncwarp();
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const float* p0 = &patch[0][2 * g4];            // pixel g4; pixel g4 + 8 is 16 columns further
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {                // k-steps 2 kp, 2 kp + 1: one ldmatrix.x4 per channel tile
      uint32_t a[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int ks = 2 * kp + h;
        a[h][0] = pack16(p0[toff[ks][0]], p0[toff[ks][1]]);
        a[h][1] = pack16(p0[toff[ks][0] + 16], p0[toff[ks][1] + 16]);
        a[h][2] = pack16(p0[toff[ks][2]], p0[toff[ks][3]]);
        a[h][3] = pack16(p0[toff[ks][2] + 16], p0[toff[ks][3] + 16]);
      }
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        uint32_t bw[4];
        asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                     : "=r"(bw[0]), "=r"(bw[1]), "=r"(bw[2]), "=r"(bw[3])
                     : "r"(w_lane + (nt * 8 * C1_WP + kp * 32) * 2));
#pragma unroll
        for (int h = 0; h < 2; ++h)
          asm volatile(
              "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
              "{%0, %1, %2, %3};"
              : "+f"(acc[nt][0]), "+f"(acc[nt][1]), "+f"(acc[nt][2]), "+f"(acc[nt][3])
              : "r"(a[h][0]), "r"(a[h][1]), "r"(a[h][2]), "r"(a[h][3]), "r"(bw[2 * h]), "r"(bw[2 * h + 1]));
      }
    }
    h16* o0 = out + ((size_t)(b * 48 + oy) * 160 + xt * 16 + g4) * 64 + 2 * t4;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 bb = *reinterpret_cast<const float2*>(&s.bias[nt * 8 + 2 * t4]);
      *reinterpret_cast<uint32_t*>(o0 + nt * 8) = pack16(fmaxf(acc[nt][0] + bb.x, 0.f), fmaxf(acc[nt][1] + bb.y, 0.f));
      *reinterpret_cast<uint32_t*>(o0 + 8 * 64 + nt * 8) = pack16(fmaxf(acc[nt][2] + bb.x, 0.f), fmaxf(acc[nt][3] + bb.y, 0.f));
    }
  }
}

// maxpool 3x3, stride 2, pad 1: fp16 NHWC [B,H,W,C] -> fp16 + fp32 NHWC [B,H/2,W/2,C]   (8 channels per thread)
__global__ void __launch_bounds__(256) maxpool_kernel(const h16* __restrict__ in, int B, int H, int W, int C,
                                                      h16* __restrict__ out16, float* __restrict__ out32) {
  const int Ho = H / 2, Wo = W / 2, cgs = C / 8;
  const size_t total = (size_t)B * Ho * Wo * cgs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i % cgs;
    const size_t pix = i / cgs;
    const int ox = pix % Wo, oy = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
    float m[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) m[c] = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
      const int iy = oy * 2 - 1 + ky;
      if (iy < 0 || iy >= H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int ix = ox * 2 - 1 + kx;
        if (ix < 0 || ix >= W) continue;
        const uint4 u = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + iy) * W + ix) * C + cg * 8);
        const float2 a = unpack16(u.x), bb = unpack16(u.y), cc = unpack16(u.z), d = unpack16(u.w);
        m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], bb.x); m[3] = fmaxf(m[3], bb.y);
        m[4] = fmaxf(m[4], cc.x); m[5] = fmaxf(m[5], cc.y); m[6] = fmaxf(m[6], d.x); m[7] = fmaxf(m[7], d.y);
      }
    }
    *reinterpret_cast<uint4*>(out16 + pix * C + cg * 8) =
        make_uint4(pack16(m[0], m[1]), pack16(m[2], m[3]), pack16(m[4], m[5]), pack16(m[6], m[7]));
    float4* o = reinterpret_cast<float4*>(out32 + pix * C + cg * 8);
    o[0] = make_float4(m[0], m[1], m[2], m[3]);
    o[1] = make_float4(m[4], m[5], m[6], m[7]);
  }
}

// im2col for a k x k convolution (k = 3, pad 1 or k = 1, pad 0) with stride s:
// fp16 NHWC [B,H,W,C] -> fp16 [B*Ho*Wo, k*k*C], column order (ky, kx, c); 16 bytes (8 channels) per thread
__global__ void __launch_bounds__(256) im2col_kernel(const h16* __restrict__ in, int B, int H, int W, int C, int k,
                                                     int stride, int pad, int Ho, int Wo, h16* __restrict__ out) {
  const int cgs = C / 8, kk = k * k;
  const size_t total = (size_t)B * Ho * Wo * kk * cgs;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i % cgs;
    const int tap = (i / cgs) % kk;
    const size_t pix = i / ((size_t)cgs * kk);
    const int ox = pix % Wo, oy = (pix / Wo) % Ho, b = pix / ((size_t)Wo * Ho);
    const int iy = oy * stride - pad + tap / k, ix = ox * stride - pad + tap % k;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(in + (((size_t)b * H + iy) * W + ix) * C + cg * 8);
    *reinterpret_cast<uint4*>(out + (pix * kk + tap) * C + cg * 8) = v;
  }
}

// AdaptiveAvgPool2d((1, None)) over the 3 rows of [B,3,10,512] (fp32 NHWC) -> fp16 [B*10, 512]
__global__ void __launch_bounds__(256) avgpool_h_kernel(const float* __restrict__ in, int B, int H, int W, int C,
                                                        h16* __restrict__ out) {
  const size_t total = (size_t)B * W * C / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c2 = i % (C / 2);
    const int x = (i / (C / 2)) % W, b = i / ((size_t)(C / 2) * W);
    float a0 = 0.f, a1 = 0.f;
    for (int y = 0; y < H; ++y) {
      const float2 v = *reinterpret_cast<const float2*>(in + (((size_t)b * H + y) * W + x) * C + 2 * c2);
      a0 += v.x; a1 += v.y;
    }
    const float inv = 1.0f / H;
    *reinterpret_cast<uint32_t*>(out + ((size_t)b * W + x) * C + 2 * c2) = pack16(a0 * inv, a1 * inv);
  }
}

// x[b, w, :] + pos[w, :]  ->  rows (w * B + b): the [10, B, d] order the reference feeds its batch_first encoder
// (src/model_res18trans.py:57-61), fp32 residual stream + fp16 GEMM operand
__global__ void __launch_bounds__(256) add_pos_permute_kernel(const float* __restrict__ x, const float* __restrict__ pos,
                                                              int B, int S, int d, float* __restrict__ o32,
                                                              h16* __restrict__ o16) {
  const size_t total = (size_t)B * S * d / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c2 = i % (d / 2);
    const size_t r = i / (d / 2);                  // b * S + w
    const int w = r % S, b = r / S;
    const float2 v = *reinterpret_cast<const float2*>(x + r * d + 2 * c2);
    const float2 pp = *reinterpret_cast<const float2*>(pos + (size_t)w * d + 2 * c2);
    const float2 y = make_float2(v.x + pp.x, v.y + pp.y);
    const size_t ro = (size_t)w * B + b;
    *reinterpret_cast<float2*>(o32 + ro * d + 2 * c2) = y;
    *reinterpret_cast<uint32_t*>(o16 + ro * d + 2 * c2) = pack16(y.x, y.y);
  }
}

// rows (w * B + b) -> [B, S, d]  (the final permute(1, 0, 2), src/model_res18trans.py:64), fp32 + fp16
__global__ void __launch_bounds__(256) permute_back_kernel(const float* __restrict__ x, int B, int S, int d,
                                                           float* __restrict__ o32, h16* __restrict__ o16) {
  const size_t total = (size_t)B * S * d / 2;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c2 = i % (d / 2);
    const size_t r = i / (d / 2);                  // w * B + b
    const int b = r % B, w = r / B;
    const float2 v = *reinterpret_cast<const float2*>(x + r * d + 2 * c2);
    const size_t ro = (size_t)b * S + w;
    *reinterpret_cast<float2*>(o32 + ro * d + 2 * c2) = v;
    *reinterpret_cast<uint32_t*>(o16 + ro * d + 2 * c2) = pack16(v.x, v.y);
  }
}

// ToTensor + Normalize(0.5, 0.5) of app/src/preprocess.py:7-12 / src/predict.py:36-41 on the device:
// uint8 [B,96,320] (grayscale, already 96 x 320) -> f32 [B,1,96,320] = (u / 255 - 0.5) / 0.5, same operation order
// and IEEE divisions as torchvision, so the result is bit-identical to the host transform.  16 pixels per thread.
__global__ void __launch_bounds__(256) preprocess_u8_kernel(const uint8_t* __restrict__ in, size_t n16, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in) + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float4* o = reinterpret_cast<float4*>(out) + i * 4;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[4];
#pragma unroll
      for (int b = 0; b < 4; ++b)
        v[b] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)((w[q] >> (8 * b)) & 0xffu), 255.0f), 0.5f), 0.5f);
      o[q] = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
}

int blocks_for(size_t total) {
  size_t b = (total + 255) / 256;
  if (b > 148 * 16) b = 148 * 16;
  return (int)(b < 1 ? 1 : b);
}

}  // namespace

int preprocess_u8(cudaStream_t st, const uint8_t* in, size_t pixels, float* out) {
  HM_CHECK(pixels % 16 == 0, "preprocess: pixel count must be a multiple of 16");
  preprocess_u8_kernel<<<blocks_for(pixels / 16), 256, 0, st>>>(in, pixels / 16, out);
  HM_LAUNCHED();
  return 0;
}
int conv7x7_bn_relu(cudaStream_t st, const float* images, int B, const float* w, const float* bias, h16* out) {
  int blocks = ceil_div(B * 48 * 10, C1_WARPS);
  if (blocks > 148 * 2) blocks = 148 * 2;
  conv7x7_kernel<<<blocks, C1_WARPS * 32, 0, st>>>(images, B, w, bias, out);
  HM_LAUNCHED();
  return 0;
}
int maxpool3x3s2(cudaStream_t st, const h16* in, int B, int H, int W, int C, h16* out16, float* out32) {
  HM_CHECK(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "maxpool: unsupported shape");
  maxpool_kernel<<<blocks_for((size_t)B * (H / 2) * (W / 2) * (C / 8)), 256, 0, st>>>(in, B, H, W, C, out16, out32);
  HM_LAUNCHED();
  return 0;
}
int im2col(cudaStream_t st, const h16* in, int B, int H, int W, int C, int k, int stride, int pad, h16* out) {
  HM_CHECK(C % 8 == 0 && (k == 1 || k == 3), "im2col: unsupported shape");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  im2col_kernel<<<blocks_for((size_t)B * Ho * Wo * k * k * (C / 8)), 256, 0, st>>>(in, B, H, W, C, k, stride, pad, Ho, Wo, out);
  HM_LAUNCHED();
  return 0;
}
int avgpool_h(cudaStream_t st, const float* in, int B, int H, int W, int C, h16* out) {
  avgpool_h_kernel<<<blocks_for((size_t)B * W * C / 2), 256, 0, st>>>(in, B, H, W, C, out);
  HM_LAUNCHED();
  return 0;
}
int add_pos_permute(cudaStream_t st, const float* x, const float* pos, int B, int S, int d, float* o32, h16* o16) {
  add_pos_permute_kernel<<<blocks_for((size_t)B * S * d / 2), 256, 0, st>>>(x, pos, B, S, d, o32, o16);
  HM_LAUNCHED();
  return 0;
}
int permute_back(cudaStream_t st, const float* x, int B, int S, int d, float* o32, h16* o16) {
  permute_back_kernel<<<blocks_for((size_t)B * S * d / 2), 256, 0, st>>>(x, B, S, d, o32, o16);
  HM_LAUNCHED();
  return 0;
}

}  // namespace hmocr
