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
// One thread = one output pixel x 8 channels; the 64 x 49 folded weights live in shared memory.
__global__ void __launch_bounds__(256) conv7x7_kernel(const float* __restrict__ img, int B, const float* __restrict__ w,
                                                      const float* __restrict__ bias, h16* __restrict__ out) {
  __shared__ float ws[49][64];
  __shared__ float bs[64];
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) ws[i % 49][i / 49] = w[i];       // w is [64][49]
  if (threadIdx.x < 64) bs[threadIdx.x] = bias[threadIdx.x];
  __syncthreads();
  const size_t total = (size_t)B * 48 * 160 * 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int cg = i & 7;
    const size_t pix = i >> 3;
    const int ox = pix % 160, oy = (pix / 160) % 48, b = pix / (160 * 48);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = bs[cg * 8 + c];
    const float* ib = img + (size_t)b * 96 * 320;
    for (int ky = 0; ky < 7; ++ky) {
      const int iy = oy * 2 - 3 + ky;
      if (iy < 0 || iy >= 96) continue;
#pragma unroll
      for (int kx = 0; kx < 7; ++kx) {
        const int ix = ox * 2 - 3 + kx;
        if (ix < 0 || ix >= 320) continue;
        const float v = __ldg(ib + iy * 320 + ix);
        const float* wr = &ws[ky * 7 + kx][cg * 8];
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[c] = fmaf(v, wr[c], acc[c]);
      }
    }
    *reinterpret_cast<uint4*>(out + pix * 64 + cg * 8) =
        make_uint4(pack16(fmaxf(acc[0], 0.f), fmaxf(acc[1], 0.f)), pack16(fmaxf(acc[2], 0.f), fmaxf(acc[3], 0.f)),
                   pack16(fmaxf(acc[4], 0.f), fmaxf(acc[5], 0.f)), pack16(fmaxf(acc[6], 0.f), fmaxf(acc[7], 0.f)));
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
  conv7x7_kernel<<<blocks_for((size_t)B * 48 * 160 * 8), 256, 0, st>>>(images, B, w, bias, out);
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
