// Image preprocessing of the reference on the GPU (SURVEY.md §8 f, row N2):
//   Grayscale(1) -> Resize((96, 320)) -> ToTensor() -> Normalize([0.5], [0.5])
//   (/root/reference/app/src/preprocess.py:6-16, /root/reference/src/predict.py:36-46)
// bit-identical to PIL + torchvision.  The arithmetic is Pillow's (third-party, not under /root/reference):
//   * Image.convert("L") for RGB input: L = (R*19595 + G*38470 + B*7471 + 0x8000) >> 16   (Convert.c, rgb2l);
//   * Image.resize(BILINEAR) = ImagingResample (Resample.c): a triangle filter whose support grows with the
//     down-scaling factor, coefficients normalised in double precision on the host and rounded to 22 fractional
//     bits, a horizontal pass and then a vertical pass, each accumulating in int32 from 1 << 21 and clipping >> 22
//     to uint8 (the intermediate image is uint8, as in Pillow);
//   * ToTensor + Normalize: ((u8 / 255) - 0.5) / 0.5, every step rounded to fp32.
// The training / evaluation loader's route (/root/reference/src/data_loader.py:31-35) is cv2.resize(gray, (320, 96))
// = OpenCV's 8-bit INTER_LINEAR (resize.cpp: 11-bit coefficients from a float source position, horizontal taps clamped
// with their weight moved inwards, vertical taps clamped by row, (((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2,
// exact 2x down-scales rerouted to the 2x2 box mean) followed by the same ToTensor + Normalize: resize_cv2_kernel.
// oracle/preprocess.py restates both algorithms in numpy and is pinned against the real libraries.
#include <cmath>
#include <vector>

#include "kernels.cuh"

namespace hmocr {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ int gray_at(const uint8_t* __restrict__ src, int channels, size_t idx) {
  if (channels == 1) return src[idx];
  const uint8_t* p = src + idx * 3;
  return (p[0] * 19595 + p[1] * 38470 + p[2] * 7471 + 0x8000) >> 16;
}

// horizontal pass (+ grayscale): src uint8 [H, W, channels] (row pitch in bytes) -> mid uint8 [H, out_w]
__global__ void __launch_bounds__(256) resize_h_kernel(const uint8_t* __restrict__ src, int channels, int H, int pitch,
                                                       const int* __restrict__ bounds, const int32_t* __restrict__ kk,
                                                       int ksize, int out_w, uint8_t* __restrict__ mid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * out_w) return;
  const int y = i / out_w, xx = i - y * out_w;
  const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
  const uint8_t* row = src + (size_t)y * pitch;
  const int32_t* k = kk + (size_t)xx * ksize;
  int acc = 1 << (PRECISION_BITS - 1);
  for (int x = 0; x < n; ++x) acc += gray_at(row, channels, xmin + x) * k[x];
  acc >>= PRECISION_BITS;
  mid[i] = (uint8_t)min(max(acc, 0), 255);
}

// vertical pass + ToTensor + Normalize: mid uint8 [H, out_w] -> out f32 [out_h, out_w]
__global__ void __launch_bounds__(256) resize_v_norm_kernel(const uint8_t* __restrict__ mid, int out_w,
                                                            const int* __restrict__ bounds,
                                                            const int32_t* __restrict__ kk, int ksize, int out_h,
                                                            float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_h * out_w) return;
  const int yy = i / out_w, x = i - yy * out_w;
  const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
  const int32_t* k = kk + (size_t)yy * ksize;
  int acc = 1 << (PRECISION_BITS - 1);
  for (int y = 0; y < n; ++y) acc += mid[(size_t)(ymin + y) * out_w + x] * k[y];
  acc >>= PRECISION_BITS;
  const float u = (float)min(max(acc, 0), 255);
  out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), 0.5f), 0.5f);
}

// cv2.resize INTER_LINEAR + ToTensor + Normalize, one thread per output pixel (no intermediate image):
// tab = [x0 | x1 | xa0 | xa1] (out_w each) then [y0 | y1 | yb0 | yb1] (out_h each)
__global__ void __launch_bounds__(256) resize_cv2_kernel(const uint8_t* __restrict__ src, int W, int area2x,
                                                         const int* __restrict__ tab, int out_h, int out_w,
                                                         float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_h * out_w) return;
  const int yy = i / out_w, xx = i - yy * out_w;
  int v;
  if (area2x) {
    const uint8_t* p = src + (size_t)(2 * yy) * W + 2 * xx;
    v = (p[0] + p[1] + p[W] + p[W + 1] + 2) >> 2;
  } else {
    const int* ty = tab + 4 * out_w;
    const int x0 = tab[xx], x1 = tab[out_w + xx], a0 = tab[2 * out_w + xx], a1 = tab[3 * out_w + xx];
    const int y0 = ty[yy], y1 = ty[out_h + yy], b0 = ty[2 * out_h + yy], b1 = ty[3 * out_h + yy];
    const uint8_t* r0 = src + (size_t)y0 * W;
    const uint8_t* r1 = src + (size_t)y1 * W;
    const int s0 = r0[x0] * a0 + r0[x1] * a1, s1 = r1[x0] * a0 + r1[x1] * a1;
    v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
  }
  const float u = (float)min(max(v, 0), 255);
  out[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(u, 255.0f), 0.5f), 0.5f);
}

// resize.cpp: source indices and 11-bit weights of one axis (clamp_weight = the horizontal rule)
void cv2_linear_coeffs(int ssize, int dsize, bool clamp_weight, int* i0, int* i1, int* w0, int* w1) {
  const double inv = (double)dsize / ssize, scale = 1.0 / inv;
  for (int d = 0; d < dsize; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)std::floor(f);
    f -= (float)s;
    if (clamp_weight) {
      if (s < 0) { f = 0.f; s = 0; }
      if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
      i0[d] = s;
      i1[d] = s + 1 < ssize ? s + 1 : ssize - 1;
    } else {
      i0[d] = s < 0 ? 0 : (s > ssize - 1 ? ssize - 1 : s);
      i1[d] = s + 1 < 0 ? 0 : (s + 1 > ssize - 1 ? ssize - 1 : s + 1);
    }
    w0[d] = (int)std::nearbyint((1.f - f) * 2048.f);       // cvRound: round half to even
    w1[d] = (int)std::nearbyint(f * 2048.f);
  }
}

// Resample.c: precompute_coeffs + normalize_coeffs_8bpc, bilinear filter over the whole axis (double precision)
void precompute_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int32_t>& kk, int* ksize_out) {
  const float in0 = 0.0f, in1 = (float)in_size;
  double scale = (double)(in1 - in0) / out_size, filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> k(ksize);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = in0 + (xx + 0.5) * scale;
    double ww = 0.0;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x) {
      if (ww != 0.0) k[x] /= ww;
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
    }
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
  }
  *ksize_out = ksize;
}

}  // namespace

size_t preprocess_table_ints(int H, int W, int out_h, int out_w) {      // upper bound of the coefficient tables, in ints
  auto ks = [](int in, int out) { double s = (double)in / out; if (s < 1.0) s = 1.0; return (size_t)((int)std::ceil(s) * 2 + 1); };
  return (size_t)out_w * (2 + ks(W, out_w)) + (size_t)out_h * (2 + ks(H, out_h));
}

// src_dev uint8 [H, W, channels] (dense), tables_dev >= preprocess_table_ints ints, mid_dev >= H * out_w bytes
int preprocess_image(cudaStream_t st, const uint8_t* src_dev, int channels, int H, int W, int out_h, int out_w,
                     int* tables_dev, uint8_t* mid_dev, float* out_dev) {
  HM_CHECK(channels == 1 || channels == 3, "preprocess: %d channels (1 = mode L, 3 = mode RGB)", channels);
  HM_CHECK(H >= 1 && W >= 1 && H <= 16384 && W <= 16384, "preprocess: image size %dx%d out of range", H, W);
  std::vector<int> bh, bv;
  std::vector<int32_t> kh, kv;
  int ksh = 0, ksv = 0;
  precompute_coeffs(W, out_w, bh, kh, &ksh);
  precompute_coeffs(H, out_h, bv, kv, &ksv);
  std::vector<int> host;
  host.reserve(bh.size() + kh.size() + bv.size() + kv.size());
  host.insert(host.end(), bh.begin(), bh.end());
  host.insert(host.end(), kh.begin(), kh.end());
  host.insert(host.end(), bv.begin(), bv.end());
  host.insert(host.end(), kv.begin(), kv.end());
  HM_CUDA(cudaMemcpyAsync(tables_dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, st));   // pageable: staged before return
  const int* d_bh = tables_dev;
  const int32_t* d_kh = tables_dev + bh.size();
  const int* d_bv = tables_dev + bh.size() + kh.size();
  const int32_t* d_kv = tables_dev + bh.size() + kh.size() + bv.size();
  resize_h_kernel<<<ceil_div(H * out_w, 256), 256, 0, st>>>(src_dev, channels, H, W * channels, d_bh, d_kh, ksh, out_w, mid_dev);
  HM_LAUNCHED();
  resize_v_norm_kernel<<<ceil_div(out_h * out_w, 256), 256, 0, st>>>(mid_dev, out_w, d_bv, d_kv, ksv, out_h, out_dev);
  HM_LAUNCHED();
  return 0;
}

// src_dev uint8 [H, W] dense, tables_dev >= 4 * (out_w + out_h) ints
int preprocess_gray_cv2(cudaStream_t st, const uint8_t* src_dev, int H, int W, int out_h, int out_w, int* tables_dev,
                        float* out_dev) {
  HM_CHECK(H >= 1 && W >= 1 && H <= 16384 && W <= 16384, "preprocess: image size %dx%d out of range", H, W);
  const int area2x = (W == 2 * out_w && H == 2 * out_h) ? 1 : 0;
  if (!area2x) {
    std::vector<int> host((size_t)4 * (out_w + out_h));
    int* tx = host.data();
    int* ty = host.data() + 4 * out_w;
    cv2_linear_coeffs(W, out_w, true, tx, tx + out_w, tx + 2 * out_w, tx + 3 * out_w);
    cv2_linear_coeffs(H, out_h, false, ty, ty + out_h, ty + 2 * out_h, ty + 3 * out_h);
    HM_CUDA(cudaMemcpyAsync(tables_dev, host.data(), host.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  }
  resize_cv2_kernel<<<ceil_div(out_h * out_w, 256), 256, 0, st>>>(src_dev, W, area2x, tables_dev, out_h, out_w, out_dev);
  HM_LAUNCHED();
  return 0;
}

}  // namespace hmocr
