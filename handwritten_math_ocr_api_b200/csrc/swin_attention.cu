// Fused shifted-window attention on tensor cores (sm_100a).            swin_transformer.py:151-214,219-227
//
// One WARP = one (window, head) at a time.  The eight warps of a CTA work on the SAME head, so the head's
// relative-position bias [49][49] sits once in the CTA's shared memory; each warp loops over windows with
// private 49-row Q/K/V tiles (12 KB), which leaves room for 16 warps per SM.  Per item:
//   * the cyclic shift, zero padding, window partition and their inverses are index arithmetic on
//     the un-shifted, un-padded qkv buffer (rolled position (r,c) reads padded position
//     ((r+sh)%Hp, (c+sw)%Wp)); a padded token has q = k = v = qkv.bias; its output is never written;
//   * Q, K, V rows (49 x 32 fp16 each) are gathered with 16-byte cp.async, 8 tokens x 4 pieces per pass;
//     the rows 49..63 the 16-row MMA tiles reach for are one shared 16-byte line of zeros (every lane
//     hands ldmatrix its own row address);
//   * S = Q K^T (mma.sync m16n8k16, 4 m-tiles x 7 n-tiles x 2 k-steps), scaled by 32^-0.5, plus bias,
//     plus the -100 region mask; softmax in fp32 registers with quad shuffles;
//   * O = P V with P re-packed from the S accumulators straight into A fragments and V read through
//     ldmatrix.trans; rows are normalised and scattered back to the original token order.
// One __syncthreads (the bias table); after that every warp owns its tiles (__syncwarp only).
// 96.5 % of the encoder FLOPs are the tcgen05 GEMMs; this kernel's job is to stop the 49x49
// attention + the roll/partition copies from costing more time than they do FLOPs.
#include "kernels.cuh"

namespace hmocr {
namespace {

constexpr int WS = 7, WN = 49, HD = 32;
constexpr int TP = 40;            // tile pitch (fp16): 80-byte rows are conflict-free for ldmatrix
constexpr int BP = 52;            // bias pitch (fp32)
constexpr int WARPS = 8;

struct WarpTile {
  h16 q[WN][TP];
  h16 k[WN][TP];
  h16 v[WN][TP];
  int tok[WN + 3];            // token row in the qkv buffer, -1 if padded
  uint8_t region[64];
};
struct CtaShared {
  float bias[WN][BP];         // this head's relative-position bias, times log2(e)
  uint4 zero[4];              // what ldmatrix reads for tile rows >= 49 (64 bytes: a row address may be advanced by 32)
  WarpTile w[WARPS];
};

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2_trans(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__global__ void __launch_bounds__(WARPS * 32, 2) window_attn_mma_kernel(const h16* __restrict__ qkv,
                                                                        const float* __restrict__ qkv_bias,
                                                                        const float* __restrict__ rel_bias, int B, int H,
                                                                        int W, int C, int heads, int sh, int sw, int Hp,
                                                                        int Wp, h16* __restrict__ ctx) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  CtaShared& cs = *reinterpret_cast<CtaShared*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpTile& s = cs.w[warp];
  pdl_launch_dependents();
  const int per_head = gridDim.x / heads;                    // CTAs working on one head
  if (per_head == 0 || (int)blockIdx.x >= per_head * heads) return;
  const int h = blockIdx.x % heads;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int nww = Wp / WS, nwh = Hp / WS, nwin = nwh * nww;
  const int total = B * nwin;
  const bool masked = (sh + sw) > 0;
  // softmax on ex2: log2(e) is folded into the q scale (32^-0.5, swin_transformer.py:188), the bias table and the mask
  const float LOG2E = 1.4426950408889634f;
  const float scale = 0.17677669529663687f * LOG2E;
  const float mask_val = -100.0f * LOG2E;

  for (int i = threadIdx.x; i < WN * WN; i += WARPS * 32)
    cs.bias[i / WN][i % WN] = __ldg(rel_bias + (size_t)h * WN * WN + i) * LOG2E;
  for (int i = threadIdx.x; i < WN; i += WARPS * 32) cs.bias[i][WN] = 0.f;   // column 49 is read (and discarded) by the float2 loads
  if (threadIdx.x < 4) cs.zero[threadIdx.x] = make_uint4(0u, 0u, 0u, 0u);
  // q = k = v of a padded token: this head's slice of qkv.bias, the 16-byte piece this lane copies
  uint4 padv[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(qkv_bias + m * C + h * HD + t4 * 8));
    const float4 c4 = __ldg(reinterpret_cast<const float4*>(qkv_bias + m * C + h * HD + t4 * 8 + 4));
    padv[m] = make_uint4(pack16(a.x, a.y), pack16(a.z, a.w), pack16(c4.x, c4.y), pack16(c4.z, c4.w));
  }
  __syncthreads();
  pdl_wait();              // qkv comes from the previous kernel

  const uint32_t zero_addr = smem_u32(&cs.zero[0]);
  const uint32_t q_addr = smem_u32(&s.q[0][0]), k_addr = smem_u32(&s.k[0][0]), v_addr = smem_u32(&s.v[0][0]);
  // ldmatrix row addresses of this lane (tile rows 0..47 are real; of rows 48..63 only row 48 exists)
  const uint32_t q_lane = q_addr + (lane & 15) * (TP * 2) + (lane >> 4) * 16;
  const uint32_t q_last = (lane & 15) == 0 ? q_addr + 48 * (TP * 2) + (lane >> 4) * 16 : zero_addr;
  const uint32_t k_lane = k_addr + (lane & 7) * (TP * 2) + ((lane >> 3) & 1) * 16;
  const uint32_t k_last = (lane & 7) == 0 ? k_addr + 48 * (TP * 2) + ((lane >> 3) & 1) * 16 : zero_addr;
  const uint32_t v_lane = v_addr + (lane & 15) * (TP * 2);
  const bool v_last_ok = (lane & 15) == 0;

  const int item0 = (blockIdx.x / heads) * WARPS + warp, item_step = per_head * WARPS;
  for (int wi = item0; wi < total; wi += item_step) {
    const int b = wi / nwin, win = wi - b * nwin;
    const int wr = win / nww, wc = win - wr * nww;
    // most windows lie in ONE region of the shift mask: only the last window row / column is cut by the roll
    const bool mixed = masked && ((sh > 0 && wr == nwh - 1) || (sw > 0 && wc == nww - 1));
#pragma unroll
    for (int pass = 0; pass < 7; ++pass) {
      const int p = pass * 8 + g4;                               // token slot of the window; this lane copies piece t4
      if (p < WN) {
        const int pi = (p * 37) >> 8;                            // p / 7 for p < 64
        const int r = wr * WS + pi, c = wc * WS + (p - pi * WS); // rolled, padded coordinates
        int pr = r + sh, pc = c + sw;                            // padded coordinates before the roll
        if (pr >= Hp) pr -= Hp;
        if (pc >= Wp) pc -= Wp;
        const int tok = (pr < H && pc < W) ? (b * H + pr) * W + pc : -1;
        if (t4 == 0) {
          s.tok[p] = tok;
          if (mixed) {
            // slices (0,-7),(-7,-s),(-s,None) written in order; s == 0 makes the last one cover everything
            const int hb = (sh == 0) ? 2 : ((r >= Hp - WS) + (r >= Hp - sh));
            const int wb = (sw == 0) ? 2 : ((c >= Wp - WS) + (c >= Wp - sw));
            s.region[p] = (uint8_t)(hb * 3 + wb);
          }
        }
        const uint32_t off = p * (TP * 2) + t4 * 16;
        if (tok >= 0) {
          const h16* src = qkv + (size_t)tok * 3 * C + h * HD + t4 * 8;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(q_addr + off), "l"(src) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(k_addr + off), "l"(src + C) : "memory");
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(v_addr + off), "l"(src + 2 * C) : "memory");
        } else {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(q_addr + off), "r"(padv[0].x), "r"(padv[0].y), "r"(padv[0].z), "r"(padv[0].w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(k_addr + off), "r"(padv[1].x), "r"(padv[1].y), "r"(padv[1].z), "r"(padv[1].w) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(v_addr + off), "r"(padv[2].x), "r"(padv[2].y), "r"(padv[2].z), "r"(padv[2].w) : "memory");
        }
      }
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();

#pragma unroll 1
    for (int mt = 0; mt < 4; ++mt) {
      // a tile of 16 query slots that are ALL padding (the window rows below the image: half of every bottom window at
      // stages 1 / 2, a quarter of every window at stage 3, half at stage 4) produces nothing that is written
      {
        const int pq = mt * 16 + (lane & 15);
        if (!__any_sync(0xffffffffu, pq < WN && s.tok[pq] >= 0)) continue;
      }
      uint32_t aq[2][4];
      const uint32_t qa = mt < 3 ? q_lane + mt * 16 * (TP * 2) : q_last;
      ldsm4(qa, aq[0]);
      ldsm4(qa + 32, aq[1]);
      const int r0 = mt * 16 + g4, r1 = r0 + 8;
      const int rb0 = min(r0, WN - 1), rb1 = min(r1, WN - 1);       // clamp for the bias / region lookups
      int reg0 = 0, reg1 = 0;
      if (mixed) { reg0 = s.region[rb0]; reg1 = s.region[rb1]; }
      float sc[7][4];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t bk[2];
        const uint32_t ka = nt < 6 ? k_lane + nt * 8 * (TP * 2) : k_last;
        ldsm2(ka, bk);
        mma16816(c, aq[0], bk);
        ldsm2(ka + 32, bk);
        mma16816(c, aq[1], bk);
        const int c0 = nt * 8 + 2 * t4;                      // columns c0, c0 + 1 (c0 is even, <= 54)
        const float2 b0 = *reinterpret_cast<const float2*>(&cs.bias[rb0][min(c0, WN - 1) & ~1]);
        const float2 b1 = *reinterpret_cast<const float2*>(&cs.bias[rb1][min(c0, WN - 1) & ~1]);
        float v[4] = {fmaf(c[0], scale, b0.x), fmaf(c[1], scale, b0.y), fmaf(c[2], scale, b1.x), fmaf(c[3], scale, b1.y)};
        if (mixed) {
          const int rc0 = s.region[min(c0, WN - 1)], rc1 = s.region[min(c0 + 1, WN - 1)];
          if (rc0 != reg0) v[0] += mask_val;
          if (rc1 != reg0) v[1] += mask_val;
          if (rc0 != reg1) v[2] += mask_val;
          if (rc1 != reg1) v[3] += mask_val;
        }
        if (c0 >= WN) { v[0] = -INFINITY; v[2] = -INFINITY; }
        if (c0 + 1 >= WN) { v[1] = -INFINITY; v[3] = -INFINITY; }
#pragma unroll
        for (int e = 0; e < 4; ++e) sc[nt][e] = v[e];
        mx0 = fmaxf(mx0, fmaxf(v[0], v[1])); mx1 = fmaxf(mx1, fmaxf(v[2], v[3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 7; ++nt) {
        const float p0 = ex2_approx(sc[nt][0] - mx0), p1 = ex2_approx(sc[nt][1] - mx0);
        const float p2 = ex2_approx(sc[nt][2] - mx1), p3 = ex2_approx(sc[nt][3] - mx1);
        sum0 += p0 + p1; sum1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack16(p0, p1);        // row g4   , keys nt*8 + 2*t4 ..
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack16(p2, p3);    // row g4+8
      }
      pa[3][2] = 0u; pa[3][3] = 0u;                            // keys 56..63 do not exist
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1); sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1); sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
      const int tok0 = (r0 < WN) ? s.tok[r0] : -1, tok1 = (r1 < WN) ? s.tok[r1] : -1;
#pragma unroll
      for (int nt2 = 0; nt2 < 4; ++nt2) {
        float o[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t bv[2];
          const uint32_t va = kk < 3 ? v_lane + kk * 16 * (TP * 2) + nt2 * 16
                                     : (v_last_ok ? v_addr + 48 * (TP * 2) + nt2 * 16 : zero_addr);
          ldsm2_trans(va, bv);
          mma16816(o, pa[kk], bv);
        }
        const int col = h * HD + nt2 * 8 + 2 * t4;
        if (tok0 >= 0) *reinterpret_cast<uint32_t*>(ctx + (size_t)tok0 * C + col) = pack16(o[0] * inv0, o[1] * inv0);
        if (tok1 >= 0) *reinterpret_cast<uint32_t*>(ctx + (size_t)tok1 * C + col) = pack16(o[2] * inv1, o[3] * inv1);
      }
    }
    __syncwarp();      // all lanes are done with the tiles before the next item overwrites them
  }
}

// ------------------------------------------------------------------------------------------------
// Full (non-causal) multi-head self-attention on tensor cores, head_dim 32, up to 256 keys: the
// nn.TransformerEncoder of the ResNet-18 variant (/root/reference/src/model_res18trans.py:62-64, which attends
// ACROSS the batch: "sequence" = the images of the batch, "batch" = the 10 feature columns).
// One CTA = one (sequence b, head h): K and V of the head sit in shared memory (padded 80-byte rows), each
// warp takes 16-query tiles and runs a flash-style online softmax over 64-key chunks; S = Q K^T and O = P V
// on mma.sync with P re-packed from the S accumulators (same fragment algebra as the window kernel above).
// ------------------------------------------------------------------------------------------------
constexpr int FA_MAXT = 256, FA_WARPS = 8;
struct FullAttnSmem {
  h16 k[FA_MAXT][TP];
  h16 v[FA_MAXT][TP];
};

__global__ void __launch_bounds__(FA_WARPS * 32) full_attn_mma_kernel(const h16* __restrict__ qkv, int T, int nhead,
                                                                     h16* __restrict__ ctx) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FullAttnSmem& s = *reinterpret_cast<FullAttnSmem*>(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g4 = lane >> 2, t4 = lane & 3;
  const int b = blockIdx.x / nhead, h = blockIdx.x - b * nhead;
  const int d = nhead * HD, pitch = 3 * d;
  const h16* base = qkv + (size_t)b * T * pitch + h * HD;
  const int Tp = (T + 63) & ~63;                               // keys are processed 64 at a time
  pdl_launch_dependents();
  pdl_wait();
  for (int i = threadIdx.x; i < Tp * 8; i += FA_WARPS * 32) {
    const int j = i >> 3, m = (i >> 2) & 1, ch = i & 3;
    h16* dst = (m == 0 ? &s.k[j][0] : &s.v[j][0]) + ch * 8;
    if (j < T) {
      const h16* src = base + (size_t)j * pitch + (m + 1) * d + ch * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  const float scale = 0.17677669529663687f * 1.4426950408889634f;     // 32^-0.5 * log2(e): softmax on ex2
  const uint32_t k_lane = smem_u32(&s.k[lane & 7][((lane >> 3) & 1) * 8]);
  const uint32_t v_lane = smem_u32(&s.v[lane & 15][0]);
  for (int mt = warp; mt * 16 < T; mt += FA_WARPS) {
    const int r0 = mt * 16 + g4, r1 = r0 + 8;
    uint32_t aq[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const h16* q0 = base + (size_t)r0 * pitch + ks * 16 + 2 * t4;
      const h16* q1 = base + (size_t)r1 * pitch + ks * 16 + 2 * t4;
      aq[ks][0] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(q0)) : 0u;
      aq[ks][1] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(q1)) : 0u;
      aq[ks][2] = r0 < T ? __ldg(reinterpret_cast<const uint32_t*>(q0 + 8)) : 0u;
      aq[ks][3] = r1 < T ? __ldg(reinterpret_cast<const uint32_t*>(q1 + 8)) : 0u;
    }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    float o[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll 1
    for (int kc = 0; kc < Tp; kc += 64) {
      float sc[8][4];
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t bk[2];
        const uint32_t ka = k_lane + (kc + nt * 8) * (TP * 2);
        ldsm2(ka, bk);
        mma16816(c, aq[0], bk);
        ldsm2(ka + 32, bk);
        mma16816(c, aq[1], bk);
        const int c0 = kc + nt * 8 + 2 * t4;
        sc[nt][0] = c0 < T ? c[0] * scale : -INFINITY;
        sc[nt][1] = c0 + 1 < T ? c[1] * scale : -INFINITY;
        sc[nt][2] = c0 < T ? c[2] * scale : -INFINITY;
        sc[nt][3] = c0 + 1 < T ? c[3] * scale : -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float n0 = fmaxf(m0, mx0), n1 = fmaxf(m1, mx1);      // finite: every chunk holds at least one real key
      const float f0 = ex2_approx(m0 - n0), f1 = ex2_approx(m1 - n1);
      m0 = n0; m1 = n1;
      float sum0 = 0.f, sum1 = 0.f;
      uint32_t pa[4][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float p0 = ex2_approx(sc[nt][0] - n0), p1 = ex2_approx(sc[nt][1] - n0);
        const float p2 = ex2_approx(sc[nt][2] - n1), p3 = ex2_approx(sc[nt][3] - n1);
        sum0 += p0 + p1; sum1 += p2 + p3;
        pa[nt >> 1][(nt & 1) * 2] = pack16(p0, p1);
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack16(p2, p3);
      }
      l0 = l0 * f0 + sum0; l1 = l1 * f1 + sum1;                  // per-lane partial sums; quad-reduced at the end
#pragma unroll
      for (int nt2 = 0; nt2 < 4; ++nt2) {
        o[nt2][0] *= f0; o[nt2][1] *= f0; o[nt2][2] *= f1; o[nt2][3] *= f1;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t bv[2];
          ldsm2_trans(v_lane + (kc + kk * 16) * (TP * 2) + nt2 * 16, bv);
          mma16816(o[nt2], pa[kk], bv);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
    h16* out0 = ctx + (size_t)(b * T + r0) * d + h * HD + 2 * t4;
    h16* out1 = ctx + (size_t)(b * T + r1) * d + h * HD + 2 * t4;
#pragma unroll
    for (int nt2 = 0; nt2 < 4; ++nt2) {
      if (r0 < T) *reinterpret_cast<uint32_t*>(out0 + nt2 * 8) = pack16(o[nt2][0] * inv0, o[nt2][1] * inv0);
      if (r1 < T) *reinterpret_cast<uint32_t*>(out1 + nt2 * 8) = pack16(o[nt2][2] * inv1, o[nt2][3] * inv1);
    }
  }
}

}  // namespace

int window_attention(cudaStream_t st, const h16* qkv, const float* qkv_bias, const float* rel_bias, int B,
                     int H, int W, int C, int heads, int shift, h16* ctx) {
  HM_CHECK(C == heads * HD, "window_attention: head_dim must be 32 (C=%d heads=%d)", C, heads);
  HM_CHECK(C % 8 == 0, "window_attention: C must be a multiple of 8");
  const int Hp = ceil_div(H, WS) * WS, Wp = ceil_div(W, WS) * WS;
  const int sh = (Hp > WS) ? shift : 0, sw = (Wp > WS) ? shift : 0;   // swin_transformer.py:158-163
  const int smem = (int)sizeof(CtaShared);
  HM_DEVICE_ONCE(HM_CUDA(cudaFuncSetAttribute(window_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
  // CTAs are bound to heads (block % heads): 2 CTAs per SM (105 KB each), rounded down to a multiple of heads;
  // fewer when there are not enough windows to give every warp one.
  const long windows = (long)B * (Hp / WS) * (Wp / WS);
  int per_head = (148 * 2) / heads;
  const int need = (int)((windows + WARPS - 1) / WARPS);
  if (need < per_head) per_head = need;
  if (per_head < 1) per_head = 1;
  const int grid = per_head * heads;
  HM_CUDA(launch_pdl(window_attn_mma_kernel, dim3(grid), dim3(WARPS * 32), (size_t)smem, st, qkv, qkv_bias, rel_bias, B, H, W, C,
                     heads, sh, sw, Hp, Wp, ctx));
  HM_LAUNCHED();
  return 0;
}

// qkv fp16 [B*T, 3*nhead*32] (row = b*T + t) -> ctx fp16 [B*T, nhead*32]; T <= 256
int mha_full_mma(cudaStream_t st, const h16* qkv, int B, int T, int nhead, h16* ctx) {
  HM_CHECK(T >= 1 && T <= FA_MAXT, "mha_full_mma: T=%d out of range", T);
  HM_DEVICE_ONCE(HM_CUDA(cudaFuncSetAttribute(full_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FullAttnSmem))));
  HM_CUDA(launch_pdl(full_attn_mma_kernel, dim3(B * nhead), dim3(FA_WARPS * 32), sizeof(FullAttnSmem), st, qkv, T, nhead, ctx));
  HM_LAUNCHED();
  return 0;
}

}  // namespace hmocr
