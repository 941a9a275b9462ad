// Persistent cluster decode kernel (sm_100a).
//
// Replaces the whole greedy loop body of /root/reference/src/inference.py:18-25 — i.e. one
// DecoderTransformer.forward step (src/model_swin.py:72-88, 8 x torch TransformerDecoderLayer,
// post-LN) + fc_out + argmax — for a range of steps [t_begin, t_end) in ONE launch.
//
// Mapping: a thread-block CLUSTER of 8 CTAs owns 16 sequences end to end.  CTA rank c is
//   * attention head c (self- and cross-attention of its 16 rows for that head), and
//   * the c-th 1/8 column slice of every projection (q/k/v of head c, 32 columns of out_proj,
//     64 columns of linear1, 32 of linear2, V/8 vocabulary columns of fc_out).
// So each SM streams only 1/8 of the decoder weights per step (1.6 MB instead of 13 MB): the
// weight slices are pre-packed per (layer, CTA) and arrive in shared memory through a 4-slot ring
// of cp.async.bulk (TMA) copies issued 3 chunks ahead, completion on mbarriers.  The 16-row
// activations are exchanged between the 8 CTAs through distributed shared memory
// (st.shared::cluster into every peer) and ordered by hardware cluster barriers
// (barrier.cluster), 6 per layer + 1 per step; no grid-wide synchronisation, no global-memory
// round trip, no host involvement between steps.  M = 16 rows is exactly one mma.sync m16n8k16
// tile; the step is bandwidth/latency bound (13-15 MFLOP per token), so tensor-core rate is
// irrelevant here and tcgen05 (M >= 64) would idle 3/4 of its rows.
// Self-attention K/V (bf16) are appended to / streamed from the HBM cache with 512-byte
// coalesced warp loads (8 keys x 64 B per load instruction); the memory K/V of cross-attention
// are read the same way from a [layer][image][head][30][32] repack.
#include "decode_persistent.cuh"

namespace hmocr {
namespace {

constexpr int ROWS = 16, CL = 8, THREADS = 256;
constexpr int D = 256, FF = 512, HD = 32, NH = 8, MEM_S = 30;
constexpr int PD = D + 8, PF = FF + 8;            // padded operand pitches (elements): conflict-free ldmatrix
constexpr int NSLOT = 4, PRE = 3;
constexpr int SLOT = DP_CH_F1;
constexpr float ATT_SCALE = 0.17677669529663687f;
constexpr float LN_EPS = 1e-5f;

struct __align__(16) Partial { float m; int idx; float s; int pad; };

struct Smem {
  alignas(128) uint8_t slot[NSLOT][SLOT];
  alignas(16) float x32[ROWS][D];          // residual stream (replicated in every CTA of the cluster)
  alignas(16) float y32[ROWS][D];          // pre-LayerNorm rows gathered from the 8 column slices
  alignas(16) __nv_bfloat16 xa[ROWS][PD];  // LayerNorm output, bf16 A operand
  alignas(16) __nv_bfloat16 ctxf[ROWS][PD];   // attention context gathered from the 8 heads
  alignas(16) __nv_bfloat16 hf[ROWS][PF];  // relu(linear1) gathered from the 8 slices
  alignas(16) float qs[ROWS][HD];          // this head's scaled query
  Partial part[CL][ROWS];                  // per-CTA argmax / sum-exp partials (gathered)
  Partial wpart[8][ROWS];                  // per-warp partials
  int tok[ROWS];
  alignas(8) uint64_t full[NSLOT];
};

// ---- PTX helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t (&r)[2]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C[16 x 8] (one n-tile) = A[16 x K] (smem, pitch PA) x W[ntile*8.., K]^T (smem, pitch K+8)
//   c[0],c[1]: row lane/4, cols 2*(lane%4), +1;  c[2],c[3]: row lane/4 + 8
template <int K, int PA>
__device__ __forceinline__ void gemm_tile(const __nv_bfloat16* A, const __nv_bfloat16* W, int ntile, int lane,
                                          float (&c)[4]) {
  float c2[4] = {0.f, 0.f, 0.f, 0.f};
  c[0] = c[1] = c[2] = c[3] = 0.f;
  const uint32_t a_addr = smem_u32(A + (lane & 15) * PA + (lane >> 4) * 8);
  const uint32_t b_addr = smem_u32(W + (ntile * 8 + (lane & 7)) * (K + 8) + ((lane >> 3) & 1) * 8);
#pragma unroll
  for (int k0 = 0; k0 < K; k0 += 32) {
    uint32_t a[4], b[2];
    ldsm_x4(a_addr + k0 * 2, a);
    ldsm_x2(b_addr + k0 * 2, b);
    mma_bf16(c, a, b);
    ldsm_x4(a_addr + (k0 + 16) * 2, a);
    ldsm_x2(b_addr + (k0 + 16) * 2, b);
    mma_bf16(c2, a, b);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i] += c2[i];
}

__device__ __forceinline__ float dot8(const float (&q)[8], const uint4 u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  float s = q[0] * a.x;
  s = fmaf(q[1], a.y, s); s = fmaf(q[2], b.x, s); s = fmaf(q[3], b.y, s);
  s = fmaf(q[4], c.x, s); s = fmaf(q[5], c.y, s); s = fmaf(q[6], d.x, s); s = fmaf(q[7], d.y, s);
  return s;
}

// One query row against n keys.  Lane = (key group jg = lane/4, dim chunk dc = lane%4): every load
// instruction of the warp fetches 8 consecutive 64-byte K (or V) rows = 512 contiguous bytes.
// out[0..7] = context channels dc*8.. (valid in every lane after the key-group reduction).
template <int NI>
__device__ __forceinline__ void attend_row(const float* q, const __nv_bfloat16* Kb, const __nv_bfloat16* Vb, int n,
                                           int lane, float (&out)[8]) {
  const int jg = lane >> 2, dc = lane & 3;
  float qv[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) qv[e] = q[dc * 8 + e];
  float sc[NI];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    sc[i] = -INFINITY;
    if (i * 8 < n) {                                   // warp-uniform
      const int j = jg + 8 * i;
      float a = 0.f;
      if (j < n) a = dot8(qv, __ldcg(reinterpret_cast<const uint4*>(Kb + (size_t)j * HD + dc * 8)));
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      if (j < n) { sc[i] = a; mx = fmaxf(mx, a); }
    }
  }
  mx = warp_max(mx);
  float den = 0.f;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    if (i * 8 < n) {
      const int j = jg + 8 * i;
      if (j < n) {
        const float p = __expf(sc[i] - mx);
        den += p;
        const uint4 u = __ldcg(reinterpret_cast<const uint4*>(Vb + (size_t)j * HD + dc * 8));
        const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
        acc[0] = fmaf(p, a.x, acc[0]); acc[1] = fmaf(p, a.y, acc[1]); acc[2] = fmaf(p, b.x, acc[2]);
        acc[3] = fmaf(p, b.y, acc[3]); acc[4] = fmaf(p, c.x, acc[4]); acc[5] = fmaf(p, c.y, acc[5]);
        acc[6] = fmaf(p, d.x, acc[6]); acc[7] = fmaf(p, d.y, acc[7]);
      }
    }
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    den += __shfl_xor_sync(0xffffffffu, den, o);
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  }
  const float inv = 1.0f / den;
#pragma unroll
  for (int e = 0; e < 8; ++e) out[e] = acc[e] * inv;
}

__device__ __forceinline__ void merge_partial(Partial& a, const Partial& b) {
  // combine (max, first-argmax, sum exp(x - max)); ties -> lower index (torch.argmax)
  if (b.m > a.m || (b.m == a.m && b.idx < a.idx)) {
    const float sa = (a.m == -INFINITY) ? 0.f : a.s * __expf(a.m - b.m);
    a.s = sa + b.s; a.m = b.m; a.idx = b.idx;
  } else {
    const float sb = (b.m == -INFINITY) ? 0.f : b.s * __expf(b.m - a.m);
    a.s += sb;
  }
}
__device__ __forceinline__ void update_partial(Partial& a, float v, int idx) {
  if (v > a.m) { a.s = a.s * __expf(a.m - v) + 1.0f; a.m = v; a.idx = idx; }
  else a.s += __expf(v - a.m);
}

template <int NI>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 1)
decode_persistent_kernel(const DecPersistParams p, int t_begin, int t_end) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (int)cluster_ctarank();              // column slice == attention head
  const int row0 = (blockIdx.x / CL) * ROWS;
  const int g4 = lane >> 2, t4 = lane & 3;           // mma fragment coordinates
  const int L = p.num_layers;
  const int cps = 8 * L + p.fc_chunks;               // weight chunks per step
  const int total_chunks = (t_end - t_begin) * cps;
  const int cols_per_cta = p.fc_chunks * 64;

  auto issue = [&](int gi) {
    const int n = gi % cps;
    const uint8_t* src;
    uint32_t bytes;
    if (n < 8 * L) {
      const int l = n >> 3, ch = n & 7;
      src = p.wblob + ((size_t)l * CL + c) * DP_LAYER_CTA_BYTES +
            (ch < 6 ? ch * DP_CH_ATT : (ch == 6 ? 6 * DP_CH_ATT : 6 * DP_CH_ATT + DP_CH_F1));
      bytes = ch < 6 ? DP_CH_ATT : (ch == 6 ? DP_CH_F1 : DP_CH_F2);
    } else {
      src = p.fcblob + ((size_t)c * p.fc_chunks + (n - 8 * L)) * DP_CH_FC;
      bytes = DP_CH_FC;
    }
    uint64_t* bar = &s.full[gi & (NSLOT - 1)];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(s.slot[gi & (NSLOT - 1)], src, bytes, bar);
  };
  // wait for chunk g (and keep the ring PRE chunks ahead); the slot being refilled held chunk g-1,
  // whose readers all passed a block/cluster barrier before anyone gets here
  auto acquire = [&](int g) -> const __nv_bfloat16* {
    if (tid == 0 && g + PRE < total_chunks) issue(g + PRE);
    mbar_wait(&s.full[g & (NSLOT - 1)], (g >> 2) & 1);
    return reinterpret_cast<const __nv_bfloat16*>(s.slot[g & (NSLOT - 1)]);
  };

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(&s.full[i], 1);
    fence_barrier_init();
  }
  // x = embedding[token at t_begin] + pos[t_begin]
  {
    const int r = tid >> 4, c0 = (tid & 15) * 16;
    const int gr = row0 + r;
    long long tk = (gr < p.rows) ? p.tokens[(size_t)gr * p.ld_tok + t_begin] : 0;
    if (tk < 0 || tk >= p.vocab) tk = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gr < p.rows) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)tk * D + c0) + q);
        const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)t_begin * D + c0) + q);
        v = make_float4(e.x + ps.x, e.y + ps.y, e.z + ps.z, e.w + ps.w);
      }
      *reinterpret_cast<float4*>(&s.x32[r][c0 + 4 * q]) = v;
      *reinterpret_cast<uint2*>(&s.xa[r][c0 + 4 * q]) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
  }
  __syncthreads();
  if (tid == 0)
    for (int i = 0; i < PRE && i < total_chunks; ++i) issue(i);
  cluster_sync_all();      // every CTA of the cluster is resident and initialised before any DSMEM store

  const uint32_t ctxf_base = smem_u32(&s.ctxf[0][0]);
  const uint32_t y32_base = smem_u32(&s.y32[0][0]);
  const uint32_t hf_base = smem_u32(&s.hf[0][0]);
  const uint32_t part_base = smem_u32(&s.part[0][0]);

  // LayerNorm of the gathered rows: warp w owns rows 2w, 2w+1; lane owns 8 consecutive columns
  auto layer_norm = [&](const float* gamma, const float* beta) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr;
      const float4 a = *reinterpret_cast<const float4*>(&s.y32[r][lane * 8]);
      const float4 b = *reinterpret_cast<const float4*>(&s.y32[r][lane * 8 + 4]);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float sum = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) sum += v[e];
      const float mean = warp_sum(sum) * (1.0f / D);
      float sq = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[e] -= mean; sq += v[e] * v[e]; }
      const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + LN_EPS);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + lane * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + lane * 8 + 4));
      v[0] = v[0] * rstd * g0.x + b0.x; v[1] = v[1] * rstd * g0.y + b0.y; v[2] = v[2] * rstd * g0.z + b0.z;
      v[3] = v[3] * rstd * g0.w + b0.w; v[4] = v[4] * rstd * g1.x + b1.x; v[5] = v[5] * rstd * g1.y + b1.y;
      v[6] = v[6] * rstd * g1.z + b1.z; v[7] = v[7] * rstd * g1.w + b1.w;
      *reinterpret_cast<float4*>(&s.x32[r][lane * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&s.x32[r][lane * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<uint4*>(&s.xa[r][lane * 8]) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  };
  // query slice of head c: qs = (xa Wq^T + b) / sqrt(32)
  auto project_q = [&](const __nv_bfloat16* W, const float* bias) {
    if (warp < 4) {
      float acc[4];
      gemm_tile<D, PD>(&s.xa[0][0], W, warp, lane, acc);
      const int col = warp * 8 + 2 * t4;
      const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
      s.qs[g4][col] = (acc[0] + b0) * ATT_SCALE; s.qs[g4][col + 1] = (acc[1] + b1) * ATT_SCALE;
      s.qs[g4 + 8][col] = (acc[2] + b0) * ATT_SCALE; s.qs[g4 + 8][col + 1] = (acc[3] + b1) * ATT_SCALE;
    }
  };
  // out-projection slice (32 columns) + bias + residual -> y32 of every CTA in the cluster
  auto project_out = [&](const __nv_bfloat16* A, auto gemm, const float* bias) {
    if (warp < 4) {
      float acc[4];
      gemm(A, acc);
      const int col = c * 32 + warp * 8 + 2 * t4;
      const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
      const float y0 = acc[0] + b0 + s.x32[g4][col], y1 = acc[1] + b1 + s.x32[g4][col + 1];
      const float y2 = acc[2] + b0 + s.x32[g4 + 8][col], y3 = acc[3] + b1 + s.x32[g4 + 8][col + 1];
      const uint32_t o0 = y32_base + (g4 * D + col) * 4, o1 = y32_base + ((g4 + 8) * D + col) * 4;
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) {
        st_cluster_v2(mapa(o0, rk), __float_as_uint(y0), __float_as_uint(y1));
        st_cluster_v2(mapa(o1, rk), __float_as_uint(y2), __float_as_uint(y3));
      }
    }
  };
  // attention of rows 2w, 2w+1 for head c; context slice -> ctxf of every CTA
  auto attention = [&](auto kv_of_row, int nkeys) {
#pragma unroll 1
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp * 2 + rr;
      const int gr = row0 + r;
      float out[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) out[e] = 0.f;
      if (gr < p.rows) {                               // warp-uniform
        const __nv_bfloat16 *Kb, *Vb;
        kv_of_row(gr, Kb, Vb);
        attend_row<NI>(&s.qs[r][0], Kb, Vb, nkeys, lane, out);
      }
      if (lane < 4) {
        const uint4 v = make_uint4(pack_bf16(out[0], out[1]), pack_bf16(out[2], out[3]), pack_bf16(out[4], out[5]),
                                   pack_bf16(out[6], out[7]));
        const uint32_t o = ctxf_base + (r * PD + c * HD + lane * 8) * 2;
#pragma unroll
        for (int rk = 0; rk < CL; ++rk) st_cluster_v4(mapa(o, rk), v);
      }
    }
  };

  int g = 0;     // weight chunk counter of this launch
  for (int t = t_begin; t < t_end; ++t) {
    for (int l = 0; l < L; ++l) {
      const float* fp = p.fparams + (size_t)l * DP_FP_LAYER;
      const size_t cache_l = (size_t)l * p.rows;
      // ---- self-attention: q, k, v of head c ---------------------------------------------------
      {
        const __nv_bfloat16* W = acquire(g);
        project_q(W, fp + DP_FP_BIN + c * HD);
        __syncthreads();
        ++g;
      }
#pragma unroll 1
      for (int kv = 0; kv < 2; ++kv) {
        const __nv_bfloat16* W = acquire(g);
        if (warp < 4) {
          float acc[4];
          gemm_tile<D, PD>(&s.xa[0][0], W, warp, lane, acc);
          const int col = warp * 8 + 2 * t4;
          const float* bias = fp + DP_FP_BIN + (1 + kv) * D + c * HD;
          const float b0 = __ldg(bias + col), b1 = __ldg(bias + col + 1);
          __nv_bfloat16* cache = kv ? p.vcache : p.kcache;
          const int r0 = row0 + g4, r1 = row0 + g4 + 8;
          if (r0 < p.rows)
            *reinterpret_cast<uint32_t*>(cache + (((cache_l + r0) * NH + c) * p.tmax + t) * HD + col) =
                pack_bf16(acc[0] + b0, acc[1] + b1);
          if (r1 < p.rows)
            *reinterpret_cast<uint32_t*>(cache + (((cache_l + r1) * NH + c) * p.tmax + t) * HD + col) =
                pack_bf16(acc[2] + b0, acc[3] + b1);
        }
        __syncthreads();      // also publishes the appended K/V row to the attention warps of this CTA
        ++g;
      }
      attention(
          [&](int gr, const __nv_bfloat16*& Kb, const __nv_bfloat16*& Vb) {
            const size_t off = ((cache_l + gr) * NH + c) * (size_t)p.tmax * HD;
            Kb = p.kcache + off; Vb = p.vcache + off;
          },
          t + 1);
      cluster_sync_all();                                                        // #1 ctxf complete
      // ---- x = LN1(x + out_proj(ctx)) ---------------------------------------------------------------
      {
        const __nv_bfloat16* W = acquire(g);
        project_out(&s.ctxf[0][0], [&](const __nv_bfloat16* A, float (&acc)[4]) { gemm_tile<D, PD>(A, W, warp, lane, acc); },
                    fp + DP_FP_BO);
        cluster_sync_all();                                                      // #2 y32 complete
        ++g;
        layer_norm(fp + DP_FP_LN1G, fp + DP_FP_LN1B);
        __syncthreads();
      }
      // ---- cross-attention over the 30 memory tokens ----------------------------------------------
      {
        const __nv_bfloat16* W = acquire(g);
        project_q(W, fp + DP_FP_BCQ + c * HD);
        __syncthreads();
        ++g;
        attention(
            [&](int gr, const __nv_bfloat16*& Kb, const __nv_bfloat16*& Vb) {
              const int img = gr / p.beam;
              const size_t off = (((size_t)l * p.images + img) * NH + c) * (size_t)MEM_S * HD;
              Kb = p.memk + off; Vb = p.memv + off;
            },
            MEM_S);
        cluster_sync_all();                                                      // #3
      }
      {
        const __nv_bfloat16* W = acquire(g);
        project_out(&s.ctxf[0][0], [&](const __nv_bfloat16* A, float (&acc)[4]) { gemm_tile<D, PD>(A, W, warp, lane, acc); },
                    fp + DP_FP_BCO);
        cluster_sync_all();                                                      // #4
        ++g;
        layer_norm(fp + DP_FP_LN2G, fp + DP_FP_LN2B);
        __syncthreads();
      }
      // ---- feed-forward: 64 columns of linear1 (+ReLU) per CTA, then 32 columns of linear2 ---------
      {
        const __nv_bfloat16* W = acquire(g);
        float acc[4];
        gemm_tile<D, PD>(&s.xa[0][0], W, warp, lane, acc);
        const int col = c * 64 + warp * 8 + 2 * t4;
        const float b0 = __ldg(fp + DP_FP_B1 + col), b1 = __ldg(fp + DP_FP_B1 + col + 1);
        const uint32_t h0 = pack_bf16(fmaxf(acc[0] + b0, 0.f), fmaxf(acc[1] + b1, 0.f));
        const uint32_t h1 = pack_bf16(fmaxf(acc[2] + b0, 0.f), fmaxf(acc[3] + b1, 0.f));
        const uint32_t o0 = hf_base + (g4 * PF + col) * 2, o1 = hf_base + ((g4 + 8) * PF + col) * 2;
#pragma unroll
        for (int rk = 0; rk < CL; ++rk) {
          st_cluster_b32(mapa(o0, rk), h0);
          st_cluster_b32(mapa(o1, rk), h1);
        }
        cluster_sync_all();                                                      // #5 hf complete
        ++g;
      }
      {
        const __nv_bfloat16* W = acquire(g);
        project_out(&s.hf[0][0], [&](const __nv_bfloat16* A, float (&acc)[4]) { gemm_tile<FF, PF>(A, W, warp, lane, acc); },
                    fp + DP_FP_B2);
        cluster_sync_all();                                                      // #6
        ++g;
        layer_norm(fp + DP_FP_LN3G, fp + DP_FP_LN3B);
        __syncthreads();
      }
    }
    // ---- fc_out slice + running (max, argmax, sum-exp) ---------------------------------------------
    Partial pa, pb;        // rows g4 and g4+8 of this thread's columns
    pa.m = pb.m = -INFINITY; pa.idx = pb.idx = 0x7fffffff; pa.s = pb.s = 0.f; pa.pad = pb.pad = 0;
#pragma unroll 1
    for (int ch = 0; ch < p.fc_chunks; ++ch) {
      const __nv_bfloat16* W = acquire(g);
      float acc[4];
      gemm_tile<D, PD>(&s.xa[0][0], W, warp, lane, acc);
      const int v0 = c * cols_per_cta + ch * 64 + warp * 8 + 2 * t4;
      if (v0 < p.vocab) {
        const float b = __ldg(p.fc_bias + v0);
        update_partial(pa, acc[0] + b, v0);
        update_partial(pb, acc[2] + b, v0);
      }
      if (v0 + 1 < p.vocab) {
        const float b = __ldg(p.fc_bias + v0 + 1);
        update_partial(pa, acc[1] + b, v0 + 1);
        update_partial(pb, acc[3] + b, v0 + 1);
      }
      __syncthreads();
      ++g;
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      Partial oa, ob;
      oa.m = __shfl_xor_sync(0xffffffffu, pa.m, o); oa.idx = __shfl_xor_sync(0xffffffffu, pa.idx, o);
      oa.s = __shfl_xor_sync(0xffffffffu, pa.s, o);
      ob.m = __shfl_xor_sync(0xffffffffu, pb.m, o); ob.idx = __shfl_xor_sync(0xffffffffu, pb.idx, o);
      ob.s = __shfl_xor_sync(0xffffffffu, pb.s, o);
      merge_partial(pa, oa);
      merge_partial(pb, ob);
    }
    if (t4 == 0) { s.wpart[warp][g4] = pa; s.wpart[warp][g4 + 8] = pb; }
    __syncthreads();
    if (tid < ROWS) {
      Partial a = s.wpart[0][tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) merge_partial(a, s.wpart[w][tid]);
      const uint32_t o = part_base + (uint32_t)((c * ROWS + tid) * sizeof(Partial));
      const uint4 v = make_uint4(__float_as_uint(a.m), (uint32_t)a.idx, __float_as_uint(a.s), 0u);
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) st_cluster_v4(mapa(o, rk), v);
    }
    cluster_sync_all();                                                          // #7 partials gathered
    if (tid < ROWS) {
      Partial a = s.part[0][tid];
#pragma unroll
      for (int k = 1; k < CL; ++k) merge_partial(a, s.part[k][tid]);
      s.tok[tid] = a.idx;
      const int gr = row0 + tid;
      if (c == 0 && gr < p.rows) {
        p.tokens[(size_t)gr * p.ld_tok + t + 1] = a.idx;
        if (p.logprob != nullptr) p.logprob[(size_t)gr * p.max_len + t] = -logf(a.s);   // log_softmax of the argmax
        if (a.idx == p.eos && !p.finished[gr]) {
          p.finished[gr] = 1;
          const int cnt = atomicAdd(&p.state->finished_count, 1) + 1;
          if (cnt == p.rows) p.state->steps_executed = t + 1;      // src/inference.py:23-25
        }
      }
    }
    __syncthreads();
    // ---- next input: embedding[token] + pos[t+1] ------------------------------------------------------
    if (t + 1 < p.max_pos) {
      const int r = tid >> 4, c0 = (tid & 15) * 16;
      const int tk = s.tok[r];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)tk * D + c0) + q);
        const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)(t + 1) * D + c0) + q);
        const float4 v = make_float4(e.x + ps.x, e.y + ps.y, e.z + ps.z, e.w + ps.w);
        *reinterpret_cast<float4*>(&s.x32[r][c0 + 4 * q]) = v;
        *reinterpret_cast<uint2*>(&s.xa[r][c0 + 4 * q]) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
      }
    }
    __syncthreads();
  }
  if (blockIdx.x == 0 && tid == 0) p.state->step = t_end;
  cluster_sync_all();      // no CTA exits while a peer may still address its shared memory
}

__global__ void repack_memkv_kernel(const __nv_bfloat16* __restrict__ memkv, int images, int L,
                                    __nv_bfloat16* __restrict__ memk, __nv_bfloat16* __restrict__ memv) {
  // memkv [img*30+s][l*512 + kv*256 + h*32 + d]  ->  mem{k,v} [l][img][h][s][d]   (16-byte chunks)
  const size_t total = (size_t)images * MEM_S * L * 2 * NH * 4;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ch = i & 3, h = (i >> 2) & 7, kv = (i >> 5) & 1;
    const size_t rest = i >> 6;
    const int l = rest % L;
    const size_t row = rest / L;              // img*30 + s
    const int sidx = row % MEM_S;
    const size_t img = row / MEM_S;
    const uint4 v = *reinterpret_cast<const uint4*>(memkv + row * (size_t)L * 512 + l * 512 + kv * 256 + h * 32 + ch * 8);
    __nv_bfloat16* dst = (kv ? memv : memk) + ((((size_t)l * images + img) * NH + h) * MEM_S + sidx) * HD + ch * 8;
    *reinterpret_cast<uint4*>(dst) = v;
  }
}

}  // namespace

int decode_persistent_init() {
  static bool done = false;
  if (done) return 0;
  HM_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<20>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  HM_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  done = true;
  return 0;
}

int decode_persistent_launch(cudaStream_t st, const DecPersistParams& p, int t_begin, int t_end) {
  HM_TRY(decode_persistent_init());
  HM_CHECK(p.tmax <= 256, "decode: max_seq_len %d > 256", p.tmax);
  HM_CHECK(t_begin >= 0 && t_end > t_begin && t_end <= p.tmax, "decode: bad step range [%d,%d)", t_begin, t_end);
  const int clusters = ceil_div(p.rows, ROWS);
  dim3 grid(clusters * CL);
  if (p.tmax <= 160)
    decode_persistent_kernel<20><<<grid, THREADS, sizeof(Smem), st>>>(p, t_begin, t_end);
  else
    decode_persistent_kernel<32><<<grid, THREADS, sizeof(Smem), st>>>(p, t_begin, t_end);
  HM_LAUNCHED();
  return 0;
}

int repack_memkv(cudaStream_t st, const __nv_bfloat16* memkv, int images, int L, __nv_bfloat16* memk,
                 __nv_bfloat16* memv) {
  const size_t total = (size_t)images * MEM_S * L * 2 * NH * 4;
  size_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  repack_memkv_kernel<<<(int)blocks, 256, 0, st>>>(memkv, images, L, memk, memv);
  HM_LAUNCHED();
  return 0;
}

}  // namespace hmocr
