// Persistent cluster decode kernel (sm_100a).
//
// Replaces the whole greedy loop body of /root/reference/src/inference.py:18-25 — i.e. one
// DecoderTransformer.forward step (src/model_swin.py:72-88, 8 x torch TransformerDecoderLayer,
// post-LN) + fc_out + argmax — for a range of steps [t_begin, t_end) in ONE launch.
//
// Mapping: a thread-block CLUSTER of 8 CTAs owns up to 16 sequences end to end.  CTA rank c is
//   * attention head c (self- and cross-attention of the cluster's rows for that head), and
//   * the c-th 1/8 column slice of every projection (q/k/v of head c, 32 columns of out_proj,
//     64 columns of linear1, 32 of linear2, V/8 vocabulary columns of fc_out).
// So each SM streams only 1/8 of the decoder weights per step (1.6 MB instead of 13 MB): the
// weight slices are pre-packed per (layer, CTA) in 16.5 KB chunks and arrive in shared memory
// through a 3-slot ring of cp.async.bulk (TMA) copies, completion on mbarriers.  The activations of
// the cluster's rows are exchanged between the 8 CTAs through distributed shared memory
// (st.shared::cluster into every peer) and ordered by hardware cluster barriers
// (barrier.cluster), 6 per layer + 1 per step; no grid-wide synchronisation, no global-memory
// round trip, no host involvement between steps.  16 rows are exactly one mma.sync m16n8k16 tile;
// the step is bandwidth/latency bound (13-15 MFLOP per token), so tensor-core rate is irrelevant
// here and tcgen05 (M >= 64) would idle most of its rows.
//
// Occupancy is part of the design: the CTA uses 105 KB of shared memory and <= 128 registers so
// that TWO CTAs (of different clusters) share an SM and 33 clusters are co-resident on a B200
// (only 15 fit at one CTA per SM).  B=256 therefore runs as 32 clusters x 8 rows in a single wave,
// and while one CTA of an SM waits on a cluster barrier or on K/V from HBM the other one computes.
//
// Self-attention K/V (bf16) are appended to / streamed from the HBM cache in a single pass
// (online softmax) with 512-byte coalesced warp loads, 8 independent 16-byte loads in flight per
// lane; the cache rows of the NEXT layer are pulled into L2 (cp.async.bulk.prefetch.L2) while the
// current layer's projections run.  The memory K/V of cross-attention are read the same way from a
// [layer][image][head][30][32] repack.
#include "decode_persistent.cuh"

namespace hmocr {
namespace {

constexpr int CL = 8, THREADS = 256, R = 16, NSLOT = 3;
constexpr int D = 256, FF = 512, HD = 32, NH = 8, MEM_S = 30;
constexpr int PD = D + 8, PF = FF + 8;            // padded operand pitches (elements): conflict-free ldmatrix
constexpr float ATT_SCALE = 0.17677669529663687f;
constexpr float LN_EPS = 1e-5f;

struct __align__(16) Partial { float m; int idx; float s; int pad; };

struct Smem {
  alignas(128) uint8_t slot[NSLOT][DP_CHUNK];
  alignas(16) float y32[R][D];             // pre-LayerNorm rows gathered from the 8 column slices
  alignas(16) __nv_bfloat16 xa[R][PD];     // LayerNorm output (full rows), bf16 A operand
  alignas(16) __nv_bfloat16 hf[R][PF];     // relu(linear1) gathered from the 8 slices; its first R*PD
                                           // elements double as the gathered attention context
  alignas(16) float x32s[R][32];           // fp32 residual stream, this CTA's 32-column slice only
  alignas(16) float qs[R][HD];             // this head's scaled query
  Partial part[CL][R];                     // per-CTA argmax / sum-exp partials (gathered)
  Partial wpart[8][R];                     // per-warp partials
  int tok[R];
  alignas(16) float fpar[2][DP_FPC];       // this CTA's bias slices of layer l / l+1
  alignas(16) float fcb[DP_FCB_MAX];       // this CTA's slice of fc_out.bias
  alignas(8) uint64_t full[NSLOT];
  alignas(8) uint64_t fpbar[2];
  alignas(8) uint64_t xbar[4];             // exchange barriers (st.async mode): context, y32, hidden, partials
};
enum { X_CTX = 0, X_Y = 1, X_HF = 2, X_PART = 3 };
static_assert(sizeof(Smem) <= 113 * 1024, "two CTAs must fit one SM");

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
// st.async: a remote (DSMEM) store that also performs complete_tx(bytes) on an mbarrier of the
// destination CTA - the sender needs no fence and no barrier, the receiver waits on its own mbarrier.
__device__ __forceinline__ void st_async_b32(uint32_t addr, uint32_t v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v),
               "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t addr, uint32_t a, uint32_t b, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(addr),
               "r"(a), "r"(b), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}
// wait for a phase whose bytes were written by other CTAs of the cluster (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  long long t0 = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(a), "r"(parity)
        : "memory");
    if (ok) return;
    if (t0 == 0) t0 = clock64();
    else if (clock64() - t0 > 4000000000LL) {
      printf("hmocr: exchange mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
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
__device__ __forceinline__ void axpy8(float (&acc)[8], float p, const uint4 u) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  acc[0] = fmaf(p, a.x, acc[0]); acc[1] = fmaf(p, a.y, acc[1]); acc[2] = fmaf(p, b.x, acc[2]);
  acc[3] = fmaf(p, b.y, acc[3]); acc[4] = fmaf(p, c.x, acc[4]); acc[5] = fmaf(p, c.y, acc[5]);
  acc[6] = fmaf(p, d.x, acc[6]); acc[7] = fmaf(p, d.y, acc[7]);
}

struct OnlineRow {       // per-lane online-softmax state of one query row over this lane's key group
  float m, den, acc[8];
  __device__ __forceinline__ void init() {
    m = -INFINITY; den = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  }
  template <int G>
  __device__ __forceinline__ void update(const float (&s)[G], const uint4 (&v)[G]) {
    float mn = m;
#pragma unroll
    for (int u = 0; u < G; ++u) mn = fmaxf(mn, s[u]);
    if (mn == -INFINITY) return;
    const float corr = __expf(m - mn);
    den *= corr;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= corr;
#pragma unroll
    for (int u = 0; u < G; ++u) {
      const float p = __expf(s[u] - mn);
      den += p;
      axpy8(acc, p, v[u]);
    }
    m = mn;
  }
  // combine the 8 key groups (lanes with the same lane%4) and normalise
  __device__ __forceinline__ void finish(float (&out)[8]) {
    float M = m;
#pragma unroll
    for (int o = 4; o < 32; o <<= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    const float sc = (m == -INFINITY) ? 0.f : __expf(m - M);
    den *= sc;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= sc;
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
};

// NR (1 or 2) query rows against n keys each, one pass.  Lane = (key group jg = lane/4, dim chunk
// dc = lane%4): every load instruction of the warp fetches 8 consecutive 64-byte K (or V) rows =
// 512 contiguous bytes; 8 independent 16-byte loads per lane (K and V of 4/NR key blocks x NR rows)
// are issued before the first use.
template <int NR, int NI>
__device__ __forceinline__ void attend(const float* (&q)[NR], const __nv_bfloat16* (&K)[NR],
                                       const __nv_bfloat16* (&V)[NR], int n, int lane, float (&out)[NR][8]) {
  constexpr int G = 4 / NR;
  const int jg = lane >> 2, dc = lane & 3;
  float qv[NR][8];
  OnlineRow row[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    row[r].init();
#pragma unroll
    for (int e = 0; e < 8; ++e) qv[r][e] = q[r][dc * 8 + e];
  }
#pragma unroll
  for (int i0 = 0; i0 < NI; i0 += G) {
    if (i0 * 8 < n) {                                   // warp-uniform
      uint4 kk[NR][G], vv[NR][G];
#pragma unroll
      for (int u = 0; u < G; ++u) {
        const int j = min(jg + 8 * (i0 + u), n - 1);    // clamped: always a valid row, masked below
        const size_t off = (size_t)j * HD + dc * 8;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          kk[r][u] = __ldcg(reinterpret_cast<const uint4*>(K[r] + off));
          vv[r][u] = __ldcg(reinterpret_cast<const uint4*>(V[r] + off));
        }
      }
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        float sc[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          float a = dot8(qv[r], kk[r][u]);
          a += __shfl_xor_sync(0xffffffffu, a, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2);
          sc[u] = ((jg + 8 * (i0 + u)) < n) ? a : -INFINITY;
        }
        row[r].template update<G>(sc, vv[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r) row[r].finish(out[r]);
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

struct LnRegs { float4 g0, g1, b0, b1; };

// ASYNC = true : exchanges use st.async + per-buffer mbarriers (point-to-point: a CTA proceeds as soon
//                as ITS inputs have arrived; no fence, no cluster-wide barrier)
// ASYNC = false: plain st.shared::cluster + barrier.cluster (first implementation, kept for A/B tests)
template <int NI, bool ASYNC>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(THREADS, 2)
decode_persistent_kernel(const DecPersistParams p, int t_begin, int t_end) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  __nv_bfloat16(*ctxf)[PD] = reinterpret_cast<__nv_bfloat16(*)[PD]>(&s.hf[0][0]);   // aliases hf (see Smem)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = (int)cluster_ctarank();              // column slice == attention head
  const int row0 = (blockIdx.x / CL) * p.rows_per_cluster;
  const int nrows = min(p.rows_per_cluster, p.rows - row0);     // valid rows of this cluster (1..16)
  const int g4 = lane >> 2, t4 = lane & 3;           // mma fragment coordinates
  const int L = p.num_layers;
  const int cps = DP_LAYER_CHUNKS * L + p.fc_chunks; // weight chunks per step
  const int total_chunks = (t_end - t_begin) * cps;
  const int cols_per_cta = p.fc_chunks * 32;

  // The copy issuer is lane 0 of warp 7 (a warp that idles in the 4-tile GEMM phases), so issuing never
  // sits on warp 0's path.  Its cursor is kept incrementally (no integer division on the hot path).
  constexpr int ISSUER = 7 * 32;
  int issued = 0, iss_n = 0, iss_slot = 0;           // meaningful in the issuer thread only
  const uint8_t* iss_src = p.wblob + (size_t)c * DP_LAYER_CTA_BYTES;
  const int layer_chunks_total = DP_LAYER_CHUNKS * L;

  auto issue_next = [&]() {
    uint32_t bytes = DP_CHUNK;
    const uint8_t* src;
    if (iss_n < layer_chunks_total) {
      src = iss_src;
      const int ch = iss_n % DP_LAYER_CHUNKS;         // compile-time divisor
      if (ch >= 8) bytes = DP_CHUNK_F2;               // the two linear2 pieces
      iss_src += bytes;
      if (ch == DP_LAYER_CHUNKS - 1) iss_src += (size_t)(CL - 1) * DP_LAYER_CTA_BYTES;   // next layer's block
    } else {
      src = p.fcblob + ((size_t)c * p.fc_chunks + (iss_n - layer_chunks_total)) * DP_CHUNK;
    }
    uint64_t* bar = &s.full[iss_slot];
    mbar_expect_tx(bar, bytes);
    bulk_g2s(s.slot[iss_slot], src, bytes, bar);
    if (++iss_slot == NSLOT) iss_slot = 0;
    if (++iss_n == cps) { iss_n = 0; iss_src = p.wblob + (size_t)c * DP_LAYER_CTA_BYTES; }
    ++issued;
  };
  // Refill the ring: every chunk < g has been released (its readers passed a block or cluster barrier),
  // so chunks up to g + NSLOT - 1 may be in flight.
  auto refill = [&](int g) {
    if (tid == ISSUER) {
      while (issued < total_chunks && issued < g + NSLOT) issue_next();
    }
  };
  auto wait_chunks = [&](int g, int n) {
    for (int i = 0; i < n; ++i) mbar_wait(&s.full[(g + i) % NSLOT], ((g + i) / NSLOT) & 1);
  };
  auto chunk_ptr = [&](int x) { return reinterpret_cast<const __nv_bfloat16*>(s.slot[x % NSLOT]); };
  // per-layer bias slices of this CTA, double buffered: layer counter gl = step * L + l
  const int total_layers = (t_end - t_begin) * L;
  auto issue_fpar = [&](int gl) {
    if (tid == ISSUER && gl < total_layers) {
      const int l = gl % L;
      mbar_expect_tx(&s.fpbar[gl & 1], DP_FPC * 4);
      bulk_g2s(s.fpar[gl & 1], p.fparams + ((size_t)l * CL + c) * DP_FPC, DP_FPC * 4, &s.fpbar[gl & 1]);
    }
  };
  // LayerNorm gamma/beta of this lane's 8 columns, fetched (L2) a phase ahead of their use
  auto load_ln = [&](int l, int which) {
    const float* g = p.lnparams + ((size_t)l * 6 + 2 * which) * D + lane * 8;
    LnRegs r;
    r.g0 = __ldg(reinterpret_cast<const float4*>(g));
    r.g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
    r.b0 = __ldg(reinterpret_cast<const float4*>(g + D));
    r.b1 = __ldg(reinterpret_cast<const float4*>(g + D + 4));
    return r;
  };

  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) mbar_init(&s.full[i], 1);
    mbar_init(&s.fpbar[0], 1);
    mbar_init(&s.fpbar[1], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < cols_per_cta; i += THREADS) s.fcb[i] = p.fc_bias[c * cols_per_cta + i];

  // x = embedding[token] + pos[t]: full rows as the bf16 A operand, this CTA's 32 columns as fp32 residual
  auto embed_rows = [&](int t, bool from_global) {
    const int r = tid >> 4, c0 = (tid & 15) * 16;
    long long tk = 0;
    const bool ok = r < nrows;
    if (ok) tk = from_global ? p.tokens[(size_t)(row0 + r) * p.ld_tok + t] : (long long)s.tok[r];
    if (tk < 0 || tk >= p.vocab) tk = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) {
        const float4 e = __ldg(reinterpret_cast<const float4*>(p.emb + (size_t)tk * D + c0) + q);
        const float4 ps = __ldg(reinterpret_cast<const float4*>(p.pos + (size_t)t * D + c0) + q);
        v = make_float4(e.x + ps.x, e.y + ps.y, e.z + ps.z, e.w + ps.w);
      }
      *reinterpret_cast<uint2*>(&s.xa[r][c0 + 4 * q]) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
      if ((c0 >> 5) == c) *reinterpret_cast<float4*>(&s.x32s[r][(c0 & 31) + 4 * q]) = v;
    }
  };
  embed_rows(t_begin, true);
  __syncthreads();
  cluster_sync_all();      // every CTA of the cluster is resident and initialised before any DSMEM store

  const uint32_t ctxf_base = smem_u32(&s.hf[0][0]);
  const uint32_t y32_base = smem_u32(&s.y32[0][0]);
  const uint32_t hf_base = smem_u32(&s.hf[0][0]);
  const uint32_t part_base = smem_u32(&s.part[0][0]);

  // LayerNorm of the gathered rows: warp w owns rows w and w+8; lane owns 8 consecutive columns
  auto layer_norm = [&](const LnRegs& ln) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int r = warp + 8 * rr;
      if (r >= nrows) break;                             // warp-uniform
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
      v[0] = v[0] * rstd * ln.g0.x + ln.b0.x; v[1] = v[1] * rstd * ln.g0.y + ln.b0.y;
      v[2] = v[2] * rstd * ln.g0.z + ln.b0.z; v[3] = v[3] * rstd * ln.g0.w + ln.b0.w;
      v[4] = v[4] * rstd * ln.g1.x + ln.b1.x; v[5] = v[5] * rstd * ln.g1.y + ln.b1.y;
      v[6] = v[6] * rstd * ln.g1.z + ln.b1.z; v[7] = v[7] * rstd * ln.g1.w + ln.b1.w;
      if ((lane >> 2) == c) {                            // this CTA's residual slice: columns c*32 .. c*32+31
        *reinterpret_cast<float4*>(&s.x32s[r][(lane & 3) * 8]) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(&s.x32s[r][(lane & 3) * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
      }
      *reinterpret_cast<uint4*>(&s.xa[r][lane * 8]) =
          make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    }
  };
  // query slice of head c (one chunk, 4 n-tiles): qs = (xa Wq^T + b) / sqrt(32)
  auto project_q = [&](const __nv_bfloat16* W, const float* bias) {
    if (warp < 4) {
      float acc[4];
      gemm_tile<D, PD>(&s.xa[0][0], W, warp, lane, acc);
      const int col = warp * 8 + 2 * t4;
      const float b0 = bias[col], b1 = bias[col + 1];
      s.qs[g4][col] = (acc[0] + b0) * ATT_SCALE; s.qs[g4][col + 1] = (acc[1] + b1) * ATT_SCALE;
      s.qs[g4 + 8][col] = (acc[2] + b0) * ATT_SCALE; s.qs[g4 + 8][col + 1] = (acc[3] + b1) * ATT_SCALE;
    }
  };
  // y = acc + bias + residual  ->  y32[r][c*32 + lc ..] of every CTA in the cluster (lc: column in the slice)
  auto scatter_y = [&](const float (&acc)[4], int lc, const float* bias) {
    const float b0 = bias[lc], b1 = bias[lc + 1];
    const int col = c * 32 + lc;
    if (g4 < nrows) {
      const float y0 = acc[0] + b0 + s.x32s[g4][lc], y1 = acc[1] + b1 + s.x32s[g4][lc + 1];
      const uint32_t o0 = y32_base + (g4 * D + col) * 4;
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) st_cluster_v2(mapa(o0, rk), __float_as_uint(y0), __float_as_uint(y1));
    }
    if (g4 + 8 < nrows) {
      const float y2 = acc[2] + b0 + s.x32s[g4 + 8][lc], y3 = acc[3] + b1 + s.x32s[g4 + 8][lc + 1];
      const uint32_t o1 = y32_base + ((g4 + 8) * D + col) * 4;
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) st_cluster_v2(mapa(o1, rk), __float_as_uint(y2), __float_as_uint(y3));
    }
  };
  // out-projection slice (one chunk, 32 columns) over the gathered context
  auto project_out = [&](const __nv_bfloat16* W, const float* bias) {
    if (warp < 4) {
      float acc[4];
      gemm_tile<D, PD>(&ctxf[0][0], W, warp, lane, acc);
      scatter_y(acc, warp * 8 + 2 * t4, bias);
    }
  };
  // attention of rows w (and w+8) for head c; context slice -> ctxf of every CTA
  auto put_ctx = [&](int r, const float (&o)[8]) {
    if (lane < 4) {
      const uint4 v = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      const uint32_t off = ctxf_base + (r * PD + c * HD + lane * 8) * 2;
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) st_cluster_v4(mapa(off, rk), v);
    }
  };
  auto attention = [&](auto kv_of_row, int nkeys) {
    if (warp >= nrows) return;                             // warp-uniform
    if (warp + 8 < nrows) {
      const float* q[2] = {&s.qs[warp][0], &s.qs[warp + 8][0]};
      const __nv_bfloat16 *K[2], *V[2];
      kv_of_row(row0 + warp, K[0], V[0]);
      kv_of_row(row0 + warp + 8, K[1], V[1]);
      float o[2][8];
      attend<2, NI>(q, K, V, nkeys, lane, o);
      put_ctx(warp, o[0]);
      put_ctx(warp + 8, o[1]);
    } else {
      const float* q[1] = {&s.qs[warp][0]};
      const __nv_bfloat16 *K[1], *V[1];
      kv_of_row(row0 + warp, K[0], V[0]);
      float o[1][8];
      attend<1, NI>(q, K, V, nkeys, lane, o);
      put_ctx(warp, o[0]);
    }
  };
  // L2 prefetch of K/V rows ahead of their attention: lane -> (row = lane/2, K or V = lane&1) of warp 6
  auto prefetch_kv = [&](const __nv_bfloat16* kbase, const __nv_bfloat16* vbase, size_t row_stride, int first_row,
                         int count, int bytes) {
    if (warp == 6 && bytes >= 16 && (lane >> 1) < count) {
      const __nv_bfloat16* src = ((lane & 1) ? vbase : kbase) + (size_t)(first_row + (lane >> 1)) * row_stride;
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(reinterpret_cast<uint64_t>(src)), "r"(bytes)
                   : "memory");
    }
  };

  int g = 0;     // weight chunk counter of this launch
  int gl = 0;    // layer counter of this launch (parity of the fpar double buffer)
  const bool tracing = p.trace != nullptr && blockIdx.x == 0 && tid == 0;
  int ti = 0;
#define TR() do { if (tracing && t == p.trace_step && ti < 1024) p.trace[ti++] = clock64(); } while (0)
  issue_fpar(0);
  refill(0);
  for (int t = t_begin; t < t_end; ++t) {
    for (int l = 0; l < L; ++l, ++gl) {
      const size_t cache_l = (size_t)l * p.rows;
      issue_fpar(gl + 1);
      mbar_wait(&s.fpbar[gl & 1], (gl >> 1) & 1);
      const float* fp = s.fpar[gl & 1];
      // ---- self-attention: q, k, v of head c (3 chunks, 12 n-tiles over 8 warps) -------------------
      TR();
      wait_chunks(g, 3);
      TR();
      // cross-attention memory of this layer -> L2 while the self-attention runs
      prefetch_kv(p.memk + (((size_t)l * p.images) * NH + c) * (size_t)MEM_S * HD,
                  p.memv + (((size_t)l * p.images) * NH + c) * (size_t)MEM_S * HD, (size_t)NH * MEM_S * HD,
                  row0 / p.beam, min(nrows, p.images - row0 / p.beam), MEM_S * HD * 2);
      for (int tt = warp; tt < 12; tt += 8) {
        const int part = tt >> 2, nt = tt & 3;              // 0 = q, 1 = k, 2 = v
        float acc[4];
        gemm_tile<D, PD>(&s.xa[0][0], chunk_ptr(g + part), nt, lane, acc);
        const int col = nt * 8 + 2 * t4;
        const float* bias = fp + DPC_BQKV + part * HD;
        const float b0 = bias[col], b1 = bias[col + 1];
        if (part == 0) {
          s.qs[g4][col] = (acc[0] + b0) * ATT_SCALE; s.qs[g4][col + 1] = (acc[1] + b1) * ATT_SCALE;
          s.qs[g4 + 8][col] = (acc[2] + b0) * ATT_SCALE; s.qs[g4 + 8][col + 1] = (acc[3] + b1) * ATT_SCALE;
        } else {
          __nv_bfloat16* cache = (part == 2) ? p.vcache : p.kcache;
          if (g4 < nrows)
            *reinterpret_cast<uint32_t*>(cache + (((cache_l + row0 + g4) * NH + c) * p.tmax + t) * HD + col) =
                pack_bf16(acc[0] + b0, acc[1] + b1);
          if (g4 + 8 < nrows)
            *reinterpret_cast<uint32_t*>(cache + (((cache_l + row0 + g4 + 8) * NH + c) * p.tmax + t) * HD + col) =
                pack_bf16(acc[2] + b0, acc[3] + b1);
        }
      }
      __syncthreads();      // releases the 3 chunks; publishes qs and the appended K/V row to this CTA
      g += 3;
      refill(g);
      TR();
      attention(
          [&](int gr, const __nv_bfloat16*& Kb, const __nv_bfloat16*& Vb) {
            const size_t off = ((cache_l + gr) * NH + c) * (size_t)p.tmax * HD;
            Kb = p.kcache + off; Vb = p.vcache + off;
          },
          t + 1);
      TR();
      cluster_sync_all();                                                        // #1 context gathered
      TR();
      // self-attention cache of the NEXT layer (next step's layer 0 after the last one) -> L2, now that
      // this layer's K/V burst is over
      {
        const int ln = (l + 1 < L) ? l + 1 : 0;
        const int keys = (l + 1 < L) ? t : t + 1;
        if (keys < p.tmax) {
          const size_t base = ((size_t)ln * p.rows * NH + c) * (size_t)p.tmax * HD;
          prefetch_kv(p.kcache + base, p.vcache + base, (size_t)NH * p.tmax * HD, row0, nrows, keys * HD * 2);
        }
      }
      // ---- x = LN1(x + out_proj(ctx)) ---------------------------------------------------------------
      LnRegs ln = load_ln(l, 0);
      wait_chunks(g, 1);
      TR();
      project_out(chunk_ptr(g), fp + DPC_BO);
      TR();
      cluster_sync_all();                                                        // #2 y32 gathered
      TR();
      g += 1;
      refill(g);
      layer_norm(ln);
      __syncthreads();
      TR();
      // ---- cross-attention over the 30 memory tokens ----------------------------------------------
      wait_chunks(g, 1);
      TR();
      project_q(chunk_ptr(g), fp + DPC_BCQ);
      __syncthreads();
      g += 1;
      refill(g);
      TR();
      attention(
          [&](int gr, const __nv_bfloat16*& Kb, const __nv_bfloat16*& Vb) {
            const int img = gr / p.beam;
            const size_t off = (((size_t)l * p.images + img) * NH + c) * (size_t)MEM_S * HD;
            Kb = p.memk + off; Vb = p.memv + off;
          },
          MEM_S);
      TR();
      cluster_sync_all();                                                        // #3
      TR();
      ln = load_ln(l, 1);
      wait_chunks(g, 1);
      TR();
      project_out(chunk_ptr(g), fp + DPC_BCO);
      TR();
      cluster_sync_all();                                                        // #4
      TR();
      g += 1;
      refill(g);
      layer_norm(ln);
      __syncthreads();
      TR();
      // ---- feed-forward: 64 columns of linear1 (+ReLU) per CTA, then 32 columns of linear2 ---------
      wait_chunks(g, 2);
      TR();
      {
        float acc[4];
        gemm_tile<D, PD>(&s.xa[0][0], chunk_ptr(g + (warp >> 2)), warp & 3, lane, acc);
        const int lc = warp * 8 + 2 * t4, col = c * 64 + lc;
        const float b0 = fp[DPC_B1 + lc], b1 = fp[DPC_B1 + lc + 1];
        if (g4 < nrows) {
          const uint32_t h0 = pack_bf16(fmaxf(acc[0] + b0, 0.f), fmaxf(acc[1] + b1, 0.f));
          const uint32_t o0 = hf_base + (g4 * PF + col) * 2;
#pragma unroll
          for (int rk = 0; rk < CL; ++rk) st_cluster_b32(mapa(o0, rk), h0);
        }
        if (g4 + 8 < nrows) {
          const uint32_t h1 = pack_bf16(fmaxf(acc[2] + b0, 0.f), fmaxf(acc[3] + b1, 0.f));
          const uint32_t o1 = hf_base + ((g4 + 8) * PF + col) * 2;
#pragma unroll
          for (int rk = 0; rk < CL; ++rk) st_cluster_b32(mapa(o1, rk), h1);
        }
      }
      TR();
      cluster_sync_all();                                                        // #5 hidden gathered
      TR();
      g += 2;
      refill(g);
      ln = load_ln(l, 2);
      wait_chunks(g, 2);
      TR();
      if (warp < 4) {
        float acc[4];
        gemm_tile<FF, PF>(&s.hf[0][0], chunk_ptr(g + (warp >> 1)), warp & 1, lane, acc);
        scatter_y(acc, warp * 8 + 2 * t4, fp + DPC_B2);
      }
      TR();
      cluster_sync_all();                                                        // #6
      TR();
      g += 2;
      refill(g);
      layer_norm(ln);
      __syncthreads();
      TR();
    }
    // ---- fc_out slice + running (max, argmax, sum-exp); 2 chunks = 64 vocabulary rows per phase -----
    TR();
    Partial pa, pb;        // rows g4 and g4+8 of this thread's columns
    pa.m = pb.m = -INFINITY; pa.idx = pb.idx = 0x7fffffff; pa.s = pb.s = 0.f; pa.pad = pb.pad = 0;
#pragma unroll 1
    for (int ch = 0; ch < p.fc_chunks; ch += 2) {
      wait_chunks(g, 2);
      const int lc = ch * 32 + warp * 8 + 2 * t4;
      const int v0 = c * cols_per_cta + lc;
      const float b0 = s.fcb[lc], b1 = s.fcb[lc + 1];
      float acc[4];
      gemm_tile<D, PD>(&s.xa[0][0], chunk_ptr(g + (warp >> 2)), warp & 3, lane, acc);
      if (v0 < p.vocab) { update_partial(pa, acc[0] + b0, v0); update_partial(pb, acc[2] + b0, v0); }
      if (v0 + 1 < p.vocab) { update_partial(pa, acc[1] + b1, v0 + 1); update_partial(pb, acc[3] + b1, v0 + 1); }
      __syncthreads();
      g += 2;
      refill(g);
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
    if (tid < R) {
      Partial a = s.wpart[0][tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) merge_partial(a, s.wpart[w][tid]);
      const uint32_t o = part_base + (uint32_t)((c * R + tid) * sizeof(Partial));
      const uint4 v = make_uint4(__float_as_uint(a.m), (uint32_t)a.idx, __float_as_uint(a.s), 0u);
#pragma unroll
      for (int rk = 0; rk < CL; ++rk) st_cluster_v4(mapa(o, rk), v);
    }
    TR();
    cluster_sync_all();                                                          // #7 partials gathered
    TR();
    if (tid < R) {
      Partial a = s.part[0][tid];
#pragma unroll
      for (int k = 1; k < CL; ++k) merge_partial(a, s.part[k][tid]);
      s.tok[tid] = a.idx;
      const int gr = row0 + tid;
      if (c == 0 && tid < nrows) {
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
    TR();
    // ---- next input: embedding[token] + pos[t+1] ------------------------------------------------------
    if (t + 1 < p.max_pos) embed_rows(t + 1, false);
    __syncthreads();
    TR();
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

int g_max_clusters = 0;

}  // namespace

int decode_persistent_init() {
  static bool done = false;
  if (done) return 0;
  HM_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<20, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  HM_CUDA(cudaFuncSetAttribute(decode_persistent_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CL * 64);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = sizeof(Smem);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  HM_CUDA(cudaOccupancyMaxActiveClusters(&g_max_clusters, decode_persistent_kernel<20, false>, &cfg));
  HM_CHECK(g_max_clusters >= 1, "device cannot host an 8-CTA decode cluster");
  done = true;
  return 0;
}

int decode_persistent_max_clusters(int* out) {
  HM_TRY(decode_persistent_init());
  *out = g_max_clusters;
  return 0;
}

int decode_persistent_launch(cudaStream_t st, DecPersistParams p, int t_begin, int t_end) {
  HM_TRY(decode_persistent_init());
  HM_CHECK(p.tmax <= 256, "decode: max_seq_len %d > 256", p.tmax);
  HM_CHECK(t_begin >= 0 && t_end > t_begin && t_end <= p.tmax, "decode: bad step range [%d,%d)", t_begin, t_end);
  HM_CHECK(p.fc_chunks % 2 == 0 && p.fc_chunks * 32 <= DP_FCB_MAX, "decode: bad fc_chunks %d", p.fc_chunks);
  // Rows per cluster: spread the rows over as many co-resident clusters as possible (fewer rows per
  // cluster = shorter attention per step, same projection latency) without ever needing a second wave
  // unless the batch exceeds 16 rows x max clusters.
  int rpc = p.rows_per_cluster;
  if (rpc <= 0) {
    rpc = ceil_div(p.rows, g_max_clusters);
    if (rpc > 16) rpc = 16;
  }
  HM_CHECK(rpc >= 1 && rpc <= 16, "decode: rows_per_cluster %d outside [1,16]", rpc);
  p.rows_per_cluster = rpc;
  dim3 grid(ceil_div(p.rows, rpc) * CL);
  if (p.tmax <= 160)
    decode_persistent_kernel<20, false><<<grid, THREADS, sizeof(Smem), st>>>(p, t_begin, t_end);
  else
    decode_persistent_kernel<32, false><<<grid, THREADS, sizeof(Smem), st>>>(p, t_begin, t_end);
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
