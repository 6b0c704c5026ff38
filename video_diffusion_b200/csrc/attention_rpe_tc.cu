// Temporal RPE attention on tensor cores (bf16 mode).
//
// The three RPE einsums of the reference (unet.py:357-378) contract q / k / attn against tables
// R[b][t][s][h][:] that depend on the (batch, frame) pair but NOT on the pixel, so over the pixels of
// one (b, t) they are plain GEMMs.  The host therefore runs them on the tcgen05 GEMM kernel with
// "grouped" weights (one weight block per (b, t) row group), using the block-diagonal-over-heads
// matrices built by rpe_expand_kernel:
//     Sk[(b,t,d)][(h,s)] = q[b,t,d,h,:] . Rk[b,t,s,h,:]          (GEMM, weights Bk)
//     Sq[(b,s,d)][(h,t)] = scale * k[b,s,d,h,:] . Rq[b,s,t,h,:]  (GEMM, weights Bq)
//     out[(b,t,d)][(h,f)] = PV + sum_s P[t][s] * Rv[b,t,s,h,f]   (GEMM, weights Bv, residual PV)
// and attn_temporal_mma_kernel does what remains per (b, pixel, head): scale*q.k^T + the two bias
// terms, mask, fp32 softmax, P.V (mma.sync, sequence length T <= 32), emitting P for the third GEMM.
#include "common.cuh"

namespace vdm {
namespace {

// ---------------------------------------------------------------- block-diagonal weight expansion
// Rows of a group g = (b, t) block: [sub*128 + h*T + j] (sub = g % gpt; zero rows pad each 128 block).
//   which 0/1 (Bk / Bq): out[tg][sub*128 + h*T + j][h'*hd + f] = (h' == h) * mul * R[(g*T + j)][h*hd + f]
//   which 2    (Bv)    : out[tg][h*hd + f][sub*128 + h'*T + j] = (h' == h) * R[(g*T + j)][h*hd + f]
__global__ void __launch_bounds__(256) rpe_expand_kernel(const float* __restrict__ r_q, const float* __restrict__ r_k,
                                                          const float* __restrict__ r_v,
                                                          const float* __restrict__ bias,   // [3][C] (q, k, v) or NULL
                                                          long long r_block_stride, int tgs,   // batching over blocks
                                                          long long qk_block_stride,
                                                          int G, int T, int heads, int hd, int gpt, float scale,
                                                          __nv_bfloat16* __restrict__ bq,
                                                          __nv_bfloat16* __restrict__ bk,
                                                          __nv_bfloat16* __restrict__ bv) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  // one thread = 8 consecutive outputs (one 16-byte store)
  const int C = heads * hd, SW = 128 * gpt;
  const int which = blockIdx.z % 3;   // 0: Bk, 1: Bq, 2: Bv
  const int blk = blockIdx.z / 3;     // attention block (tables / biases / outputs are strided by block)
  const int tg = blockIdx.y;
  const float* R = (which == 0 ? r_k : (which == 1 ? r_q : r_v)) + (size_t)blk * r_block_stride;
  const float* bs = bias ? bias + ((size_t)blk * 3 + (which == 0 ? 1 : (which == 1 ? 0 : 2))) * C : nullptr;
  __nv_bfloat16* out = (which == 0 ? bk : (which == 1 ? bq : bv)) +
                       (which < 2 ? (size_t)blk * qk_block_stride : (size_t)blk * tgs * SW * C) + (size_t)tg * SW * C;
  const float mul = which == 1 ? scale : 1.0f;
  const int nvec = SW * C / 8;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += gridDim.x * blockDim.x) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    if (which < 2) {                      // out[n][c .. c+7]
      const int n = v / (C / 8), c = (v - n * (C / 8)) * 8;
      const int sub = n >> 7, r = n & 127;
      const int g = tg * gpt + sub, h = r / T, j = r - h * T;
      if (g < G && h < heads && c / hd == h) {
        const float* src = R + ((size_t)g * T + j) * C + c;
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
        const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = mul * (x[i] + (bs ? bs[c + i] : 0.f));
      }
    } else {                              // out[c][n .. n+7]
      const int c = v / (SW / 8), n0 = (v - c * (SW / 8)) * 8;
      const int hc = c / hd;
      const float bc = bs ? bs[c] : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + i, sub = n >> 7, r = n & 127;
        const int g = tg * gpt + sub, h = r / T, j = r - h * T;
        if (g < G && h == hc) o[i] = __ldg(R + ((size_t)g * T + j) * C + c) + bc;
      }
    }
    uint4 pk;
    pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
    pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
    reinterpret_cast<uint4*>(out)[v] = pk;
  }
}

// The same operands when the buffers were zeroed once and keep their shape: only the live (block-diagonal) entries
// are rewritten.  grid = (gpt * ceil(C / 64), tile groups, 3 * blocks).
//   Bk / Bq: the CTAs of one tile group stride over its live 16-byte vectors (coalesced 32-byte reads of R)
//   Bv     : CTA = one (b, t) group x 64 channels; the [T][64] slab of R is transposed through shared memory
__global__ void __launch_bounds__(256) rpe_expand_live_kernel(const float* __restrict__ r_q, const float* __restrict__ r_k,
                                                               const float* __restrict__ r_v, const float* __restrict__ bias,
                                                               long long r_block_stride, int tgs,
                                                               long long qk_block_stride, int G, int T, int heads,
                                                               int hd, int gpt, float scale, __nv_bfloat16* __restrict__ bq,
                                                               __nv_bfloat16* __restrict__ bk, __nv_bfloat16* __restrict__ bv) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  __shared__ float tile[32][65];
  const int C = heads * hd, SW = 128 * gpt;
  const int which = blockIdx.z % 3, blk = blockIdx.z / 3, tg = blockIdx.y;
  const float* R = (which == 0 ? r_k : (which == 1 ? r_q : r_v)) + (size_t)blk * r_block_stride;
  const float* bs = bias ? bias + ((size_t)blk * 3 + (which == 0 ? 1 : (which == 1 ? 0 : 2))) * C : nullptr;
  __nv_bfloat16* out = (which == 0 ? bk : (which == 1 ? bq : bv)) +
                       (which < 2 ? (size_t)blk * qk_block_stride : (size_t)blk * tgs * SW * C) + (size_t)tg * SW * C;
  if (which < 2) {
    const float mul = which == 1 ? scale : 1.0f;
    const int hd8 = hd / 8, HT = heads * T;
    const int live = gpt * HT * hd8;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < live; v += gridDim.x * blockDim.x) {
      const int rl = v / hd8, f = (v - rl * hd8) * 8;
      const int sub = rl / HT, hj = rl - sub * HT;
      const int h = hj / T, j = hj - h * T;
      const int g = tg * gpt + sub;
      if (g >= G) continue;
      const int c = h * hd + f;
      const float* src = R + ((size_t)g * T + j) * C + c;
      const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
      const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = mul * (x[i] + (bs ? bs[c + i] : 0.f));
      uint4 pk;
      pk.x = pack_bf16x2(o[0], o[1]); pk.y = pack_bf16x2(o[2], o[3]);
      pk.z = pack_bf16x2(o[4], o[5]); pk.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(out + (size_t)(sub * 128 + hj) * C + c) = pk;
    }
  } else {
    const int nchunk = (C + 63) / 64;
    const int sub = blockIdx.x / nchunk, c0 = (blockIdx.x - sub * nchunk) * 64;
    const int g = tg * gpt + sub;
    if (g >= G) return;
    const int cw = min(64, C - c0);               // channels in this chunk (multiple of 8)
    const int v4 = cw / 4;
    for (int i = threadIdx.x; i < T * v4; i += blockDim.x) {
      const int j = i / v4, q = (i - j * v4) * 4;
      const float4 a = __ldg(reinterpret_cast<const float4*>(R + ((size_t)g * T + j) * C + c0 + q));
      tile[j][q] = a.x; tile[j][q + 1] = a.y; tile[j][q + 2] = a.z; tile[j][q + 3] = a.w;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cw * T; i += blockDim.x) {
      const int cl = i / T, j = i - cl * T;
      const int c = c0 + cl;
      const float val = tile[j][cl] + (bs ? bs[c] : 0.f);
      out[(size_t)c * SW + sub * 128 + (c / hd) * T + j] = __float2bfloat16_rn(val);
    }
  }
}

// ---------------------------------------------------------------- per-(b, pixel, head) attention
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* smem) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp16(void* smem, const void* gmem, bool valid) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a), "l"(gmem), "r"(bytes) : "memory");
}

// CTA = 4 warps = 2 problems (b, d, h) x 2 warps of 16 query frames each (T <= 32).
template <int HD>
__global__ void __launch_bounds__(128) attn_temporal_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                const float* __restrict__ sk,
                                                                const float* __restrict__ sq,
                                                                const float* __restrict__ mask, int pad_interact,
                                                                int n_prob, int T, int D, int heads, int gpt,
                                                                __nv_bfloat16* __restrict__ pm,
                                                                float* __restrict__ pv) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  constexpr int LDS = HD + 8, CH = HD / 8, ROWS = 32;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int prob_local = warp >> 1, w2 = warp & 1;
  const int pid = blockIdx.x * 2 + prob_local;
  const bool live = pid < n_prob;
  const int C = heads * HD, SW = 128 * gpt;
  const int h = live ? pid % heads : 0;
  const int d = live ? (pid / heads) % D : 0;
  const int b = live ? pid / (heads * D) : 0;
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw) + prob_local * 3 * ROWS * LDS;
  __nv_bfloat16* sK = sQ + ROWS * LDS;
  __nv_bfloat16* sV = sK + ROWS * LDS;
  const float scale = rsqrtf((float)HD);

  // stage q, k, v rows t < T of this problem (64 threads per problem); rows >= T are zero
  {
    const int t64 = tid & 63;
    for (int idx = t64; idx < ROWS * CH; idx += 64) {
      const int r = idx / CH, c = idx - r * CH;
      const bool ok = live && r < T;
      const __nv_bfloat16* src = qkv + ((size_t)(b * T + (ok ? r : 0)) * D + d) * (3 * C) + h * HD + c * 8;
      cp16(sQ + r * LDS + c * 8, src, ok);
      cp16(sK + r * LDS + c * 8, src + C, ok);
      cp16(sV + r * LDS + c * 8, src + 2 * C, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  // The RPE bias terms and the mask do not depend on the staged tiles: fetch them while the cp.async copies are in
  // flight, so the kernel waits for one memory round trip instead of two.
  const int r0 = w2 * 16 + (lane >> 2);
  float bias_k[2][8], bias_q[2][8], allow[2][8];
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8;
    const bool row_ok = live && r < T;
    const int gr = b * T + (row_ok ? r : 0);
    const float* sk_row = sk + ((size_t)gr * D + d) * SW + (gr % gpt) * 128 + h * T;
    const float m_r = row_ok ? mask[b * T + r] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + 2 * (lane & 3) + e;
        float bk = 0.f, bq = 0.f, al = 0.f;
        if (row_ok && col < T) {
          const int gs = b * T + col;
          bq = sq[((size_t)gs * D + d) * SW + (gs % gpt) * 128 + h * T + r];
          bk = sk_row[col];
          const float m_s = mask[b * T + col];
          al = m_r * m_s;
          if (pad_interact) al += (1.f - m_r) * (1.f - m_s);
          else if (col == r) al = 1.f;
        }
        bias_k[half][nt * 2 + e] = bk;
        bias_q[half][nt * 2 + e] = bq;
        allow[half][nt * 2 + e] = al;
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (!live) return;

  // S = Q K^T : this warp's 16 query rows x 32 keys
  float s[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk) {
    uint32_t qa[4];
    ldsm_x4(qa, sQ + (w2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8);
#pragma unroll
    for (int nt = 0; nt < 4; nt += 2) {
      uint32_t kb[4];
      ldsm_x4(kb, sK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LDS + kk * 16 + ((lane >> 3) & 1) * 8);
      mma16816(s[nt], qa, kb[0], kb[1]);
      mma16816(s[nt + 1], qa, kb[2], kb[3]);
    }
  }
  // logits = scale*(q.k + Sk) + Sq, mask, softmax over the key axis (rows r0 and r0 + 8 of this lane)
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8;
    const bool row_ok = r < T;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = nt * 8 + 2 * (lane & 3) + e;
        float v = -INFINITY;
        if (row_ok && col < T && allow[half][nt * 2 + e] != 0.f)
          v = scale * (s[nt][half * 2 + e] + bias_k[half][nt * 2 + e]) + bias_q[half][nt * 2 + e];
        s[nt][half * 2 + e] = v;
        mx[half] = fmaxf(mx[half], v);
      }
    }
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    mx[half] = fmaxf(mx[half], __shfl_xor_sync(0xffffffffu, mx[half], 1));
    mx[half] = fmaxf(mx[half], __shfl_xor_sync(0xffffffffu, mx[half], 2));
    if (mx[half] == -INFINITY) mx[half] = 0.f;      // padded query rows: all keys masked
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float pexp = __expf(s[nt][half * 2 + e] - mx[half]);
        s[nt][half * 2 + e] = pexp;
        sum[half] += pexp;
      }
    sum[half] += __shfl_xor_sync(0xffffffffu, sum[half], 1);
    sum[half] += __shfl_xor_sync(0xffffffffu, sum[half], 2);
    const float inv = sum[half] > 0.f ? 1.f / sum[half] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][half * 2] *= inv;
      s[nt][half * 2 + 1] *= inv;
    }
  }
  // P (normalised) -> global for the Rv GEMM
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8;
    if (r < T) {
      const int gr = b * T + r;
      __nv_bfloat16* prow = pm + ((size_t)gr * D + d) * SW + (gr % gpt) * 128 + h * T;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = nt * 8 + 2 * (lane & 3) + e;
          if (col < T) prow[col] = __float2bfloat16_rn(s[nt][half * 2 + e]);
        }
    }
  }
  // O = P V
  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 2; ++kk) {
    uint32_t pa[4];
    pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
    pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
    pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int nt = 0; nt < HD / 8; nt += 2) {
      uint32_t vb[4];
      ldsm_x4_t(vb, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + nt * 8 + (lane >> 4) * 8);
      mma16816(o[nt], pa, vb[0], vb[1]);
      mma16816(o[nt + 1], pa, vb[2], vb[3]);
    }
  }
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = r0 + half * 8;
    if (r < T) {
      float* orow = pv + ((size_t)(b * T + r) * D + d) * C + h * HD + 2 * (lane & 3);
#pragma unroll
      for (int nt = 0; nt < HD / 8; ++nt)
        *reinterpret_cast<float2*>(orow + nt * 8) = make_float2(o[nt][half * 2], o[nt][half * 2 + 1]);
    }
  }
}

template <int HD>
int launch_attn(const void* qkv, const float* sk, const float* sq, const float* mask, int pad, int B, int T, int D,
                int heads, int gpt, void* pm, float* pv, cudaStream_t stream) {
  const size_t smem = (size_t)2 * 3 * 32 * (HD + 8) * sizeof(__nv_bfloat16);
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_mma_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attn_temporal_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  const int n_prob = B * D * heads;
  launch_kernel(attn_temporal_mma_kernel<HD>, (n_prob + 1) / 2, 128, smem, (cudaStream_t)stream, 1, 
      (const __nv_bfloat16*)qkv, sk, sq, mask, pad, n_prob, T, D, heads, gpt, (__nv_bfloat16*)pm, pv);
  VDM_AFTER_LAUNCH("attn_temporal_tc");
  return 0;
}

}  // namespace
}  // namespace vdm

using namespace vdm;

extern "C" int vdm_rpe_expand(const float* r_q, const float* r_k, const float* r_v, const float* bias,
                              int32_t n_blocks, int64_t r_block_stride, int64_t qk_block_stride, int32_t B,
                              int32_t T, int32_t heads,
                              int32_t hd, int32_t groups_per_tile, int32_t zero_fill, void* bq, void* bk,
                              void* bv, vdm_stream_t stream) {
  VDM_REQUIRE(r_q && r_k && r_v && bq && bk && bv, "rpe_expand: NULL pointer");
  VDM_REQUIRE(heads * T <= 128 && groups_per_tile >= 1, "rpe_expand: heads*T = %d must be <= 128", heads * T);
  const int G = B * T, gpt = groups_per_tile;
  const int tgs = (G + gpt - 1) / gpt;
  VDM_REQUIRE(hd % 8 == 0, "rpe_expand: head_dim must be a multiple of 8");
  const int nvec = 128 * gpt * heads * hd / 8;
  VDM_REQUIRE(n_blocks >= 1, "rpe_expand: n_blocks must be >= 1");
  if (qk_block_stride == 0) qk_block_stride = (int64_t)tgs * 128 * gpt * heads * hd;
  if (!zero_fill) {
    VDM_REQUIRE(T <= 32, "rpe_expand: T=%d must be <= 32", T);
    dim3 grid(gpt * ((heads * hd + 63) / 64), tgs, 3 * n_blocks);
    launch_kernel(rpe_expand_live_kernel, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, r_q, r_k, r_v, bias, r_block_stride, tgs, qk_block_stride, G, T, heads, hd,
                                                                   gpt, 1.0f / sqrtf((float)hd), (__nv_bfloat16*)bq,
                                                                   (__nv_bfloat16*)bk, (__nv_bfloat16*)bv);
    VDM_AFTER_LAUNCH("rpe_expand");
    return 0;
  }
  dim3 grid(std::min((nvec + 255) / 256, 32), tgs, 3 * n_blocks);
  launch_kernel(rpe_expand_kernel, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, r_q, r_k, r_v, bias, r_block_stride, tgs, qk_block_stride, G, T, heads, hd, gpt, 1.0f / sqrtf((float)hd),
                                                            (__nv_bfloat16*)bq, (__nv_bfloat16*)bk, (__nv_bfloat16*)bv);
  VDM_AFTER_LAUNCH("rpe_expand");
  return 0;
}

extern "C" int vdm_attn_temporal_tc(const void* qkv, const float* sk, const float* sq, const float* mask,
                                    int32_t allow_pad_interactions, int32_t B, int32_t T, int32_t HW, int32_t heads,
                                    int32_t hd, int32_t groups_per_tile, void* pm, float* pv, vdm_stream_t stream) {
  VDM_REQUIRE(qkv && sk && sq && mask && pm && pv, "attn_temporal_tc: NULL pointer");
  VDM_REQUIRE(T >= 1 && T <= 32 && heads * T <= 128, "attn_temporal_tc: T=%d, heads=%d unsupported", T, heads);
  switch (hd) {
    case 32: return launch_attn<32>(qkv, sk, sq, mask, allow_pad_interactions, B, T, HW, heads, groups_per_tile, pm, pv, (cudaStream_t)stream);
    case 64: return launch_attn<64>(qkv, sk, sq, mask, allow_pad_interactions, B, T, HW, heads, groups_per_tile, pm, pv, (cudaStream_t)stream);
    case 96: return launch_attn<96>(qkv, sk, sq, mask, allow_pad_interactions, B, T, HW, heads, groups_per_tile, pm, pv, (cudaStream_t)stream);
    case 128: return launch_attn<128>(qkv, sk, sq, mask, allow_pad_interactions, B, T, HW, heads, groups_per_tile, pm, pv, (cudaStream_t)stream);
  }
  set_error("attn_temporal_tc: head_dim=%d not in {32, 64, 96, 128}", hd);
  return -1;
}
