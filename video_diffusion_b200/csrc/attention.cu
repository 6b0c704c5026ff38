// Attention cores of the FactorizedAttentionBlock (unet.py:236-268, 471-540).
//
// temporal: sequence = frames (T <= 32), one problem per (batch, pixel, head).  The RPE terms
//   logits[t][s] = scale*q_t.(k_s + Rk[t][s]) + scale*k_s.Rq[s][t];   out_t = sum_s P[t][s]*(v_s + Rv[t][s])
// are fused in: a CTA owns one (batch, head) and PIX pixels, keeps their K/V in shared memory and
// streams the per-query slices of the three R tables through shared memory once for all pixels.
//
// spatial (fp32 reference-accuracy kernel): sequence = pixels (L <= 256), plain softmax(QK^T)V;
// K/V of one (image, head) in shared memory, one warp per query.  The bf16 tensor-core version
// lives in attention_tc.cu.
#include "common.cuh"

namespace vdm {
namespace {

constexpr int PIX = 8;  // pixels (= warps) per CTA in the temporal kernel

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// grid: (ceil(HW/PIX), heads, B); block: PIX warps.
// smem floats: K[PIX][T][hs] V[PIX][T][hs] Rk[T][hs] Rq[T][hs] Rv[T][hs] q[PIX][hs] P[PIX][32], hs = hd+4
template <typename OutT>
__global__ void __launch_bounds__(PIX * 32) attn_temporal_kernel(const float* __restrict__ qkv,
                                                                 const float* __restrict__ r_q,
                                                                 const float* __restrict__ r_k,
                                                                 const float* __restrict__ r_v,
                                                                 const float* __restrict__ mask, int pad_interact,
                                                                 int T, int HW, int heads, int hd,
                                                                 OutT* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) float sm[];
  const int hs = hd + 4;
  const int C = heads * hd;
  float* Ks = sm;
  float* Vs = Ks + PIX * T * hs;
  float* Rk = Vs + PIX * T * hs;
  float* Rq = Rk + T * hs;
  float* Rv = Rq + T * hs;
  float* Qs = Rv + T * hs;
  float* Ps = Qs + PIX * hs;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int pix = blockIdx.x * PIX + warp;
  const bool active = pix < HW;
  const float scale = rsqrtf((float)hd);
  const int hd4 = hd >> 2;

  // K, V of this warp's pixel: rows t = 0..T-1
  if (active) {
    for (int idx = lane; idx < T * hd4; idx += 32) {
      const int t = idx / hd4, f = (idx - t * hd4) * 4;
      const float* row = qkv + ((size_t)(b * T + t) * HW + pix) * (3 * C) + h * hd + f;
      *reinterpret_cast<float4*>(Ks + (warp * T + t) * hs + f) = __ldg(reinterpret_cast<const float4*>(row + C));
      *reinterpret_cast<float4*>(Vs + (warp * T + t) * hs + f) = __ldg(reinterpret_cast<const float4*>(row + 2 * C));
    }
  }
  const float m_self_lane = (lane < T) ? mask[b * T + lane] : 0.f;

  for (int t = 0; t < T; ++t) {
    __syncthreads();  // previous iteration's readers of Rk/Rq/Rv are done
    // R slices for query frame t: Rk[t][s][:], Rq[s][t][:], Rv[t][s][:]  (rows of [B*T*T][C])
    for (int idx = threadIdx.x; idx < T * hd4; idx += blockDim.x) {
      const int s = idx / hd4, f = (idx - s * hd4) * 4;
      const size_t row_ts = ((size_t)(b * T + t) * T + s) * C + h * hd + f;
      const size_t row_st = ((size_t)(b * T + s) * T + t) * C + h * hd + f;
      *reinterpret_cast<float4*>(Rk + s * hs + f) = __ldg(reinterpret_cast<const float4*>(r_k + row_ts));
      *reinterpret_cast<float4*>(Rq + s * hs + f) = __ldg(reinterpret_cast<const float4*>(r_q + row_st));
      *reinterpret_cast<float4*>(Rv + s * hs + f) = __ldg(reinterpret_cast<const float4*>(r_v + row_ts));
    }
    if (active) {
      const float* qrow = qkv + ((size_t)(b * T + t) * HW + pix) * (3 * C) + h * hd;
      for (int f = lane * 4; f < hd; f += 128) {
        float4 q = __ldg(reinterpret_cast<const float4*>(qrow + f));
        q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;   // q *= self.scale (unet.py:487)
        *reinterpret_cast<float4*>(Qs + warp * hs + f) = q;
      }
    }
    __syncthreads();
    if (!active) continue;
    // ---- logits: lane = key frame s
    float logit = -INFINITY;
    if (lane < T) {
      const float* kr = Ks + (warp * T + lane) * hs;
      const float* rk = Rk + lane * hs;
      const float* rq = Rq + lane * hs;
      const float* qs = Qs + warp * hs;
      float a0 = 0.f, a1 = 0.f;
      for (int f = 0; f < hd; f += 4) {
        const float4 q = *reinterpret_cast<const float4*>(qs + f);
        const float4 k = *reinterpret_cast<const float4*>(kr + f);
        const float4 x = *reinterpret_cast<const float4*>(rk + f);
        const float4 y = *reinterpret_cast<const float4*>(rq + f);
        a0 = fmaf(q.x, k.x + x.x, a0); a0 = fmaf(q.y, k.y + x.y, a0);
        a0 = fmaf(q.z, k.z + x.z, a0); a0 = fmaf(q.w, k.w + x.w, a0);
        a1 = fmaf(k.x, y.x, a1); a1 = fmaf(k.y, y.y, a1); a1 = fmaf(k.z, y.z, a1); a1 = fmaf(k.w, y.w, a1);
      }
      logit = a0 + a1 * scale;
      // allowed = m_t*m_s (+ (1-m_t)(1-m_s) | diagonal)   (unet.py:511-524)
      const float mt = mask[b * T + t], ms = m_self_lane;
      float allowed = mt * ms;
      if (pad_interact) allowed += (1.f - mt) * (1.f - ms);
      else if (lane == t) allowed = 1.f;
      if (allowed == 0.f) logit = -INFINITY;
    }
    const float mx = warp_max(logit);
    const float e = (lane < T) ? __expf(logit - mx) : 0.f;
    const float denom = warp_sum(e);
    Ps[warp * 32 + lane] = e / denom;
    __syncwarp();
    // ---- output: lane = 4 head-dim elements
    OutT* orow = out + ((size_t)(b * T + t) * HW + pix) * C + h * hd;
    for (int f = lane * 4; f < hd; f += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < T; ++s) {
        const float pw = Ps[warp * 32 + s];
        const float4 v = *reinterpret_cast<const float4*>(Vs + (warp * T + s) * hs + f);
        const float4 r = *reinterpret_cast<const float4*>(Rv + s * hs + f);
        acc.x = fmaf(pw, v.x + r.x, acc.x); acc.y = fmaf(pw, v.y + r.y, acc.y);
        acc.z = fmaf(pw, v.z + r.z, acc.z); acc.w = fmaf(pw, v.w + r.w, acc.w);
      }
      if constexpr (sizeof(OutT) == 2) {
        uint2 pk;
        pk.x = pack_bf16x2(acc.x, acc.y);
        pk.y = pack_bf16x2(acc.z, acc.w);
        *reinterpret_cast<uint2*>(orow + f) = pk;
      } else {
        *reinterpret_cast<float4*>(orow + f) = acc;
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------ spatial, fp32
// grid: (ceil(L / QPB), heads, n_img); 8 warps; K, V [L][hs] in smem; one query per warp at a time.
constexpr int QPB = 64;
template <typename OutT>
__global__ void __launch_bounds__(256) attn_spatial_f32_kernel(const float* __restrict__ qkv, int L, int heads,
                                                                int hd, OutT* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) float sm[];
  const int hs = hd + 4, C = heads * hd, hd4 = hd >> 2;
  float* Ks = sm;
  float* Vs = Ks + L * hs;
  float* Qs = Vs + L * hs;       // [8][hs]
  float* Ps = Qs + 8 * hs;       // [8][L]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, n = blockIdx.z;
  const float scale = rsqrtf((float)hd);
  for (int idx = threadIdx.x; idx < L * hd4; idx += blockDim.x) {
    const int s = idx / hd4, f = (idx - s * hd4) * 4;
    const float* row = qkv + ((size_t)n * L + s) * (3 * C) + h * hd + f;
    *reinterpret_cast<float4*>(Ks + s * hs + f) = __ldg(reinterpret_cast<const float4*>(row + C));
    *reinterpret_cast<float4*>(Vs + s * hs + f) = __ldg(reinterpret_cast<const float4*>(row + 2 * C));
  }
  __syncthreads();
  const int q_end = min(L, (int)(blockIdx.x + 1) * QPB);
  for (int qi = blockIdx.x * QPB + warp; qi < q_end; qi += 8) {
    const float* qrow = qkv + ((size_t)n * L + qi) * (3 * C) + h * hd;
    for (int f = lane * 4; f < hd; f += 128) {
      float4 q = __ldg(reinterpret_cast<const float4*>(qrow + f));
      q.x *= scale; q.y *= scale; q.z *= scale; q.w *= scale;
      *reinterpret_cast<float4*>(Qs + warp * hs + f) = q;
    }
    __syncwarp();
    float mx = -INFINITY;
    for (int s = lane; s < L; s += 32) {
      const float* kr = Ks + s * hs;
      float a = 0.f;
      for (int f = 0; f < hd; f += 4) {
        const float4 q = *reinterpret_cast<const float4*>(Qs + warp * hs + f);
        const float4 k = *reinterpret_cast<const float4*>(kr + f);
        a = fmaf(q.x, k.x, a); a = fmaf(q.y, k.y, a); a = fmaf(q.z, k.z, a); a = fmaf(q.w, k.w, a);
      }
      Ps[warp * L + s] = a;
      mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int s = lane; s < L; s += 32) {
      const float e = __expf(Ps[warp * L + s] - mx);
      Ps[warp * L + s] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    __syncwarp();
    OutT* orow = out + ((size_t)n * L + qi) * C + h * hd;
    for (int f = lane * 4; f < hd; f += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < L; ++s) {
        const float pw = Ps[warp * L + s];
        const float4 v = *reinterpret_cast<const float4*>(Vs + s * hs + f);
        acc.x = fmaf(pw, v.x, acc.x); acc.y = fmaf(pw, v.y, acc.y);
        acc.z = fmaf(pw, v.z, acc.z); acc.w = fmaf(pw, v.w, acc.w);
      }
      acc.x *= inv; acc.y *= inv; acc.z *= inv; acc.w *= inv;
      if constexpr (sizeof(OutT) == 2) {
        uint2 pk;
        pk.x = pack_bf16x2(acc.x, acc.y);
        pk.y = pack_bf16x2(acc.z, acc.w);
        *reinterpret_cast<uint2*>(orow + f) = pk;
      } else {
        *reinterpret_cast<float4*>(orow + f) = acc;
      }
    }
    __syncwarp();
  }
}

template <typename K>
int set_smem(K kernel, size_t bytes, const char* name) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("%s: cannot reserve %zu bytes of shared memory: %s", name, bytes, cudaGetErrorString(e));
    return (int)e;
  }
  return 0;
}

}  // namespace

int attn_spatial_tc(const void* qkv, int n_img, int L, int heads, int hd, void* out_a, int out_dtype,
                    cudaStream_t stream);
bool attn_spatial_sm100_supported(int L, int heads, int hd);
int attn_spatial_sm100(const void* qkv, int n_img, int L, int heads, int hd, void* out_a, cudaStream_t stream);

}  // namespace vdm

using namespace vdm;

extern "C" int vdm_attn_temporal(const float* qkv, const float* r_q, const float* r_k, const float* r_v,
                                 const float* mask, int32_t allow_pad_interactions, int32_t B, int32_t T, int32_t HW,
                                 int32_t heads, int32_t hd, void* out_a, int32_t out_dtype, vdm_stream_t stream) {
  VDM_REQUIRE(qkv && r_q && r_k && r_v && mask && out_a, "attn_temporal: NULL pointer");
  VDM_REQUIRE(T >= 1 && T <= 32, "attn_temporal: T=%d must be in [1,32]", T);
  VDM_REQUIRE(hd % 4 == 0 && hd <= 128, "attn_temporal: head_dim=%d must be a multiple of 4, <= 128", hd);
  const int hs = hd + 4;
  const size_t smem = sizeof(float) * ((size_t)2 * PIX * T * hs + 3 * (size_t)T * hs + PIX * hs + PIX * 32);
  VDM_REQUIRE(smem <= 227 * 1024, "attn_temporal: T=%d, head_dim=%d needs %zu B of shared memory", T, hd, smem);
  dim3 grid((HW + PIX - 1) / PIX, heads, B);
  int rc;
  if (out_dtype == VDM_BF16) {
    if ((rc = set_smem(attn_temporal_kernel<__nv_bfloat16>, smem, "attn_temporal"))) return rc;
    launch_kernel(attn_temporal_kernel<__nv_bfloat16>, grid, PIX * 32, smem, (cudaStream_t)(cudaStream_t)stream, 1, 
        qkv, r_q, r_k, r_v, mask, allow_pad_interactions, T, HW, heads, hd, (__nv_bfloat16*)out_a);
  } else {
    if ((rc = set_smem(attn_temporal_kernel<float>, smem, "attn_temporal"))) return rc;
    launch_kernel(attn_temporal_kernel<float>, grid, PIX * 32, smem, (cudaStream_t)(cudaStream_t)stream, 1, 
        qkv, r_q, r_k, r_v, mask, allow_pad_interactions, T, HW, heads, hd, (float*)out_a);
  }
  VDM_AFTER_LAUNCH("attn_temporal");
  return 0;
}

extern "C" int vdm_attn_spatial(const void* qkv, int32_t qkv_dtype, int32_t n_img, int32_t L, int32_t heads, int32_t hd,
                                void* out_a, int32_t out_dtype, vdm_stream_t stream) {
  VDM_REQUIRE(qkv && out_a, "attn_spatial: NULL pointer");
  VDM_REQUIRE(hd % 4 == 0 && hd <= 128, "attn_spatial: head_dim=%d must be a multiple of 4, <= 128", hd);
  if (qkv_dtype == VDM_BF16) {
    // tcgen05 / TMEM kernel where the shape fits its envelope, else the mma.sync flash kernel
    if (out_dtype == VDM_BF16 && attn_spatial_sm100_supported(L, heads, hd) &&
        (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out_a) & 15) == 0)
      return attn_spatial_sm100(qkv, n_img, L, heads, hd, out_a, (cudaStream_t)stream);
    return attn_spatial_tc(qkv, n_img, L, heads, hd, out_a, out_dtype, (cudaStream_t)stream);
  }
  const int hs = hd + 4;
  const size_t smem = sizeof(float) * ((size_t)2 * L * hs + 8 * hs + 8 * (size_t)L);
  VDM_REQUIRE(smem <= 227 * 1024, "attn_spatial: L=%d, head_dim=%d needs %zu B of shared memory", L, hd, smem);
  dim3 grid((L + QPB - 1) / QPB, heads, n_img);
  int rc;
  if (out_dtype == VDM_BF16) {
    if ((rc = set_smem(attn_spatial_f32_kernel<__nv_bfloat16>, smem, "attn_spatial"))) return rc;
    launch_kernel(attn_spatial_f32_kernel<__nv_bfloat16>, grid, 256, smem, (cudaStream_t)(cudaStream_t)stream, 1, (const float*)qkv, L, heads, hd,
                                                                                    (__nv_bfloat16*)out_a);
  } else {
    if ((rc = set_smem(attn_spatial_f32_kernel<float>, smem, "attn_spatial"))) return rc;
    launch_kernel(attn_spatial_f32_kernel<float>, grid, 256, smem, (cudaStream_t)(cudaStream_t)stream, 1, (const float*)qkv, L, heads, hd,
                                                                             (float*)out_a);
  }
  VDM_AFTER_LAUNCH("attn_spatial");
  return 0;
}

// ---- head-averaged attention maps (logging; unet.py:464-468) ------------------------------------------------
// out[g][i][j] = | mean_h softmax_j( logits_h[i][j] ) | for sequence g = (outer, inner).  One warp per (g, i); lane l
// owns keys j = l, l + 32, ...  Recomputes the logits from q, k (and the RPE tables / frame mask of the temporal
// attention) -- only launched when the caller asks for return_attn_weights=True.
namespace vdm {
namespace {

constexpr int AW_WARPS = 4, AW_MAXK = 8;   // L <= 32 * AW_MAXK

__device__ __forceinline__ float aw_ld(const float* p) { return *p; }
__device__ __forceinline__ float aw_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename InT>
__global__ void __launch_bounds__(AW_WARPS * 32) attn_weights_mean_kernel(
    const InT* __restrict__ qkv, long long n_seq, int n_inner, long long outer_stride, long long inner_stride,
    long long seq_stride, int L, int heads, int hd, const float* __restrict__ r_q, const float* __restrict__ r_k,
    const float* __restrict__ mask, int pad_interact, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  __shared__ float qs[AW_WARPS][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wi = (long long)blockIdx.x * AW_WARPS + warp;
  if (wi >= n_seq * L) return;
  const long long g = wi / L;
  const int i = (int)(wi - g * L);
  const long long outer = g / n_inner, inner = g - outer * n_inner;
  const InT* base = qkv + outer * outer_stride + inner * inner_stride;
  const int C = heads * hd;
  const float scale = rsqrtf((float)hd);
  const float mi = mask ? mask[outer * L + i] : 1.f;
  float acc[AW_MAXK];
#pragma unroll
  for (int u = 0; u < AW_MAXK; ++u) acc[u] = 0.f;
  for (int h = 0; h < heads; ++h) {
    __syncwarp();
    for (int d = lane; d < hd; d += 32) qs[warp][d] = aw_ld(base + (long long)i * seq_stride + h * hd + d) * scale;
    __syncwarp();
    float lg[AW_MAXK];
    float mx = -INFINITY;
#pragma unroll
    for (int u = 0; u < AW_MAXK; ++u) {
      const int j = lane + 32 * u;
      lg[u] = -INFINITY;
      if (j < L) {
        const InT* kj = base + (long long)j * seq_stride + C + h * hd;
        float dot = 0.f;
        if (r_k) {
          // q'.k_j + q'.Rk[b][i][j][h] + (scale k_j).Rq[b][j][i][h]   (unet.py:489-509)
          const float* rk = r_k + ((outer * L + i) * L + j) * C + h * hd;
          const float* rq = r_q + ((outer * L + j) * L + i) * C + h * hd;
          float dq = 0.f;
          for (int d = 0; d < hd; ++d) {
            const float kv = aw_ld(kj + d);
            dot = fmaf(qs[warp][d], kv + rk[d], dot);
            dq = fmaf(kv, rq[d], dq);
          }
          dot = fmaf(dq, scale, dot);
        } else {
          for (int d = 0; d < hd; ++d) dot = fmaf(qs[warp][d], aw_ld(kj + d), dot);
        }
        bool ok = true;
        if (mask) {
          const float mj = mask[outer * L + j];
          float allowed = mi * mj;
          if (pad_interact) allowed += (1.f - mi) * (1.f - mj);
          else if (j == i) allowed = 1.f;
          ok = allowed != 0.f;
        }
        lg[u] = ok ? dot : -INFINITY;
        mx = fmaxf(mx, lg[u]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < AW_MAXK; ++u) {
      lg[u] = (lane + 32 * u < L) ? __expf(lg[u] - mx) : 0.f;
      sum += lg[u];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int u = 0; u < AW_MAXK; ++u) acc[u] = fmaf(lg[u], inv, acc[u]);
  }
  float* o = out + wi * L;
#pragma unroll
  for (int u = 0; u < AW_MAXK; ++u)
    if (lane + 32 * u < L) o[lane + 32 * u] = fabsf(acc[u] / (float)heads);
}

}  // namespace
}  // namespace vdm

extern "C" int vdm_attn_weights_mean(const void* qkv, int32_t qkv_dtype, int64_t n_outer, int64_t n_inner,
                                     int64_t outer_stride, int64_t inner_stride, int64_t seq_stride, int32_t L,
                                     int32_t heads, int32_t hd, const float* r_q, const float* r_k, const float* mask,
                                     int32_t allow_pad_interactions, float* out, vdm_stream_t stream) {
  VDM_REQUIRE(qkv && out && n_outer > 0 && n_inner > 0, "attn_weights_mean: bad arguments");
  VDM_REQUIRE(L >= 1 && L <= 32 * AW_MAXK, "attn_weights_mean: sequence length %d must be in [1,%d]", L, 32 * AW_MAXK);
  VDM_REQUIRE(hd >= 1 && hd <= 128, "attn_weights_mean: head_dim=%d must be <= 128", hd);
  VDM_REQUIRE((r_q == nullptr) == (r_k == nullptr), "attn_weights_mean: r_q and r_k go together");
  const long long n_seq = (long long)n_outer * n_inner;
  const long long warps = n_seq * L;
  const unsigned grid = (unsigned)((warps + AW_WARPS - 1) / AW_WARPS);
  if (qkv_dtype == VDM_BF16)
    launch_kernel(attn_weights_mean_kernel<__nv_bfloat16>, grid, AW_WARPS * 32, 0, (cudaStream_t)(cudaStream_t)stream, 1, 
        (const __nv_bfloat16*)qkv, n_seq, (int)n_inner, outer_stride, inner_stride, seq_stride, L, heads, hd, r_q, r_k,
        mask, allow_pad_interactions, out);
  else
    launch_kernel(attn_weights_mean_kernel<float>, grid, AW_WARPS * 32, 0, (cudaStream_t)(cudaStream_t)stream, 1, 
        (const float*)qkv, n_seq, (int)n_inner, outer_stride, inner_stride, seq_stride, L, heads, hd, r_q, r_k, mask,
        allow_pad_interactions, out);
  VDM_AFTER_LAUNCH("attn_weights_mean");
  return 0;
}
