// HBM-bound producer kernels around the GEMMs: GroupNorm statistics / apply (+SiLU,
// +scale-shift, + concat, + x2 upsample, + stride-2 parity split), temporal GroupNorm,
// conditioning mix + input-conv im2col, timestep sinusoid, RPE-net hidden layer.
#include "common.cuh"

namespace vdm {
namespace {

// ------------------------------------------------------------------ GroupNorm statistics
// grid (pixel chunks, n_img); block = (C/V) x rows threads; each thread owns V channels.
template <int V>
__global__ void gn_stats_kernel(const float* __restrict__ s1, int C1, const float* __restrict__ s2, int C2, int HW,
                                int pix_per_block, double* __restrict__ stats) {
  __shared__ double sg[32][2];
  const int C = C1 + C2, CV = C / V, cpg = C / 32;
  const int cq = threadIdx.x % CV, prow = threadIdx.x / CV, rows = blockDim.x / CV;
  const int n = blockIdx.y;
  if (threadIdx.x < 64) (&sg[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  const int c = cq * V;
  const float* src;
  int cs, ld;
  if (c < C1) { src = s1; cs = c; ld = C1; } else { src = s2; cs = c - C1; ld = C2; }
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  double s = 0.0, ss = 0.0;
  if (prow < rows) {
    for (int p = p0 + prow; p < p1; p += rows) {
      const float* ptr = src + ((size_t)n * HW + p) * ld + cs;
      float v[V];
      if constexpr (V == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(ptr));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
      } else {
        const float2 t = __ldg(reinterpret_cast<const float2*>(ptr));
        v[0] = t.x; v[1] = t.y;
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        s += (double)v[i];
        ss += (double)v[i] * (double)v[i];
      }
    }
    const int g = c / cpg;  // V divides cpg, so the V channels share a group
    atomicAdd(&sg[g][0], s);
    atomicAdd(&sg[g][1], ss);
  }
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(&stats[(size_t)n * 64 + threadIdx.x], (&sg[0][0])[threadIdx.x]);
}

// ------------------------------------------------------------------ GroupNorm apply
struct ApplyParams {
  const float* s1; int C1;
  const float* s2; int C2;
  int n_img, H, W;
  const double* stats;
  const float* gamma; const float* beta;
  const float* ss; int ld_ss;
  int silu, out_mode;
  void* out; float* copy;
  int pix_per_block;
};

// grid (pixel chunks, n_img); dynamic smem: 2*C floats (per-channel multiplier / offset)
template <typename OutT>
__global__ void __launch_bounds__(256) gn_apply_kernel(const ApplyParams p) {
  extern __shared__ float sm[];
  const int C = p.C1 + p.C2, cpg = C / 32, HW = p.H * p.W;
  float* mulc = sm;
  float* addc = sm + C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 1.f, b = 0.f;
    if (p.stats) {
      const int g = c / cpg;
      const double cnt = (double)HW * cpg;
      const double mean = p.stats[(size_t)n * 64 + g * 2] / cnt;
      double var = p.stats[(size_t)n * 64 + g * 2 + 1] / cnt - mean * mean;
      if (var < 0) var = 0;
      const float rstd = (float)(1.0 / sqrt(var + 1e-5));
      a = rstd * p.gamma[c];
      b = p.beta[c] - (float)mean * a;
    }
    if (p.ss) {
      const float sc = 1.0f + p.ss[(size_t)n * p.ld_ss + c];
      a *= sc;
      b = b * sc + p.ss[(size_t)n * p.ld_ss + C + c];
    }
    mulc[c] = a;
    addc[c] = b;
  }
  __syncthreads();
  const int C8 = C / 8;
  const int p0 = blockIdx.x * p.pix_per_block;
  const int npix = min(HW, p0 + p.pix_per_block) - p0;
  OutT* out = reinterpret_cast<OutT*>(p.out);
  for (int idx = threadIdx.x; idx < npix * C8; idx += blockDim.x) {
    const int pl = idx / C8, c = (idx - pl * C8) * 8;
    const int pix = p0 + pl;
    const float* src = (c < p.C1) ? p.s1 + ((size_t)n * HW + pix) * p.C1 + c
                                  : p.s2 + ((size_t)n * HW + pix) * p.C2 + (c - p.C1);
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
    const float4 v1 = __ldg(reinterpret_cast<const float4*>(src + 4));
    float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float y = fmaf(v[i], mulc[c + i], addc[c + i]);
      if (p.silu) y = silu_precise(y);
      v[i] = y;
    }
    if (p.copy) {
      float* cp = p.copy + ((size_t)n * HW + pix) * C + c;
      *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(cp + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    auto store8 = [&](size_t row) {
      OutT* o = out + row * C + c;
      if constexpr (sizeof(OutT) == 2) {
        uint4 pk;
        pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]);
        pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(o) = pk;
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    };
    if (p.out_mode == 0) {
      store8((size_t)n * HW + pix);
    } else {
      const int y = pix / p.W, x = pix - y * p.W;
      if (p.out_mode == 1) {  // nearest x2
        const int W2 = 2 * p.W;
        const size_t r0 = ((size_t)n * 2 * p.H + 2 * y) * W2 + 2 * x;
        store8(r0); store8(r0 + 1); store8(r0 + W2); store8(r0 + W2 + 1);
      } else {                // parity planes of a stride-2 conv input
        const int Hh = p.H / 2, Wh = p.W / 2;
        const int plane = (y & 1) * 2 + (x & 1);
        store8((((size_t)n * 4 + plane) * Hh + (y >> 1)) * Wh + (x >> 1));
      }
    }
  }
}

// ------------------------------------------------------------------ temporal GroupNorm
// x: [B][T][HW][C]; one thread per (b, pixel, group): stats over T frames x cpg channels.
template <typename OutT>
__global__ void __launch_bounds__(256) gn_temporal_kernel(const float* __restrict__ x, int T, int HW, int C,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ out_f32,
                                                           OutT* __restrict__ out_a) {
  const int g = threadIdx.x & 31;
  const int pix = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int b = blockIdx.y;
  if (pix >= HW) return;
  const int cpg = C / 32;
  const size_t frame_stride = (size_t)HW * C;
  const float* base = x + (size_t)b * T * frame_stride + (size_t)pix * C + g * cpg;
  float s = 0.f;
  for (int t = 0; t < T; ++t)
    for (int j = 0; j < cpg; ++j) s += base[t * frame_stride + j];
  const float mean = s / (float)(T * cpg);
  float ss = 0.f;
  for (int t = 0; t < T; ++t)
    for (int j = 0; j < cpg; ++j) {
      const float d = base[t * frame_stride + j] - mean;
      ss = fmaf(d, d, ss);
    }
  const float rstd = rsqrtf(ss / (float)(T * cpg) + 1e-5f);
  const size_t obase = (size_t)b * T * frame_stride + (size_t)pix * C + g * cpg;
  for (int t = 0; t < T; ++t)
    for (int j = 0; j < cpg; ++j) {
      const int c = g * cpg + j;
      const float y = (base[t * frame_stride + j] - mean) * rstd * gamma[c] + beta[c];
      if (out_f32) out_f32[obase + t * frame_stride + j] = y;
      store_elem<OutT>(out_a + obase + t * frame_stride + j, y);
    }
}

__global__ void __launch_bounds__(256) add_spatial_encoding_kernel(const float* h, const float* __restrict__ enc,
                                                                    float* out, long long total4, long long per_img4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(h)[i];
    const float4 e = __ldg(reinterpret_cast<const float4*>(enc) + (i % per_img4));
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    reinterpret_cast<float4*>(out)[i] = v;
  }
}

// ------------------------------------------------------------------ conditioning mix + im2col
// One thread per output pixel (b, f, y, x): gathers the 3x3 neighbourhood of the 5-channel
// conditioned input and writes one 64-wide im2col row (k = tap*5 + c, zero padded).
template <typename OutT>
__global__ void __launch_bounds__(128) cond_mix_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                                        const float* __restrict__ obs, const float* __restrict__ lat,
                                                        const float* __restrict__ kinda, const float* __restrict__ t,
                                                        int B, int F, int H, int W, OutT* __restrict__ a_out,
                                                        float* __restrict__ t_frame, float* __restrict__ attn_mask) {
  const int HW = H * W;
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= (long long)B * F * HW) return;
  const int n = (int)(m / HW), pix = (int)(m - (long long)n * HW);
  const int y = pix / W, xx = pix - y * W;
  const float o = obs[n], l = lat[n], k = kinda[n];
  const float any = fminf(o + l + k, 1.0f);
  if (pix == 0) {
    const float tb = t[n / F];
    t_frame[n] = 0.0f * o + tb * (1.0f - o);
    attn_mask[n] = any;
  }
  float row[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) row[i] = 0.f;
  const float wx = l + (1.0f - any);  // x*latent + x*(1-anything)
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
      const int iy = y + r - 1, ix = xx + s - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const int tap = r * 3 + s;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const size_t off = (((size_t)n * 3 + c) * H + iy) * W + ix;
        // x*latent_mask + x0*obs_mask + x*(1-anything_mask), in the reference's op order
        row[tap * 5 + c] = __fadd_rn(__fadd_rn(__fmul_rn(x[off], l), __fmul_rn(x0[off], o)), __fmul_rn(x[off], 1.0f - any));
      }
      row[tap * 5 + 3] = o;
      row[tap * 5 + 4] = k;
    }
  (void)wx;
  OutT* dst = a_out + (size_t)m * 64;
  if constexpr (sizeof(OutT) == 2) {
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
      uint4 pk;
      pk.x = pack_bf16x2(row[i], row[i + 1]); pk.y = pack_bf16x2(row[i + 2], row[i + 3]);
      pk.z = pack_bf16x2(row[i + 4], row[i + 5]); pk.w = pack_bf16x2(row[i + 6], row[i + 7]);
      *reinterpret_cast<uint4*>(dst + i) = pk;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 64; i += 4)
      *reinterpret_cast<float4*>(dst + i) = make_float4(row[i], row[i + 1], row[i + 2], row[i + 3]);
  }
}

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, int dim, float* __restrict__ out) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int row = idx / half, i = idx - row * half;
  // freqs = exp(-ln(10000) * i / half) in fp32 (nn.py:98-101)
  const float freq = expf(__fdiv_rn(__fmul_rn(-9.210340371976184f, (float)i), (float)half));
  const float arg = __fmul_rn(t[row], freq);
  out[(size_t)row * dim + i] = cosf(arg);
  out[(size_t)row * dim + half + i] = sinf(arg);
  if ((dim & 1) && i == 0) out[(size_t)row * dim + dim - 1] = 0.f;
}

// hidden[net][(b*T+i)*T+j][c]
template <typename OutT>
__global__ void __launch_bounds__(256) rpe_hidden_kernel(const float* __restrict__ e_t, int ld_et,
                                                          const long long* __restrict__ fi, const float* __restrict__ wd,
                                                          const float* __restrict__ bd, int B, int T, int C,
                                                          OutT* __restrict__ out) {
  const int net = blockIdx.z;
  const int row = blockIdx.y;  // (b*T + i)*T + j
  const int bi = row / T, j = row - bi * T;
  const int b = bi / T;
  const long long d = fi[bi] - fi[b * T + j];
  const float df = (float)d;
  const float f0 = log1pf(fmaxf(df, 0.f)), f1 = log1pf(fmaxf(-df, 0.f)), f2 = (d == 0) ? 1.f : 0.f;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < C; c += gridDim.x * blockDim.x) {
    const float* w = wd + ((size_t)net * C + c) * 3;
    const float ed = fmaf(f2, w[2], fmaf(f1, w[1], f0 * w[0])) + bd[net * C + c];
    const float h = e_t[(size_t)bi * ld_et + net * C + c] + ed;
    store_elem<OutT>(out + ((size_t)net * B * T * T + row) * C + c, silu_precise(h));
  }
}

}  // namespace
}  // namespace vdm

using namespace vdm;

extern "C" int vdm_gn_stats(const float* src1, int32_t C1, const float* src2, int32_t C2, int32_t n_img, int32_t HW,
                            double* stats, vdm_stream_t stream) {
  const int C = C1 + C2;
  VDM_REQUIRE(src1 && stats && C1 > 0 && (C2 == 0 || src2), "gn_stats: NULL pointer");
  VDM_REQUIRE(C % 64 == 0 && C1 % 8 == 0 && C2 % 8 == 0 && C <= 2048, "gn_stats: unsupported channels %d+%d", C1, C2);
  const int cpg = C / 32;
  const int V = (cpg % 4 == 0) ? 4 : 2;
  const int CV = C / V;
  VDM_REQUIRE(CV <= 1024, "gn_stats: too many channels");
  const int rows = CV >= 256 ? 1 : 256 / CV;
  const int threads = CV * rows;
  int ppb = 256;
  while (ppb > 16 && (long long)((HW + ppb - 1) / ppb) * n_img < 2LL * num_sms()) ppb >>= 1;
  dim3 grid((HW + ppb - 1) / ppb, n_img);
  if (V == 4)
    gn_stats_kernel<4><<<grid, threads, 0, (cudaStream_t)stream>>>(src1, C1, src2, C2, HW, ppb, stats);
  else
    gn_stats_kernel<2><<<grid, threads, 0, (cudaStream_t)stream>>>(src1, C1, src2, C2, HW, ppb, stats);
  VDM_AFTER_LAUNCH("gn_stats");
  return 0;
}

extern "C" int vdm_gn_apply(const vdm_gn_apply_args* a, vdm_stream_t stream) {
  const int C = a->C1 + a->C2;
  VDM_REQUIRE(a->src1 && a->out && (a->C2 == 0 || a->src2), "gn_apply: NULL pointer");
  VDM_REQUIRE(C % 64 == 0 && a->C1 % 8 == 0 && a->C2 % 8 == 0, "gn_apply: unsupported channels %d+%d", a->C1, a->C2);
  VDM_REQUIRE(!a->stats || (a->gamma && a->beta), "gn_apply: gamma/beta missing");
  VDM_REQUIRE(a->out_mode >= 0 && a->out_mode <= 2, "gn_apply: bad out_mode");
  VDM_REQUIRE(a->out_mode != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "gn_apply: parity split needs even H, W");
  VDM_REQUIRE(a->out_mode == 0 || a->out_f32_copy == nullptr, "gn_apply: fp32 copy only with plain output");
  ApplyParams p{a->src1, a->C1, a->src2, a->C2, a->n_img, a->H, a->W, a->stats, a->gamma, a->beta,
                a->scale_shift, a->ld_ss, a->silu, a->out_mode, a->out, a->out_f32_copy, 0};
  const int HW = a->H * a->W;
  int ppb = 64;
  while (ppb > 4 && (long long)((HW + ppb - 1) / ppb) * a->n_img < 4LL * num_sms()) ppb >>= 1;
  p.pix_per_block = ppb;
  dim3 grid((HW + ppb - 1) / ppb, a->n_img);
  const size_t smem = 2 * (size_t)C * sizeof(float);
  if (a->out_dtype == VDM_BF16)
    gn_apply_kernel<__nv_bfloat16><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  else
    gn_apply_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>(p);
  VDM_AFTER_LAUNCH("gn_apply");
  return 0;
}

extern "C" int vdm_gn_temporal(const float* x, int32_t B, int32_t T, int32_t HW, int32_t C, const float* gamma,
                               const float* beta, float* out_f32, void* out_a, int32_t out_dtype, vdm_stream_t stream) {
  VDM_REQUIRE(x && gamma && beta && out_a, "gn_temporal: NULL pointer");
  VDM_REQUIRE(C % 32 == 0, "gn_temporal: C must be a multiple of 32");
  dim3 grid((HW + 7) / 8, B);
  if (out_dtype == VDM_BF16)
    gn_temporal_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(x, T, HW, C, gamma, beta, out_f32,
                                                                             (__nv_bfloat16*)out_a);
  else
    gn_temporal_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(x, T, HW, C, gamma, beta, out_f32, (float*)out_a);
  VDM_AFTER_LAUNCH("gn_temporal");
  return 0;
}

extern "C" int vdm_add_spatial_encoding(const float* h, const float* enc, float* out, int32_t n_img, int32_t HW,
                                        int32_t C, vdm_stream_t stream) {
  VDM_REQUIRE(h && enc && out && C % 4 == 0, "add_spatial_encoding: bad arguments");
  const long long per4 = (long long)HW * C / 4, total4 = per4 * n_img;
  const int grid = (int)std::min<long long>((total4 + 255) / 256, (long long)num_sms() * 16);
  add_spatial_encoding_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, enc, out, total4, per4);
  VDM_AFTER_LAUNCH("add_spatial_encoding");
  return 0;
}

extern "C" int vdm_cond_mix(const float* x, const float* x0, const float* obs_mask, const float* latent_mask,
                            const float* kinda_marg_mask, const float* t, int32_t B, int32_t F, int32_t H, int32_t W,
                            void* a_out, int32_t out_dtype, float* t_frame, float* attn_mask, vdm_stream_t stream) {
  VDM_REQUIRE(x && x0 && obs_mask && latent_mask && kinda_marg_mask && t && a_out && t_frame && attn_mask,
              "cond_mix: NULL pointer");
  const long long M = (long long)B * F * H * W;
  const int grid = (int)((M + 127) / 128);
  if (out_dtype == VDM_BF16)
    cond_mix_kernel<__nv_bfloat16><<<grid, 128, 0, (cudaStream_t)stream>>>(x, x0, obs_mask, latent_mask, kinda_marg_mask,
                                                                          t, B, F, H, W, (__nv_bfloat16*)a_out, t_frame,
                                                                          attn_mask);
  else
    cond_mix_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(x, x0, obs_mask, latent_mask, kinda_marg_mask, t, B, F,
                                                                  H, W, (float*)a_out, t_frame, attn_mask);
  VDM_AFTER_LAUNCH("cond_mix");
  return 0;
}

extern "C" int vdm_timestep_embedding(const float* t_frame, int32_t n, int32_t dim, float* out, vdm_stream_t stream) {
  VDM_REQUIRE(t_frame && out && n > 0 && dim >= 2, "timestep_embedding: bad arguments");
  const int total = n * (dim / 2);
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t_frame, n, dim, out);
  VDM_AFTER_LAUNCH("timestep_embedding");
  return 0;
}

extern "C" int vdm_rpe_hidden(const float* e_t, int32_t ld_et, const int64_t* frame_indices, const float* wd,
                              const float* bd, int32_t B, int32_t T, int32_t C, void* out, int32_t out_dtype,
                              vdm_stream_t stream) {
  VDM_REQUIRE(e_t && frame_indices && wd && bd && out, "rpe_hidden: NULL pointer");
  dim3 grid((C + 255) / 256, B * T * T, 3);
  if (out_dtype == VDM_BF16)
    rpe_hidden_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(e_t, ld_et, (const long long*)frame_indices,
                                                                            wd, bd, B, T, C, (__nv_bfloat16*)out);
  else
    rpe_hidden_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(e_t, ld_et, (const long long*)frame_indices, wd, bd,
                                                                    B, T, C, (float*)out);
  VDM_AFTER_LAUNCH("rpe_hidden");
  return 0;
}
