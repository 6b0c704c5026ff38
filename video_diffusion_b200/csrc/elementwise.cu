// HBM-bound producer kernels around the GEMMs: GroupNorm statistics / apply (+SiLU,
// +scale-shift, + concat, + x2 upsample, + stride-2 parity split), temporal GroupNorm,
// conditioning mix + input-conv im2col, timestep sinusoid, RPE-net hidden layer.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

#ifndef GN_MIN_BLOCKS
#define GN_MIN_BLOCKS 7      // 128-thread blocks of <= 72 registers: three of them fit beside a persistent GEMM CTA
#endif
#ifndef GN_THREADS
#define GN_THREADS 128      // default / maximum block size of gn_apply (profiles: co-residency experiments)
#endif

namespace vdm {
namespace {

// The residual stream is fp32 (reference-accuracy mode, tests) or fp16 (the bf16 model): IoT = float | __half.
template <typename IoT>
__device__ __forceinline__ float4 load4(const IoT* p) {
  if constexpr (sizeof(IoT) == 2) {
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p));
    const float2 lo = unpack_f16x2(h.x), hi = unpack_f16x2(h.y);
    return make_float4(lo.x, lo.y, hi.x, hi.y);
  } else {
    return __ldg(reinterpret_cast<const float4*>(p));
  }
}
template <typename IoT>
__device__ __forceinline__ void store4(IoT* p, float4 v) {
  if constexpr (sizeof(IoT) == 2) {
    uint2 h;
    h.x = pack_f16x2(v.x, v.y);
    h.y = pack_f16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = h;
  } else {
    *reinterpret_cast<float4*>(p) = v;
  }
}

// ------------------------------------------------------------------ GroupNorm statistics
// Per-(image, channel) sum and sum of squares: stats[n][0][c], stats[n][1][c] (double).
// grid (pixel chunks, n_img); block = (C/4) x rows threads; each thread owns 4 channels.
template <typename IoT>
__global__ void gn_stats_kernel(const IoT* __restrict__ src, int C, int HW, int pix_per_block,
                                double* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int C4 = C / 4;
  const int cq = threadIdx.x % C4, prow = threadIdx.x / C4, rows = blockDim.x / C4;
  const int n = blockIdx.y;
  const int p0 = blockIdx.x * pix_per_block;
  const int p1 = min(HW, p0 + pix_per_block);
  double s[4] = {0, 0, 0, 0}, ss[4] = {0, 0, 0, 0};
  for (int p = p0 + prow; p < p1; p += rows) {
    const float4 t = load4(src + ((size_t)n * HW + p) * C + cq * 4);
    const float v[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      s[i] += (double)v[i];
      ss[i] += (double)v[i] * (double)v[i];
    }
  }
  double* d = stats + (size_t)n * 2 * C + cq * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    atomicAdd(d + i, s[i]);
    atomicAdd(d + C + i, ss[i]);
  }
}

// ------------------------------------------------------------------ GroupNorm apply
struct ApplyParams {
  const void* s1; int C1;
  const float* s2; int C2;
  int n_img, H, W;
  const void* st1; const void* st2; int st_kind, st_kind2;   // VDM_F64: double sums, VDM_I64: fixed-point 2^-24
  const float* gamma; const float* beta;
  const float* ss; int ld_ss;
  int silu, out_mode;
  void* out; void* out_raw; void* copy;
  int copy_f16;
  int pix_per_block;
  const unsigned int* wait_done;   // per-image completion counters of the producer of s1 / st1 (vdm_gn_apply_args.wait_done)
  unsigned int wait_count;
};

// grid (pixel chunks, n_img); block = (C/8) x rows threads: a thread owns 8 channels for a strided set
// of pixels, so its per-channel multiplier / offset live in registers for the whole loop.
// dynamic smem: 2*C doubles (per-channel sums) + 64 floats (group mean / rstd)
// MODE: 0 plain, 1 nearest-x2, 2 stride-2 parity planes; RAW / COPY: optional extra outputs.
// InT: float, __half (fp16 residual stream; two-source concat allowed like float) or __nv_bfloat16 (conv output kept in
// bf16, single source).
template <typename OutT, typename InT, int MODE, bool RAW, bool COPY>
__global__ void __launch_bounds__(GN_THREADS, GN_MIN_BLOCKS) gn_apply_kernel(const ApplyParams p) {
  pdl_launch_dependents();
  // polled mode (wait_done): this grid runs BESIDE its producer and waits per image, below, instead of for the whole grid
  if (p.wait_done == nullptr) pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int C = p.C1 + p.C2, cpg = C / 32, HW = p.H * p.W, C8 = C / 8;
  double* chs = reinterpret_cast<double*>(sm_raw);
  double* chss = chs + C;
  float* gmean = reinterpret_cast<float*>(chss + C);
  float* grstd = gmean + 32;
  const int n = blockIdx.y;
  const int c8 = threadIdx.x % C8, prow = threadIdx.x / C8, rows = blockDim.x / C8;
  const int c = c8 * 8;
  float a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = 1.f; b[i] = 0.f; }
  // Everything that does not depend on the statistics is requested first, so the small launches (8x8, 16x16 levels)
  // wait for one memory round trip instead of a chain of four: the first pixel rows of this thread go to L2, the
  // affine parameters into registers.
  {
    const int ldp = (c < p.C1) ? p.C1 : p.C2;
    constexpr size_t esz = sizeof(InT);
    const char* base = (c < p.C1) ? reinterpret_cast<const char*>(p.s1) + ((size_t)n * HW * p.C1 + c) * esz
                                  : reinterpret_cast<const char*>(p.s2) + ((size_t)n * HW * p.C2 + (c - p.C1)) * esz;
    const int pf0 = blockIdx.x * p.pix_per_block + prow;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int pix = pf0 + u * rows;
      if (pix < HW) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)pix * ldp * esz));
    }
  }
  float gam[8], bet[8];
  if (p.st1 != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { gam[i] = __ldg(p.gamma + c + i); bet[i] = __ldg(p.beta + c + i); }
  }
  if (p.wait_done != nullptr) {
    // the producer (a conv kernel whose CTAs are all resident) publishes image n with release atomics once its rows
    // and statistics are globally visible; it can only make progress, so this wait is bounded
    if (threadIdx.x == 0) {
      const unsigned int* flag = p.wait_done + n;
      const long long t0 = clock64();
      for (;;) {
        unsigned int v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v >= p.wait_count) break;
        __nanosleep(200);
        if (clock64() - t0 > 4000000000LL) {     // ~2 s: a protocol bug must fail loudly, never hang the box
          printf("vdm gn_apply: image %d never completed (%u of %u)\n", n, v, p.wait_count);
          __trap();
        }
      }
    }
    __syncthreads();
  }
  if (p.st1 != nullptr) {
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const bool first = ch < p.C1;
      const void* st = first ? p.st1 : p.st2;
      const int Cs = first ? p.C1 : p.C2, cs = first ? ch : ch - p.C1;
      const size_t i0 = (size_t)n * 2 * Cs + cs;
      if ((first ? p.st_kind : p.st_kind2) == VDM_F64) {
        chs[ch] = reinterpret_cast<const double*>(st)[i0];
        chss[ch] = reinterpret_cast<const double*>(st)[i0 + Cs];
      } else {
        chs[ch] = (double)reinterpret_cast<const long long*>(st)[i0] * (1.0 / 16777216.0);
        chss[ch] = (double)reinterpret_cast<const long long*>(st)[i0 + Cs] * (1.0 / 16777216.0);
      }
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      double s = 0, q = 0;
      for (int j = 0; j < cpg; ++j) { s += chs[threadIdx.x * cpg + j]; q += chss[threadIdx.x * cpg + j]; }
      const double cnt = (double)HW * cpg;
      const double mean = s / cnt;
      double var = q / cnt - mean * mean;
      if (var < 0) var = 0;
      gmean[threadIdx.x] = (float)mean;
      grstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c + i) / cpg;
      a[i] = grstd[g] * gam[i];
      b[i] = bet[i] - gmean[g] * a[i];
    }
  }
  if (p.ss != nullptr) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float sc = 1.0f + p.ss[(size_t)n * p.ld_ss + c + i];
      a[i] *= sc;
      b[i] = b[i] * sc + p.ss[(size_t)n * p.ld_ss + C + c + i];
    }
  }
  const bool silu = p.silu != 0;
  // bf16 input (a conv output kept only in bf16) is single-source; fp32 / fp16 input may be a two-source concat
  const int ld = (c < p.C1) ? p.C1 : p.C2;
  const InT* src = (c < p.C1) ? reinterpret_cast<const InT*>(p.s1) + (size_t)n * HW * p.C1 + c
                              : reinterpret_cast<const InT*>(p.s2) + (size_t)n * HW * p.C2 + (c - p.C1);
  OutT* const out = reinterpret_cast<OutT*>(p.out) + c;
  OutT* const out_raw = RAW ? reinterpret_cast<OutT*>(p.out_raw) + (size_t)n * HW * C + c : nullptr;
  const size_t copy_off = (size_t)n * HW * C + c;
  const int p0 = blockIdx.x * p.pix_per_block;
  const int p1 = min(HW, p0 + p.pix_per_block);
  auto load8 = [&](const InT* ptr, float (&x)[8]) {
    if constexpr (std::is_same<InT, __half>::value) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(ptr));
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = unpack_f16x2(w[i]);
        x[2 * i] = f.x;
        x[2 * i + 1] = f.y;
      }
    } else if constexpr (sizeof(InT) == 2) {
      const uint4 r = __ldg(reinterpret_cast<const uint4*>(ptr));
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {       // bf16 -> fp32 is a 16-bit shift
        x[2 * i] = __uint_as_float(w[i] << 16);
        x[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
      }
    } else {
      const float4 v0 = __ldg(reinterpret_cast<const float4*>(ptr));
      const float4 v1 = __ldg(reinterpret_cast<const float4*>(ptr + 4));
      x[0] = v0.x; x[1] = v0.y; x[2] = v0.z; x[3] = v0.w; x[4] = v1.x; x[5] = v1.y; x[6] = v1.z; x[7] = v1.w;
    }
  };
  auto store8 = [&](OutT* o, const float (&v)[8]) {
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
  auto process = [&](int pix, const float (&x)[8]) {
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      y[i] = fmaf(x[i], a[i], b[i]);
      if (silu) y[i] = (sizeof(OutT) == 2) ? silu_tanh(y[i]) : silu_precise(y[i]);
    }
    if constexpr (RAW) store8(out_raw + (size_t)pix * C, x);
    if constexpr (COPY) {
      const size_t o = copy_off + (size_t)pix * C;
      if (p.copy_f16) {
        uint4 pk;
        pk.x = pack_f16x2(y[0], y[1]); pk.y = pack_f16x2(y[2], y[3]);
        pk.z = pack_f16x2(y[4], y[5]); pk.w = pack_f16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.copy) + o) = pk;
      } else {
        float* cp = reinterpret_cast<float*>(p.copy) + o;
        *reinterpret_cast<float4*>(cp) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(cp + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
    if constexpr (MODE == 0) {
      if (!COPY || p.out != nullptr) store8(out + ((size_t)n * HW + pix) * C, y);   // (COPY alone: `out` may be NULL)
    } else {
      const int yy = pix / p.W, xx = pix - yy * p.W;
      if constexpr (MODE == 1) {  // nearest x2
        const int W2 = 2 * p.W;
        const size_t r0 = ((size_t)n * 2 * p.H + 2 * yy) * W2 + 2 * xx;
        store8(out + r0 * C, y); store8(out + (r0 + 1) * C, y);
        store8(out + (r0 + W2) * C, y); store8(out + (r0 + W2 + 1) * C, y);
      } else {                    // parity planes of a stride-2 conv input
        const int Hh = p.H / 2, Wh = p.W / 2;
        const int plane = (yy & 1) * 2 + (xx & 1);
        store8(out + ((((size_t)n * 4 + plane) * Hh + (yy >> 1)) * Wh + (xx >> 1)) * C, y);
      }
    }
  };
  // four pixels per iteration: all loads are in flight before any is consumed
  constexpr int U = 4;
  int pix = p0 + prow;
  for (; pix + (U - 1) * rows < p1; pix += U * rows) {
    float x[U][8];
    if constexpr (std::is_same<InT, __half>::value) {   // all raw loads first, then the conversions (see gn_temporal)
      uint4 raw[U];
#pragma unroll
      for (int u = 0; u < U; ++u) raw[u] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(pix + u * rows) * ld));
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t w[4] = {raw[u].x, raw[u].y, raw[u].z, raw[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = unpack_f16x2(w[i]);
          x[u][2 * i] = f.x;
          x[u][2 * i + 1] = f.y;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u) load8(src + (size_t)(pix + u * rows) * ld, x[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) process(pix + u * rows, x[u]);
  }
  for (; pix < p1; pix += rows) {
    float x[8];
    load8(src + (size_t)pix * ld, x);
    process(pix, x);
  }
}

// The prologue of gn_apply_kernel as a kernel of its own: per-(image, channel) multiplier / offset of GroupNorm32
// (+ scale/shift).  Same arithmetic in the same order, so a consumer that applies act(a*x + b) itself (the GEMM's
// transform warps) produces bit-identical operands.  grid = n_img, dynamic smem as gn_apply_kernel.
__global__ void __launch_bounds__(256) gn_coef_kernel(const void* st1, int st_kind, int C1, const void* st2, int st_kind2,
                                                       int C2, int HW, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, const float* __restrict__ ss,
                                                       int ld_ss, float2* __restrict__ coef) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int C = C1 + C2, cpg = C / 32;
  double* chs = reinterpret_cast<double*>(sm_raw);
  double* chss = chs + C;
  float* gmean = reinterpret_cast<float*>(chss + C);
  float* grstd = gmean + 32;
  const int n = blockIdx.x;
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const bool first = ch < C1;
    const void* st = first ? st1 : st2;
    const int Cs = first ? C1 : C2, cs = first ? ch : ch - C1;
    const size_t i0 = (size_t)n * 2 * Cs + cs;
    if ((first ? st_kind : st_kind2) == VDM_F64) {
      chs[ch] = reinterpret_cast<const double*>(st)[i0];
      chss[ch] = reinterpret_cast<const double*>(st)[i0 + Cs];
    } else {
      chs[ch] = (double)reinterpret_cast<const long long*>(st)[i0] * (1.0 / 16777216.0);
      chss[ch] = (double)reinterpret_cast<const long long*>(st)[i0 + Cs] * (1.0 / 16777216.0);
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    double s = 0, q = 0;
    for (int j = 0; j < cpg; ++j) { s += chs[threadIdx.x * cpg + j]; q += chss[threadIdx.x * cpg + j]; }
    const double cnt = (double)HW * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0) var = 0;
    gmean[threadIdx.x] = (float)mean;
    grstd[threadIdx.x] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g = ch / cpg;
    float a = grstd[g] * __ldg(gamma + ch);
    float b = __ldg(beta + ch) - gmean[g] * a;
    if (ss != nullptr) {
      const float sc = 1.0f + ss[(size_t)n * ld_ss + ch];
      a *= sc;
      b = b * sc + ss[(size_t)n * ld_ss + C + ch];
    }
    coef[(size_t)n * C + ch] = make_float2(a, b);
  }
}

// ------------------------------------------------------------------ temporal GroupNorm
// x: [B][T][HW][C]; GroupNorm over (C/32 channels x T frames) per (b, pixel).  One warp per pixel:
// pass 1 streams the pixel's T x C values with coalesced float4 loads and accumulates per-channel
// sums in shared memory, 32 lanes fold them into group statistics, pass 2 re-reads (L1/L2-hot),
// normalises and writes the fp32 residual copy and the GEMM A operand.
template <typename OutT, typename IoT>
__global__ void __launch_bounds__(256) gn_temporal_kernel(const IoT* __restrict__ x, int T, int HW, int C,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, IoT* __restrict__ out_f32,
                                                           OutT* __restrict__ out_a) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) float smt[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pix = blockIdx.x * 8 + warp;
  const int b = blockIdx.y;
  if (pix >= HW) return;
  float* chs = smt + warp * (2 * C + 64);
  float* chq = chs + C;
  float* ga = chq + C;          // per-group multiplier (rstd) and mean
  float* gm = ga + 32;
  const int cpg = C / 32, C4 = C / 4;
  const size_t frame_stride = (size_t)HW * C;
  const IoT* base = x + (size_t)b * T * frame_stride + (size_t)pix * C;
  // Per-channel sums are taken about a pivot (the channel's value in the first frame), so channels with
  // |mean| >> std lose no precision to E[x^2] - mean^2 cancellation (the reference's GroupNorm is two-pass).
  // chs <- per-channel mean, chq <- per-channel sum of squared deviations from that mean.
  for (int q4 = lane; q4 < C4; q4 += 32) {
    const float4 v0 = load4(base + q4 * 4);
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
    for (int t = 1; t < T; ++t) {
      float4 v = load4(base + t * frame_stride + q4 * 4);
      v.x -= v0.x; v.y -= v0.y; v.z -= v0.z; v.w -= v0.w;
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
    }
    const float inv_t = 1.0f / (float)T;
    *reinterpret_cast<float4*>(chs + q4 * 4) =
        make_float4(v0.x + s.x * inv_t, v0.y + s.y * inv_t, v0.z + s.z * inv_t, v0.w + s.w * inv_t);
    *reinterpret_cast<float4*>(chq + q4 * 4) =
        make_float4(q.x - s.x * s.x * inv_t, q.y - s.y * s.y * inv_t, q.z - s.z * s.z * inv_t, q.w - s.w * s.w * inv_t);
  }
  __syncwarp();
  {
    float s = 0.f, m2 = 0.f;
    for (int j = 0; j < cpg; ++j) s += chs[lane * cpg + j];
    const float mean = s / (float)cpg;
    for (int j = 0; j < cpg; ++j) {
      const float d = chs[lane * cpg + j] - mean;
      m2 += chq[lane * cpg + j] + (float)T * d * d;
    }
    const float var = fmaxf(m2 / (float)(T * cpg), 0.f);
    gm[lane] = mean;
    ga[lane] = rsqrtf(var + 1e-5f);
  }
  __syncwarp();
  for (int q4 = lane; q4 < C4; q4 += 32) {
    const int c = q4 * 4;
    float a[4], bb[4], mu[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int g = (c + i) / cpg;
      a[i] = ga[g] * gamma[c + i];
      bb[i] = beta[c + i];
      mu[i] = gm[g];
    }
    for (int t = 0; t < T; ++t) {
      const size_t off = (size_t)b * T * frame_stride + t * frame_stride + (size_t)pix * C + c;
      const float4 v = load4(x + off);
      const float y0 = fmaf(v.x - mu[0], a[0], bb[0]), y1 = fmaf(v.y - mu[1], a[1], bb[1]);
      const float y2 = fmaf(v.z - mu[2], a[2], bb[2]), y3 = fmaf(v.w - mu[3], a[3], bb[3]);
      if (out_f32) store4(out_f32 + off, make_float4(y0, y1, y2, y3));
      if (out_a == nullptr) {
        // the normalised copy alone (fp16 stream: it is itself the A operand of the qkv projection)
      } else if constexpr (sizeof(OutT) == 2) {
        uint2 pk;
        pk.x = pack_bf16x2(y0, y1);
        pk.y = pack_bf16x2(y2, y3);
        *reinterpret_cast<uint2*>(out_a + off) = pk;
      } else {
        *reinterpret_cast<float4*>(out_a + off) = make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// Single-pass variant for C % 128 == 0 and T <= 32: a warp owns one (b, pixel, slab of `gps` groups); each lane keeps
// its float4 column of all T frames in registers, the lanes of a group fold their sums with shuffles, and the values
// are normalised and written without touching memory twice.  B * HW * n_slabs warps: enough parallelism for the
// 8x8 level too (the kernel above launches only B * HW / 8 blocks there).
template <typename OutT, int TMAX, typename IoT>
__global__ void __launch_bounds__(128) gn_temporal_regs_kernel(const IoT* __restrict__ x, int B, int T, int HW, int C,
                                                                const float* __restrict__ gamma,
                                                                const float* __restrict__ beta,
                                                                IoT* __restrict__ out_f32, OutT* __restrict__ out_a,
                                                                int gps, int n_slabs) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long wg = (long long)blockIdx.x * 4 + warp;
  if (wg >= (long long)B * HW * n_slabs) return;
  const int slab = (int)(wg % n_slabs);
  const int pix = (int)((wg / n_slabs) % HW), b = (int)(wg / ((long long)n_slabs * HW));
  const int cpg = C / 32, l4 = cpg / 4;            // channels / float4 lanes per group
  const int g = slab * gps + lane / l4;
  const bool active = lane / l4 < gps && g < 32;
  const int c = active ? g * cpg + (lane % l4) * 4 : 0;
  const size_t frame_stride = (size_t)HW * C;
  const size_t off0 = (size_t)b * T * frame_stride + (size_t)pix * C + c;
  // TMAX >= T: the frame loops are fully unrolled, every load is in flight before the first use.  fp16 input stays
  // PACKED in registers (two per frame instead of four) and is unpacked in each of the three passes -- the
  // conversions are exact and cheap, and the halved register count doubles the warps (= bytes in flight) per SM,
  // which is what bounds this latency-limited kernel.
  constexpr bool PACKED = sizeof(IoT) == 2;
  float4 v[PACKED ? 1 : TMAX];
  uint2 raw[PACKED ? TMAX : 1];
  const IoT* src = x + off0;
  if constexpr (PACKED) {
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      if (t < T && active) raw[t] = __ldg(reinterpret_cast<const uint2*>(src + t * frame_stride));
    }
  } else {
#pragma unroll
    for (int t = 0; t < TMAX; ++t) {
      if (t < T && active) v[t] = load4(src + t * frame_stride);
    }
  }
  auto val = [&](int t) -> float4 {
    if constexpr (PACKED) {
      const float2 lo = unpack_f16x2(raw[t].x), hi = unpack_f16x2(raw[t].y);
      return make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
      return v[t];
    }
  };
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T && active) {
      const float4 u = val(t);
      s += (u.x + u.y) + (u.z + u.w);
    }
  }
  const int base = lane - lane % l4;
  float gs = 0.f;
  for (int j = 0; j < l4; ++j) gs += __shfl_sync(0xffffffffu, s, (base + j) & 31);
  const float cnt = (float)(T * cpg);
  const float mean = gs / cnt;
  // two-pass variance like the reference's GroupNorm: the values are still in registers, so the centred second
  // pass costs no memory traffic and channels with |mean| >> std keep their precision
  float q = 0.f;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T && active) {
      const float4 u = val(t);
      const float dx = u.x - mean, dy = u.y - mean, dz = u.z - mean, dw = u.w - mean;
      q = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, fmaf(dw, dw, q))));
    }
  }
  float gq = 0.f;
  for (int j = 0; j < l4; ++j) gq += __shfl_sync(0xffffffffu, q, (base + j) & 31);
  if (!active) return;
  const float rstd = rsqrtf(fmaxf(gq / cnt, 0.f) + 1e-5f);
  const float4 gm = *reinterpret_cast<const float4*>(gamma + c), bt = *reinterpret_cast<const float4*>(beta + c);
  const float a0 = rstd * gm.x, a1 = rstd * gm.y, a2 = rstd * gm.z, a3 = rstd * gm.w;
  const float b0 = bt.x, b1 = bt.y, b2 = bt.z, b3 = bt.w;
#pragma unroll
  for (int t = 0; t < TMAX; ++t) {
    if (t < T) {
      const size_t off = off0 + t * frame_stride;
      const float4 u = val(t);
      const float y0 = fmaf(u.x - mean, a0, b0), y1 = fmaf(u.y - mean, a1, b1);
      const float y2 = fmaf(u.z - mean, a2, b2), y3 = fmaf(u.w - mean, a3, b3);
      if (out_f32) store4(out_f32 + off, make_float4(y0, y1, y2, y3));
      if (out_a == nullptr) {
        // the normalised copy alone (fp16 stream: it is itself the A operand of the qkv projection)
      } else if constexpr (sizeof(OutT) == 2) {
        uint2 pk;
        pk.x = pack_bf16x2(y0, y1);
        pk.y = pack_bf16x2(y2, y3);
        *reinterpret_cast<uint2*>(out_a + off) = pk;
      } else {
        *reinterpret_cast<float4*>(out_a + off) = make_float4(y0, y1, y2, y3);
      }
    }
  }
}

// out = (h + enc[pixel]) + frame_emb[image]; either addend may be absent (unet.py:841-844, 914-926)
template <typename IoT>
__global__ void __launch_bounds__(256) add_spatial_encoding_kernel(const IoT* h, const float* __restrict__ enc,
                                                                    const float* __restrict__ frame_emb, IoT* out,
                                                                    long long total4, long long per_img4, int C4) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
       i += (long long)gridDim.x * blockDim.x) {
    float4 v;
    if constexpr (sizeof(IoT) == 2) {     // out may alias h: plain loads
      const uint2 hh = reinterpret_cast<const uint2*>(h)[i];
      const float2 lo = unpack_f16x2(hh.x), hi = unpack_f16x2(hh.y);
      v = make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
      v = reinterpret_cast<const float4*>(h)[i];
    }
    if (enc) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(enc) + (i % per_img4));
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    }
    if (frame_emb) {
      const float4 e = __ldg(reinterpret_cast<const float4*>(frame_emb) + (i / per_img4) * C4 + (i % C4));
      v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    }
    store4(out + 4 * i, v);
  }
}

// ------------------------------------------------------------------ conditioning mix + im2col
// One thread per output pixel (b, f, y, x): gathers the 3x3 neighbourhood of the conditioned input and writes one
// 64-wide im2col row (k = tap*CIN + c, zero padded).  MODE follows cond_emb_type (unet.py:975-1019):
//   0 'channel'   CIN 5: x*lat + observed*obs + x*(1-any) | obs indicator | kinda_marg indicator; t_frame = t*(1-obs)
//   1 'duplicate' CIN 6: x*lat + x*(1-any) | x0*obs;                                               t_frame = t
//   2 't=0'       CIN 3: x unchanged;                                                              t_frame = t
template <typename OutT, int MODE>
__global__ void __launch_bounds__(128) cond_mix_kernel(const float* __restrict__ x, const float* __restrict__ x0,
                                                        const float* __restrict__ obs, const float* __restrict__ lat,
                                                        const float* __restrict__ kinda, const float* __restrict__ t,
                                                        int B, int F, int H, int W, OutT* __restrict__ a_out,
                                                        float* __restrict__ t_frame, float* __restrict__ attn_mask) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  constexpr int CIN = MODE == 0 ? 5 : (MODE == 1 ? 6 : 3);
  const int HW = H * W;
  const long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (m >= (long long)B * F * HW) return;
  const int n = (int)(m / HW), pix = (int)(m - (long long)n * HW);
  const int y = pix / W, xx = pix - y * W;
  const float o = obs[n], l = lat[n], k = kinda[n];
  const float any = fminf(o + l + k, 1.0f);
  if (pix == 0) {
    const float tb = t[n / F];
    t_frame[n] = MODE == 0 ? 0.0f * o + tb * (1.0f - o) : tb;
    attn_mask[n] = any;
  }
  float row[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) row[i] = 0.f;
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
        if constexpr (MODE == 0) {
          // x*latent_mask + observed*obs_mask + x*(1-anything_mask), in the reference's op order
          row[tap * CIN + c] =
              __fadd_rn(__fadd_rn(__fmul_rn(x[off], l), __fmul_rn(x0[off], o)), __fmul_rn(x[off], 1.0f - any));
        } else if constexpr (MODE == 1) {
          row[tap * CIN + c] = __fadd_rn(__fmul_rn(x[off], l), __fmul_rn(x[off], 1.0f - any));
          row[tap * CIN + 3 + c] = __fmul_rn(x0[off], o);
        } else {
          row[tap * CIN + c] = x[off];
        }
      }
      if constexpr (MODE == 0) {
        row[tap * CIN + 3] = o;
        row[tap * CIN + 4] = k;
      }
    }
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

__global__ void timestep_embedding_kernel(const float* __restrict__ t, int n, int dim, float neg_log_period,
                                          float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * half) return;
  const int row = idx / half, i = idx - row * half;
  // freqs = exp(-ln(max_period) * i / half) in fp32 (nn.py:98-101)
  const float freq = expf(__fdiv_rn(__fmul_rn(neg_log_period, (float)i), (float)half));
  const float arg = __fmul_rn(t[row], freq);
  out[(size_t)row * dim + i] = cosf(arg);
  out[(size_t)row * dim + half + i] = sinf(arg);
  if ((dim & 1) && i == 0) out[(size_t)row * dim + dim - 1] = 0.f;
}

// hidden[net][(b*T+i)*T+j][c]; one thread = 8 channels of one (net, b, i) and walks j, so the distance-embedding
// weights, its bias and the diffusion-time term stay in registers (unet.py:283-296)
template <typename OutT>
__global__ void __launch_bounds__(256) rpe_hidden_kernel(const float* __restrict__ e_t, int ld_et,
                                                          const int* __restrict__ et_off, int n_nets,
                                                          const long long* __restrict__ fi, const float* __restrict__ wd,
                                                          const float* __restrict__ bd, int B, int T, int C,
                                                          OutT* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int C8 = C / 8;
  const int BT = B * T;
  const int total = n_nets * BT * C8;
  for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < total; v += gridDim.x * blockDim.x) {
    const int c = (v % C8) * 8;
    const int rr = v / C8;
    const int bi = rr % BT, net = rr / BT;
    const int b = bi / T;
    // nets come in (q, k, v) triples, one triple per attention block; et_off gives each block's first column
    const float* et = e_t + (size_t)bi * ld_et + (et_off ? et_off[net / 3] : 0) + (net % 3) * C + c;
    float w0[8], w1[8], w2[8], base[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float* w = wd + ((size_t)net * C + c + i) * 3;
      w0[i] = w[0]; w1[i] = w[1]; w2[i] = w[2];
      base[i] = bd[net * C + c + i];
    }
    float e[8];
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(et)), bb = __ldg(reinterpret_cast<const float4*>(et + 4));
      e[0] = a.x; e[1] = a.y; e[2] = a.z; e[3] = a.w; e[4] = bb.x; e[5] = bb.y; e[6] = bb.z; e[7] = bb.w;
    }
    const long long fi_i = fi[bi];
    OutT* o = out + ((size_t)net * BT * T + (size_t)bi * T) * C + c;
    for (int j = 0; j < T; ++j, o += C) {
      const long long d = fi_i - fi[b * T + j];
      const float df = (float)d;
      const float f0 = log1pf(fmaxf(df, 0.f)), f1 = log1pf(fmaxf(-df, 0.f)), f2 = (d == 0) ? 1.f : 0.f;
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ed = fmaf(f2, w2[i], fmaf(f1, w1[i], f0 * w0[i])) + base[i];
        const float x = e[i] + ed;
        if constexpr (sizeof(OutT) == 2) y[i] = silu_f(x); else y[i] = silu_precise(x);
      }
      if constexpr (sizeof(OutT) == 2) {
        uint4 pk;
        pk.x = pack_bf16x2(y[0], y[1]); pk.y = pack_bf16x2(y[2], y[3]);
        pk.z = pack_bf16x2(y[4], y[5]); pk.w = pack_bf16x2(y[6], y[7]);
        *reinterpret_cast<uint4*>(o) = pk;
      } else {
        *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
      }
    }
  }
}

// Lookup-table RPE (use_rpe_net=False, unet.py:326-347): R[net][(b*T+i)*T+j][:] = table[net][bucket(d)][:] with
// d = frame_indices[b][i] - frame_indices[b][j] and the piecewise bucket function of eq. 18 in arXiv 2107.14222,
// evaluated in float32 with the reference's operation order (no FMA contraction) and truncation; a negative bucket
// indexes the table from its end, like the reference's tensor indexing.
__global__ void __launch_bounds__(256) rpe_lookup_kernel(const float* __restrict__ tables, const long long* __restrict__ fi,
                                                          int B, int T, int C, int n_buckets, double alpha, float alpha_f,
                                                          float beta_f, float beta_minus_alpha, float log_ratio,
                                                          float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int C4 = C / 4, rows = B * T * T;
  const long long total = (long long)3 * rows * C4;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(v % C4);
    const long long rr = v / C4;
    const int row = (int)(rr % rows), net = (int)(rr / rows);
    const int bi = row / T, j = row - bi * T, b = bi / T;
    const long long d = fi[bi] - fi[b * T + j];
    const long long ad = d < 0 ? -d : d;
    long long bucket = d;
    if ((double)ad > alpha) {
      const float coef = __fdiv_rn(logf(__fdiv_rn((float)ad, alpha_f)), log_ratio);
      const float val = __fadd_rn(alpha_f, __fmul_rn(coef, beta_minus_alpha));
      const long long mag = (long long)(int)fminf(beta_f, val);
      bucket = d < 0 ? -mag : mag;
    }
    if (bucket < 0) bucket += n_buckets;
    const float4 t = __ldg(reinterpret_cast<const float4*>(tables + ((size_t)net * n_buckets + bucket) * C) + c4);
    reinterpret_cast<float4*>(out + ((size_t)net * rows + row) * C)[c4] = t;
  }
}

}  // namespace
}  // namespace vdm

using namespace vdm;

extern "C" int vdm_gn_stats(const float* src, int32_t C, int32_t n_img, int32_t HW, double* stats,
                            vdm_stream_t stream) {
  return vdm_gn_stats_t(src, VDM_F32, C, n_img, HW, stats, stream);
}

extern "C" int vdm_gn_stats_t(const void* src, int32_t src_dtype, int32_t C, int32_t n_img, int32_t HW, double* stats,
                              vdm_stream_t stream) {
  VDM_REQUIRE(src && stats && C > 0, "gn_stats: NULL pointer");
  VDM_REQUIRE(src_dtype == VDM_F32 || src_dtype == VDM_F16, "gn_stats: source must be fp32 or fp16");
  VDM_REQUIRE(C % 32 == 0 && C <= 4096, "gn_stats: unsupported channel count %d", C);
  const int C4 = C / 4;
  const int rows = C4 >= 256 ? 1 : 256 / C4;
  const int threads = C4 * rows;
  int ppb = 256;
  while (ppb > 16 && (long long)((HW + ppb - 1) / ppb) * n_img < 2LL * num_sms()) ppb >>= 1;
  dim3 grid((HW + ppb - 1) / ppb, n_img);
  if (src_dtype == VDM_F16)
    launch_kernel(gn_stats_kernel<__half>, grid, threads, 0, (cudaStream_t)stream, 1, (const __half*)src, C, HW, ppb, stats);
  else
    launch_kernel(gn_stats_kernel<float>, grid, threads, 0, (cudaStream_t)stream, 1, (const float*)src, C, HW, ppb, stats);
  VDM_AFTER_LAUNCH("gn_stats");
  return 0;
}

extern "C" int vdm_gn_apply(const vdm_gn_apply_args* a, vdm_stream_t stream) {
  const int C = a->C1 + a->C2;
  VDM_REQUIRE(a->src1 && (a->out || (a->out_f32_copy && a->out_mode == 0)) && (a->C2 == 0 || a->src2), "gn_apply: NULL pointer");
  VDM_REQUIRE(C % 32 == 0 && a->C1 % 8 == 0 && a->C2 % 8 == 0 && C <= 2048, "gn_apply: unsupported channels %d+%d",
              a->C1, a->C2);
  VDM_REQUIRE(!a->stats1 || (a->gamma && a->beta), "gn_apply: gamma/beta missing");
  VDM_REQUIRE(!a->stats1 || a->C2 == 0 || a->stats2, "gn_apply: stats2 missing");
  VDM_REQUIRE(!a->stats1 || a->stats_dtype == VDM_F64 || a->stats_dtype == VDM_I64, "gn_apply: bad stats_dtype");
  VDM_REQUIRE(a->out_mode >= 0 && a->out_mode <= 2, "gn_apply: bad out_mode");
  VDM_REQUIRE(a->out_mode != 2 || (a->H % 2 == 0 && a->W % 2 == 0), "gn_apply: parity split needs even H, W");
  VDM_REQUIRE(a->out_mode == 0 || a->out_f32_copy == nullptr, "gn_apply: fp32 copy only with plain output");
  ApplyParams p{a->src1, a->C1, a->src2, a->C2, a->n_img, a->H, a->W, a->stats1, a->stats2,
                a->stats_dtype, a->stats2_dtype, a->gamma, a->beta, a->scale_shift, a->ld_ss, a->silu, a->out_mode,
                a->out, a->out_raw, a->out_f32_copy, a->copy_dtype == VDM_F16 ? 1 : 0, 0, a->wait_done, a->wait_count};
  VDM_REQUIRE(a->wait_done == nullptr || (a->stats1 != nullptr && a->C2 == 0),
              "gn_apply: wait_done takes a single normalised source");
  const int HW = a->H * a->W;
  const int C8 = C / 8;
  int tpb = GN_THREADS;                // VDM_GN_THREADS: block size (profiles/overlap_probe.py)
  if (const char* e = getenv("VDM_GN_THREADS")) tpb = std::min(GN_THREADS, std::max(32, atoi(e)));
  const int rows = C8 >= tpb ? 1 : tpb / C8;
  const int threads = C8 * rows;
  int ppb = rows * 32;
  while (ppb > rows * 8 && (long long)((HW + ppb - 1) / ppb) * a->n_img < 4LL * num_sms()) ppb >>= 1;
  if (const char* e = getenv("VDM_GN_PPB")) ppb = std::max(rows, atoi(e) / rows * rows);
  p.pix_per_block = ppb;
  dim3 grid((HW + ppb - 1) / ppb, a->n_img);
  const size_t smem = 2 * (size_t)C * sizeof(double) + 64 * sizeof(float);
  VDM_REQUIRE(a->src1_dtype != VDM_BF16 || (a->C2 == 0 && a->out_dtype == VDM_BF16),
              "gn_apply: bf16 input needs a single source and bf16 output");
  const bool raw = a->out_raw != nullptr, copy = a->out_f32_copy != nullptr;
  VDM_REQUIRE(a->out_mode == 0 || (!raw && !copy), "gn_apply: extra outputs only with the plain layout");
  VDM_REQUIRE(!(raw && copy), "gn_apply: out_raw and out_f32_copy are mutually exclusive");
#define VDM_GN_LAUNCH(OUT, IN, MODE, RAW, COPY) \
  launch_kernel_ex(gn_apply_kernel<OUT, IN, MODE, RAW, COPY>, grid, threads, smem, (cudaStream_t)stream, 1, \
                   pdl_enabled() || a->wait_done != nullptr, p)
#define VDM_GN_BY_MODE(OUT, IN)                                              \
  do {                                                                       \
    if (a->out_mode == 1) VDM_GN_LAUNCH(OUT, IN, 1, false, false);           \
    else if (a->out_mode == 2) VDM_GN_LAUNCH(OUT, IN, 2, false, false);      \
    else if (raw) VDM_GN_LAUNCH(OUT, IN, 0, true, false);                    \
    else if (copy) VDM_GN_LAUNCH(OUT, IN, 0, false, true);                   \
    else VDM_GN_LAUNCH(OUT, IN, 0, false, false);                            \
  } while (0)
  VDM_REQUIRE(a->src1_dtype != VDM_F16 || a->out_dtype == VDM_BF16, "gn_apply: an fp16 stream feeds bf16 operands only");
  VDM_REQUIRE(a->copy_dtype == VDM_F32 || a->copy_dtype == VDM_F16, "gn_apply: bad copy_dtype");
  if (a->src1_dtype == VDM_BF16) VDM_GN_BY_MODE(__nv_bfloat16, __nv_bfloat16);
  else if (a->src1_dtype == VDM_F16) VDM_GN_BY_MODE(__nv_bfloat16, __half);
  else if (a->out_dtype == VDM_BF16) VDM_GN_BY_MODE(__nv_bfloat16, float);
  else VDM_GN_BY_MODE(float, float);
#undef VDM_GN_BY_MODE
#undef VDM_GN_LAUNCH
  VDM_AFTER_LAUNCH("gn_apply");
  return 0;
}

extern "C" int vdm_gn_coef(const void* stats1, int32_t stats_dtype, int32_t C1, const void* stats2, int32_t stats2_dtype,
                           int32_t C2, int32_t n_img, int32_t HW, const float* gamma, const float* beta,
                           const float* scale_shift, int32_t ld_ss, float* coef, vdm_stream_t stream) {
  const int C = C1 + C2;
  VDM_REQUIRE(stats1 && gamma && beta && coef && (C2 == 0 || stats2), "gn_coef: NULL pointer");
  VDM_REQUIRE(C % 32 == 0 && C <= 2048 && n_img > 0 && HW > 0, "gn_coef: unsupported channels %d+%d", C1, C2);
  VDM_REQUIRE((stats_dtype == VDM_F64 || stats_dtype == VDM_I64) && (C2 == 0 || stats2_dtype == VDM_F64 || stats2_dtype == VDM_I64),
              "gn_coef: bad stats dtype");
  const size_t smem = 2 * (size_t)C * sizeof(double) + 64 * sizeof(float);
  launch_kernel(gn_coef_kernel, n_img, 256, smem, (cudaStream_t)(cudaStream_t)stream, 1, stats1, stats_dtype, C1, stats2, stats2_dtype, C2, HW, gamma,
                                                            beta, scale_shift, ld_ss, reinterpret_cast<float2*>(coef));
  VDM_AFTER_LAUNCH("gn_coef");
  return 0;
}

template <typename OutT, typename IoT>
static int gn_temporal_launch(const IoT* x, int B, int T, int HW, int C, const float* gamma, const float* beta,
                              IoT* out_res, OutT* out_a, cudaStream_t stream) {
  if (C % 128 == 0 && T <= 32) {   // register-resident single pass
    const int l4 = C / 128;                       // float4 lanes per group
    int gps = 32 / l4;                            // groups a warp can hold ...
    while (32 % gps) --gps;                       // ... evened out so every slab has the same number
    const int n_slabs = 32 / gps;
    const long long warps = (long long)B * HW * n_slabs;
    const unsigned grid = (unsigned)((warps + 3) / 4);
#define VDM_GNT(TM) \
  launch_kernel(gn_temporal_regs_kernel<OutT, TM, IoT>, grid, 128, 0, stream, 1, x, B, T, HW, C, gamma, beta, out_res, out_a, gps, n_slabs)
    if (T <= 8) VDM_GNT(8);
    else if (T <= 16) VDM_GNT(16);
    else if (T <= 24) VDM_GNT(24);
    else VDM_GNT(32);
#undef VDM_GNT
    VDM_AFTER_LAUNCH("gn_temporal");
    return 0;
  }
  dim3 grid((HW + 7) / 8, B);
  const size_t smem = 8 * (2 * (size_t)C + 64) * sizeof(float);
  static PerDevice<bool> cfg;
  if (!cfg.get()) {
    cudaFuncSetAttribute(gn_temporal_kernel<OutT, IoT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * (2 * 1024 + 64) * 4);
    cfg.get() = true;
  }
  launch_kernel(gn_temporal_kernel<OutT, IoT>, grid, 256, smem, stream, 1, x, T, HW, C, gamma, beta, out_res, out_a);
  VDM_AFTER_LAUNCH("gn_temporal");
  return 0;
}

extern "C" int vdm_gn_temporal(const float* x, int32_t B, int32_t T, int32_t HW, int32_t C, const float* gamma,
                               const float* beta, float* out_f32, void* out_a, int32_t out_dtype, vdm_stream_t stream) {
  return vdm_gn_temporal_t(x, VDM_F32, B, T, HW, C, gamma, beta, out_f32, out_a, out_dtype, stream);
}

extern "C" int vdm_gn_temporal_t(const void* x, int32_t io_dtype, int32_t B, int32_t T, int32_t HW, int32_t C,
                                 const float* gamma, const float* beta, void* out_res, void* out_a, int32_t out_dtype,
                                 vdm_stream_t stream) {
  VDM_REQUIRE(x && gamma && beta && (out_a || out_res), "gn_temporal: NULL pointer");
  VDM_REQUIRE(C % 32 == 0 && C <= 1024, "gn_temporal: C=%d must be a multiple of 32, <= 1024", C);
  VDM_REQUIRE(io_dtype == VDM_F32 || (io_dtype == VDM_F16 && out_dtype == VDM_BF16),
              "gn_temporal: x / out_res are fp32, or fp16 with a bf16 operand output");
  cudaStream_t st = (cudaStream_t)stream;
  if (io_dtype == VDM_F16)
    return gn_temporal_launch<__nv_bfloat16, __half>((const __half*)x, B, T, HW, C, gamma, beta, (__half*)out_res,
                                                     (__nv_bfloat16*)out_a, st);
  if (out_dtype == VDM_BF16)
    return gn_temporal_launch<__nv_bfloat16, float>((const float*)x, B, T, HW, C, gamma, beta, (float*)out_res,
                                                    (__nv_bfloat16*)out_a, st);
  return gn_temporal_launch<float, float>((const float*)x, B, T, HW, C, gamma, beta, (float*)out_res, (float*)out_a, st);
}

extern "C" int vdm_add_spatial_encoding(const float* h, const float* enc, const float* frame_emb, float* out,
                                        int32_t n_img, int32_t HW, int32_t C, vdm_stream_t stream) {
  return vdm_add_spatial_encoding_t(h, VDM_F32, enc, frame_emb, out, n_img, HW, C, stream);
}

extern "C" int vdm_add_spatial_encoding_t(const void* h, int32_t io_dtype, const float* enc, const float* frame_emb,
                                          void* out, int32_t n_img, int32_t HW, int32_t C, vdm_stream_t stream) {
  VDM_REQUIRE(h && (enc || frame_emb) && out && C % 4 == 0, "add_spatial_encoding: bad arguments");
  VDM_REQUIRE(io_dtype == VDM_F32 || io_dtype == VDM_F16, "add_spatial_encoding: h / out must be fp32 or fp16");
  const long long per4 = (long long)HW * C / 4, total4 = per4 * n_img;
  const int grid = (int)std::min<long long>((total4 + 255) / 256, (long long)num_sms() * 16);
  if (io_dtype == VDM_F16)
    launch_kernel(add_spatial_encoding_kernel<__half>, grid, 256, 0, (cudaStream_t)stream, 1, (const __half*)h, enc, frame_emb,
                  (__half*)out, total4, per4, C / 4);
  else
    launch_kernel(add_spatial_encoding_kernel<float>, grid, 256, 0, (cudaStream_t)stream, 1, (const float*)h, enc, frame_emb,
                  (float*)out, total4, per4, C / 4);
  VDM_AFTER_LAUNCH("add_spatial_encoding");
  return 0;
}

template <int MODE>
static void launch_cond_mix(const float* x, const float* x0, const float* obs, const float* lat, const float* kinda,
                            const float* t, int B, int F, int H, int W, void* a_out, int out_dtype, float* t_frame,
                            float* attn_mask, cudaStream_t stream) {
  const long long M = (long long)B * F * H * W;
  const int grid = (int)((M + 127) / 128);
  if (out_dtype == VDM_BF16)
    launch_kernel(cond_mix_kernel<__nv_bfloat16, MODE>, grid, 128, 0, (cudaStream_t)stream, 1, x, x0, obs, lat, kinda, t, B, F, H, W,
                                                                  (__nv_bfloat16*)a_out, t_frame, attn_mask);
  else
    launch_kernel(cond_mix_kernel<float, MODE>, grid, 128, 0, (cudaStream_t)stream, 1, x, x0, obs, lat, kinda, t, B, F, H, W, (float*)a_out,
                                                          t_frame, attn_mask);
}

extern "C" int vdm_cond_mix(const float* x, const float* x0, const float* obs_mask, const float* latent_mask,
                            const float* kinda_marg_mask, const float* t, int32_t B, int32_t F, int32_t H, int32_t W,
                            int32_t mode, void* a_out, int32_t out_dtype, float* t_frame, float* attn_mask,
                            vdm_stream_t stream) {
  VDM_REQUIRE(x && x0 && obs_mask && latent_mask && kinda_marg_mask && t && a_out && t_frame && attn_mask,
              "cond_mix: NULL pointer");
  VDM_REQUIRE(mode >= 0 && mode <= 2, "cond_mix: mode must be 0 (channel), 1 (duplicate) or 2 (t=0)");
  cudaStream_t s = (cudaStream_t)stream;
  if (mode == 0) launch_cond_mix<0>(x, x0, obs_mask, latent_mask, kinda_marg_mask, t, B, F, H, W, a_out, out_dtype, t_frame, attn_mask, s);
  else if (mode == 1) launch_cond_mix<1>(x, x0, obs_mask, latent_mask, kinda_marg_mask, t, B, F, H, W, a_out, out_dtype, t_frame, attn_mask, s);
  else launch_cond_mix<2>(x, x0, obs_mask, latent_mask, kinda_marg_mask, t, B, F, H, W, a_out, out_dtype, t_frame, attn_mask, s);
  VDM_AFTER_LAUNCH("cond_mix");
  return 0;
}

// One launch for the inputs of a forward: the caller's tensors into the address-stable workspace the CUDA graph reads
// (two big fp32 tensors through float4 copies, the per-frame masks / timesteps / frame indices by the first block).
__global__ void __launch_bounds__(256) stage_inputs_kernel(const float4* __restrict__ x, const float4* __restrict__ x0,
                                                            float4* __restrict__ wx, float4* __restrict__ wx0,
                                                            long long n4, const float* __restrict__ obs,
                                                            const float* __restrict__ lat, const float* __restrict__ kinda,
                                                            float* __restrict__ wobs, float* __restrict__ wlat,
                                                            float* __restrict__ wkinda, int BF, const float* __restrict__ t,
                                                            float* __restrict__ wt, int B, const long long* __restrict__ fi,
                                                            long long* __restrict__ wfi) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 a = __ldg(x + i), b = __ldg(x0 + i);
    wx[i] = a;
    wx0[i] = b;
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < BF; i += blockDim.x) {
      wobs[i] = obs[i];
      wlat[i] = lat[i];
      wkinda[i] = kinda[i];
      if (fi != nullptr) wfi[i] = fi[i];
    }
    for (int i = threadIdx.x; i < B; i += blockDim.x) wt[i] = t[i];
  }
}

extern "C" int vdm_stage_inputs(const float* x, const float* x0, const float* obs_mask, const float* latent_mask,
                                const float* kinda_marg_mask, const float* t, const int64_t* frame_indices, int32_t B,
                                int32_t F, int64_t elems, float* ws_x, float* ws_x0, float* ws_obs, float* ws_lat,
                                float* ws_kinda, float* ws_t, int64_t* ws_fi, vdm_stream_t stream) {
  VDM_REQUIRE(x && x0 && obs_mask && latent_mask && kinda_marg_mask && t && ws_x && ws_x0 && ws_obs && ws_lat &&
                  ws_kinda && ws_t && (frame_indices == nullptr || ws_fi),
              "stage_inputs: NULL pointer");
  VDM_REQUIRE(elems > 0 && elems % 4 == 0 && B > 0 && F > 0, "stage_inputs: elems=%lld must be a positive multiple of 4",
              (long long)elems);
  const long long n4 = elems / 4;
  const int grid = (int)std::min<long long>((n4 + 255) / 256, 4LL * num_sms());
  launch_kernel(stage_inputs_kernel, grid, 256, 0, (cudaStream_t)stream, 1, reinterpret_cast<const float4*>(x),
                reinterpret_cast<const float4*>(x0), reinterpret_cast<float4*>(ws_x), reinterpret_cast<float4*>(ws_x0), n4,
                obs_mask, latent_mask, kinda_marg_mask, ws_obs, ws_lat, ws_kinda, B * F, t, ws_t, B,
                reinterpret_cast<const long long*>(frame_indices), reinterpret_cast<long long*>(ws_fi));
  VDM_AFTER_LAUNCH("stage_inputs");
  return 0;
}

// respace.py:113-119 as one launch: new_ts = timestep_map[clamp(t)] (optionally * 1000 / original_num_steps, fp32).
__global__ void map_timesteps_kernel(const long long* __restrict__ t, const long long* __restrict__ tmap, int n_map,
                                     float scale, float* __restrict__ out, int B) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    long long k = t[i];
    k = k < 0 ? 0 : (k >= n_map ? n_map - 1 : k);
    out[i] = (float)tmap[k] * scale;
  }
}

extern "C" int vdm_map_timesteps(const int64_t* t, const int64_t* timestep_map, int32_t n_map, float scale, float* out,
                                 int32_t B, vdm_stream_t stream) {
  VDM_REQUIRE(t && timestep_map && out && n_map > 0 && B > 0, "map_timesteps: bad arguments");
  launch_kernel(map_timesteps_kernel, (B + 127) / 128, 128, 0, (cudaStream_t)stream, 1,
                reinterpret_cast<const long long*>(t), reinterpret_cast<const long long*>(timestep_map), n_map, scale, out, B);
  VDM_AFTER_LAUNCH("map_timesteps");
  return 0;
}

extern "C" int vdm_timestep_embedding(const float* t_frame, int32_t n, int32_t dim, double max_period, float* out,
                                      vdm_stream_t stream) {
  VDM_REQUIRE(t_frame && out && n > 0 && dim >= 2 && max_period > 0, "timestep_embedding: bad arguments");
  const int total = n * (dim / 2);
  launch_kernel(timestep_embedding_kernel, (total + 127) / 128, 128, 0, (cudaStream_t)(cudaStream_t)stream, 1, t_frame, n, dim,
                                                                                  (float)(-std::log(max_period)), out);
  VDM_AFTER_LAUNCH("timestep_embedding");
  return 0;
}

extern "C" int vdm_rpe_hidden(const float* e_t, int32_t ld_et, const int32_t* et_offsets, int32_t n_blocks,
                              const int64_t* frame_indices, const float* wd, const float* bd, int32_t B, int32_t T,
                              int32_t C, void* out, int32_t out_dtype, vdm_stream_t stream) {
  VDM_REQUIRE(e_t && frame_indices && wd && bd && out, "rpe_hidden: NULL pointer");
  VDM_REQUIRE(C % 8 == 0, "rpe_hidden: C must be a multiple of 8");
  VDM_REQUIRE(n_blocks >= 1, "rpe_hidden: n_blocks must be >= 1");
  VDM_REQUIRE(ld_et % 4 == 0 && ((uintptr_t)e_t & 15) == 0, "rpe_hidden: e_t rows must be 16-byte aligned");
  const int n_nets = 3 * n_blocks;
  const long long total = (long long)n_nets * B * T * (C / 8);
  VDM_REQUIRE(total < (1LL << 31), "rpe_hidden: problem too large");
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  if (out_dtype == VDM_BF16)
    launch_kernel(rpe_hidden_kernel<__nv_bfloat16>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, e_t, ld_et, et_offsets, n_nets, (const long long*)frame_indices,
                                                                            wd, bd, B, T, C, (__nv_bfloat16*)out);
  else
    launch_kernel(rpe_hidden_kernel<float>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, e_t, ld_et, et_offsets, n_nets, (const long long*)frame_indices, wd, bd,
                                                                    B, T, C, (float*)out);
  VDM_AFTER_LAUNCH("rpe_hidden");
  return 0;
}

extern "C" int vdm_rpe_lookup(const float* tables, const int64_t* frame_indices, int32_t B, int32_t T, int32_t C,
                              int32_t n_buckets, double alpha, double beta, double gamma, float* out,
                              vdm_stream_t stream) {
  VDM_REQUIRE(tables && frame_indices && out, "rpe_lookup: NULL pointer");
  VDM_REQUIRE(C % 4 == 0 && B > 0 && T > 0, "rpe_lookup: bad shape");
  VDM_REQUIRE(alpha > 0 && (double)n_buckets >= 2 * beta + 1, "rpe_lookup: table has %d rows, needs 2*beta+1", n_buckets);
  const long long total = (long long)3 * B * T * T * (C / 4);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 8);
  launch_kernel(rpe_lookup_kernel, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, tables, (const long long*)frame_indices, B, T, C, n_buckets,
                                                            alpha, (float)alpha, (float)beta, (float)(beta - alpha),
                                                            (float)log(gamma / alpha), out);
  VDM_AFTER_LAUNCH("rpe_lookup");
  return 0;
}
