// 3x3 convolution with a handful of output channels (the U-Net's `out` head, unet.py:745-749: C -> 3 or 6).
//
// On the tcgen05 GEMM this layer wastes the tensor core (a 128 x 16 tile keeps 3 columns) and re-reads its
// activation tile from L2 once per tap.  With so few output channels the layer is a pure streaming problem (one read
// of the activations, 2 bytes per element; the output is 1 % of that), so the kernel is built around reading every
// activation ONCE, straight from global memory into mma.sync fragments -- no shared-memory staging of the activations:
//
//   1. per-pixel products for all nine taps at once:  Y[pixel][tap * N + n] = sum_c x[pixel][c] * w[n][tap][c]
//      -- a plain GEMM over the tile's pixels (its rows plus one row above and below), 9 N <= 72 columns, K = C.
//      mma.sync m16n8k16 leaves the K order free as long as both operands agree, so within a 32-channel block lane
//      (g, q) takes channels q*8 .. q*8+7 of rows g and g+8 as its A fragments of two k-steps: one 16-byte global
//      load per row and block, and the matching B fragment is one 16-byte piece of a weight row (fragment-major copy
//      in shared memory, one LDS.128 per block and column tile).
//   2. Y (fp32, column-major so that both the fragment stores and the gather are bank-conflict free) goes to shared
//      memory; out[n][y][x] = bias[n] + sum_tap Y[(y + dy, x + dx)][tap * N + n], taps outside the image skipped
//      (= the conv's zero padding), written planar (NCHW fp32), the layout the sampler consumes.
//
// A tile is R full image rows, so the only re-read is the two halo rows ((R + 2) / R, mostly L2 hits).
//
// With `coef` (vdm_gemm_args.a1_coef) the GroupNorm-apply + SiLU in front of the head (unet.py:745-748) happens in
// registers between the global load and the mma: x is then the RAW fp16 residual stream and every value becomes
// silu(a * x + b) with the per-(image, channel) pairs of vdm_gn_coef (same arithmetic as gn_apply, so the bf16
// operands are the ones the two-launch path would produce).  The standalone GroupNorm-apply pass over the 64x64
// stream and its bf16 copy disappear.
#include "common.cuh"

namespace vdm {
namespace {

constexpr int HEAD_THREADS = 256;
// coefficient pairs of 8 channels sit 80 bytes apart: the four 16-byte pieces a quarter warp reads (q = 0..3) then
// fall into different banks
constexpr int CSTRIDE = 10;

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// act(a * x + b) of eight fp16 values -> four bf16 pairs; csm8 = the eight staged coefficient pairs of these channels.
// With ACT the staged pairs are (a / 2, b / 2), so h = fma(a / 2, x, b / 2) is EXACTLY half of gn_apply's
// v = fma(a, x, b), and h * (tanh(h) + 1) rounds exactly like its silu_tanh(v) = v * fma(0.5, tanh(v / 2), 0.5)
// (scaling by two commutes with rounding): bit-identical operands for four instructions per value instead of five.
template <bool ACT>
__device__ __forceinline__ uint4 norm_act8(const uint4 v, const float2* csm8) {
  const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
  uint32_t r[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = unpack_f16x2(wv[i]);
    const float4 ab = *reinterpret_cast<const float4*>(csm8 + 2 * i);   // (a0, b0, a1, b1)
    float y0 = fmaf(ab.x, f.x, ab.y), y1 = fmaf(ab.z, f.y, ab.w);
    if (ACT) {
      float t0, t1;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(y0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(y1));
      y0 = y0 * (t0 + 1.0f);
      y1 = y1 * (t1 + 1.0f);
    }
    r[i] = pack_bf16x2(y0, y1);
  }
  return make_uint4(r[0], r[1], r[2], r[3]);
}

// XF: x is the raw fp16 stream, normalised + activated in registers.  KB = C / 32, NT = column tiles of 8 (>= 9 N / 8).
// (Keeping a warp's next pass in flight under the current one -- 32 more registers at C = 128 -- and the maximum
// shared-memory carveout were measured: 74.6 vs 78.2 us / slower for the pre-activated operand, and 89 vs 78 us.)
template <bool XF, int KB, int NT>
__global__ void __launch_bounds__(HEAD_THREADS, 2)
conv3x3_head_kernel(const uint16_t* __restrict__ x,          // [n_img][H][W][C]  bf16 (or raw fp16 with XF)
                    const __nv_bfloat16* __restrict__ w,     // [N][9 * C]
                    const float* __restrict__ bias, int H, int W, int R, int N, int P,
                    const float2* __restrict__ coef, int act, float* __restrict__ out /* [n_img][N][H*W] */) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  constexpr int C = KB * 32;
  extern __shared__ __align__(16) uint8_t smem[];
  float* ysm = reinterpret_cast<float*>(smem);                                   // [9 N][P]
  uint4* bsm = reinterpret_cast<uint4*>(ysm + (size_t)9 * N * P);                // [KB][NT][32 lanes]
  float2* csm = reinterpret_cast<float2*>(bsm + KB * NT * 32);                   // [C / 8][CSTRIDE] (XF only)
  const int tiles_per_img = (H + R - 1) / R;
  const int img = blockIdx.x / tiles_per_img;
  const int y0 = (blockIdx.x - img * tiles_per_img) * R;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;

  // ---- the tile's passes: 16 consecutive pixels of a row each; rows outside the image are never gathered, so the
  //      valid passes are one contiguous range
  const uint16_t* ximg = x + (size_t)img * H * W * C;
  const int hr_lo = y0 == 0 ? 1 : 0, hr_hi = min(R + 2, H - y0 + 1);
  const int mt_hi = hr_hi * W / 16;
  int mt = hr_lo * W / 16 + warp;
  uint4 alo[KB], ahi[KB];
  auto load_pass = [&](int m) {
    const int hp = m * 16, hr = hp / W;
    const uint16_t* row_lo = ximg + ((size_t)(y0 - 1 + hr) * W + (hp - hr * W) + g) * C + q * 8;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      alo[kb] = ldg_stream16(row_lo + kb * 32);
      ahi[kb] = ldg_stream16(row_lo + (size_t)8 * C + kb * 32);
    }
  };

  // ---- weights as B fragments: lane (g, q) of column tile j holds channels q*8..q*8+7 of column 8 j + g
  for (int i = tid; i < KB * NT * 32; i += HEAD_THREADS) {
    const int l = i & 31, j = (i >> 5) % NT, kb = i / (32 * NT);
    const int col = 8 * j + (l >> 2);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (col < 9 * N) {
      const int tap = col / N, n = col - tap * N;
      v = __ldg(reinterpret_cast<const uint4*>(w + ((size_t)n * 9 + tap) * C + kb * 32 + (l & 3) * 8));
    }
    bsm[i] = v;
  }
  if constexpr (XF)
    for (int c = tid; c < C; c += HEAD_THREADS) {
      float2 ab = __ldg(coef + (size_t)img * C + c);
      if (act) { ab.x *= 0.5f; ab.y *= 0.5f; }   // see norm_act8
      csm[(c >> 3) * CSTRIDE + (c & 7)] = ab;
    }
  __syncthreads();

  // ---- Y = x W^T over the tile's rows and its halo rows
  for (; mt < mt_hi; mt += HEAD_THREADS / 32) {
    const int hp = mt * 16;                       // first pixel of the pass, halo-tile numbering ((R + 2) x W)
    uint4 clo[KB], chi[KB];
    load_pass(mt);
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) { clo[kb] = alo[kb]; chi[kb] = ahi[kb]; }
    float acc[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      uint4 lo = clo[kb], hi = chi[kb];
      if constexpr (XF) {
        const float2* c8 = csm + (kb * 4 + q) * CSTRIDE;
        if (act) { lo = norm_act8<true>(lo, c8); hi = norm_act8<true>(hi, c8); }
        else { lo = norm_act8<false>(lo, c8); hi = norm_act8<false>(hi, c8); }
      }
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const uint4 b = bsm[(kb * NT + j) * 32 + lane];
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
            "{%0, %1, %2, %3};"
            : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
            : "r"(lo.x), "r"(hi.x), "r"(lo.y), "r"(hi.y), "r"(b.x), "r"(b.y));
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
            "{%0, %1, %2, %3};"
            : "+f"(acc[j][0]), "+f"(acc[j][1]), "+f"(acc[j][2]), "+f"(acc[j][3])
            : "r"(lo.z), "r"(hi.z), "r"(lo.w), "r"(hi.w), "r"(b.z), "r"(b.w));
      }
    }
    // accumulator layout: acc[.][0,1] = (row g, cols 2 q + {0,1}); acc[.][2,3] = row g + 8
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int col = 8 * j + 2 * q + e;
        if (col < 9 * N) {
          float* yc = ysm + (size_t)col * P + hp + g;
          yc[0] = acc[j][e];
          yc[8] = acc[j][2 + e];
        }
      }
  }
  __syncthreads();

  // ---- gather the nine taps (fixed order), planar store: a thread keeps its image column and walks rows / channels
  const int groups = HEAD_THREADS / W;
  const int grp = tid / W, xx = tid - grp * W;
  if (grp < groups) {
    const bool okl = xx > 0, okr = xx < W - 1;
    const size_t tap_stride = (size_t)N * P;
    for (int r = grp; r < R && y0 + r < H; r += groups) {
      const int y = y0 + r;
      const bool oku = y > 0, okd = y < H - 1;
      const float* yb = ysm + r * W + xx;          // halo row r = image row y - 1
      float* orow = out + (size_t)img * N * H * W + (size_t)y * W + xx;
      for (int n = 0; n < N; ++n) {
        float v[9];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3, dx = tap - dy * 3;
          const bool ok = (dy == 0 ? oku : (dy == 2 ? okd : true)) && (dx == 0 ? okl : (dx == 2 ? okr : true));
          v[tap] = 0.f;
          if (ok) v[tap] = yb[tap * tap_stride + (size_t)n * P + dy * W + dx - 1];
        }
        float s = bias ? bias[n] : 0.f;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) s += v[tap];
        orow[(size_t)n * H * W] = s;
      }
    }
  }
}

struct HeadCfg {
  int H, W, R, N, P;
  size_t smem;
  int grid;
};

template <bool XF, int KB, int NT>
int launch_head(const vdm_gemm_args* a, const HeadCfg& c, cudaStream_t stream) {
  static PerDevice<size_t> configured;
  auto* kernel = conv3x3_head_kernel<XF, KB, NT>;
  if (c.smem > configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
    if (e != cudaSuccess) {
      set_error("conv3x3_small_n: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = c.smem;
  }
  launch_kernel(kernel, c.grid, HEAD_THREADS, c.smem, stream, 1, reinterpret_cast<const uint16_t*>(a->a1),
                reinterpret_cast<const __nv_bfloat16*>(a->w), a->bias, c.H, c.W, c.R, c.N, c.P,
                reinterpret_cast<const float2*>(a->a1_coef), (int)a->a1_act, a->out_f32);
  return 0;
}

template <bool XF, int KB>
int launch_head_nt(const vdm_gemm_args* a, const HeadCfg& c, int nt, cudaStream_t stream) {
  switch (nt) {
    case 4: return launch_head<XF, KB, 4>(a, c, stream);
    case 7: return launch_head<XF, KB, 7>(a, c, stream);
    default: return launch_head<XF, KB, 9>(a, c, stream);
  }
}

template <bool XF>
int launch_head_kb(const vdm_gemm_args* a, const HeadCfg& c, int kb, int nt, cudaStream_t stream) {
  switch (kb) {
    case 2: return launch_head_nt<XF, 2>(a, c, nt, stream);
    case 4: return launch_head_nt<XF, 4>(a, c, nt, stream);
    case 6: return launch_head_nt<XF, 6>(a, c, nt, stream);
    case 8: return launch_head_nt<XF, 8>(a, c, nt, stream);
    default: return -100;
  }
}

}  // namespace

// Returns -100 if the shape is not handled here (the caller falls back to the tensor-core
// GEMM), 0 on success.
int conv3x3_small_n(const vdm_gemm_args* a, cudaStream_t stream) {
  const int W = a->W, H = a->H, C = a->C1, N = a->N;
  const bool xf = a->a1_coef != nullptr;
  if (!(a->dtype == VDM_BF16 && a->io_dtype != VDM_F16 && a->taps == 9 && a->a1_mode == 0 && a->C2 == 0 && a->out_nchw && N >= 1 && N <= 8 && C % 64 == 0 && C <= 256 &&
        W % 16 == 0 && W <= HEAD_THREADS && a->out_f32 && !a->out_bf16 && !a->residual && !a->rowbias && !a->stats_out &&
        (!xf || a->a1_raw_dtype == VDM_F16) && (xf || a->a1_raw_dtype == 0)))
    return -100;
  // R image rows per CTA: as many as keep the fp32 product tile (9 N columns of (R + 2) W pixels) near 70 KB -- two
  // CTAs per SM -- and at most 16; fewer rows re-read more halo ((R + 2) / R)
  HeadCfg c{};
  c.H = H; c.W = W; c.N = N;
  int R = (int)(70 * 1024 / ((size_t)36 * N * W)) - 2;
  R = R < 1 ? 1 : (R > 16 ? 16 : R);
  if (R > H) R = H;
  c.R = R;
  c.P = (R + 2) * W;
  c.P += (36 - c.P % 32) % 32;            // P = 4 (mod 32): the fragment stores of a warp hit 32 different banks
  const int nt = 9 * N <= 32 ? 4 : (9 * N <= 56 ? 7 : 9);
  c.smem = (size_t)9 * N * c.P * sizeof(float) + (size_t)(C / 32) * nt * 32 * sizeof(uint4) +
           (xf ? (size_t)(C / 8) * CSTRIDE * sizeof(float2) : 0);
  if (c.smem > 200 * 1024) return -100;
  c.grid = a->n_img * ((H + R - 1) / R);
  const int rc = xf ? launch_head_kb<true>(a, c, C / 32, nt, stream) : launch_head_kb<false>(a, c, C / 32, nt, stream);
  if (rc != 0) return rc;
  VDM_AFTER_LAUNCH("conv3x3_small_n");
  return 0;
}

}  // namespace vdm
