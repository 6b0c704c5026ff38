// 3x3 convolution with a handful of output channels (the U-Net's `out` head, unet.py:745-749: C -> 3 or 6).
//
// On the tcgen05 GEMM this layer wastes the tensor core (a 128 x 16 tile keeps 3 columns) and re-reads its
// activation tile from L2 once per tap.  Here a CTA stages the 128-pixel tile WITH its one-pixel halo in shared
// memory once (cp.async, zero fill outside the image = the conv padding), and each of its 8 warps computes
// 16 pixels x 8 channels with mma.sync m16n8k16: ldmatrix takes one address per row, so the nine taps are just
// nine different row addresses into the same halo tile.  Output is written planar (NCHW fp32), the layout the
// sampler consumes.
//
// With `coef` (vdm_gemm_args.a1_coef) the GroupNorm-apply + SiLU in front of the head (unet.py:745-748) happens while
// the tile is staged: x is then the RAW fp16 residual stream, every staged value becomes silu(a * x + b) with the
// per-(image, channel) pairs of vdm_gn_coef, and positions outside the image stay zero (the conv pads the ACTIVATED
// tensor).  The standalone GroupNorm-apply pass over the 64x64 stream and its bf16 copy disappear.
#include "common.cuh"

namespace vdm {
namespace {

constexpr int TILE_PIX = 128;
constexpr int PAD = 8;   // bf16 elements of padding per pixel / weight row: 16 B shifts keep ldmatrix conflict-free

__device__ __forceinline__ void cp16_zfill(void* smem, const void* gmem, bool valid) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a), "l"(gmem), "r"(bytes) : "memory");
}

template <bool XF>   // XF: x is the raw fp16 stream, normalised + activated while it is staged
__global__ void __launch_bounds__(256) conv3x3_small_n_kernel(const uint16_t* __restrict__ x,          // [n_img][H][W][C]
                                                               const __nv_bfloat16* __restrict__ w,   // [N][9*C]
                                                               const float* __restrict__ bias, int H, int W, int TW,
                                                               int C, int N, const float2* __restrict__ coef, int act,
                                                               float* __restrict__ out /* [n_img][N][H*W] */) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  extern __shared__ __align__(16) uint8_t smem[];
  // tile = R image rows x TW columns (column strips keep the halo tile small: several CTAs per SM, little re-reading)
  const int R = TILE_PIX / TW;
  const int Wp = TW + 2, Cp = C + PAD;
  __nv_bfloat16* halo = reinterpret_cast<__nv_bfloat16*>(smem);                 // [(R+2)][Wp][Cp]
  __nv_bfloat16* wsm = halo + (size_t)(R + 2) * Wp * Cp;                        // [9][8][Cp]
  const int strips = W / TW;
  const int tiles_per_img = H * W / TILE_PIX;
  const int img = blockIdx.x / tiles_per_img;
  const int t_in_img = blockIdx.x - img * tiles_per_img;
  const int y0 = (t_in_img / strips) * R;
  const int x0 = (t_in_img % strips) * TW;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C8 = C / 8;

  // ---- stage the halo tile and the weights (index = pixel * C8 + chunk, advanced without divisions)
  const uint16_t* ximg = x + (size_t)img * H * W * C;
  if constexpr (!XF) {
    const int total = (R + 2) * Wp * C8;
    const int step_pix = (int)blockDim.x / C8, step_c8 = (int)blockDim.x % C8;
    const int step_y = step_pix / Wp, step_x = step_pix % Wp;
    int c8 = tid % C8, pix = tid / C8;
    int hy = pix / Wp, hx = pix % Wp;
    for (int i = tid; i < total; i += blockDim.x) {
      const int xx = x0 + hx - 1, yy = y0 + hy - 1;
      const bool ok = xx >= 0 && xx < W && yy >= 0 && yy < H;
      cp16_zfill(halo + (size_t)pix * Cp + c8 * 8, ok ? ximg + ((size_t)yy * W + xx) * C + c8 * 8 : ximg, ok);
      c8 += step_c8; pix += step_pix; hx += step_x; hy += step_y;
      if (c8 >= C8) { c8 -= C8; ++pix; ++hx; }
      if (hx >= Wp) { hx -= Wp; ++hy; }
      if (hx >= Wp) { hx -= Wp; ++hy; }
    }
  } else {
    // the image's (a, b) pairs behind the weights; then batches of four 16-byte loads in flight per thread
    float2* csm = reinterpret_cast<float2*>(wsm + (size_t)9 * 8 * Cp);
    for (int c = tid; c < C; c += blockDim.x) csm[c] = __ldg(coef + (size_t)img * C + c);
    __syncthreads();
    const int total = (R + 2) * Wp * C8;
    for (int i0 = tid; i0 < total; i0 += 4 * blockDim.x) {
      uint4 v[4];
      int pixs[4], c8s[4];
      bool oks[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * blockDim.x;
        const int pix = i / C8, c8 = i - pix * C8;
        const int hy = pix / Wp, hx = pix - hy * Wp;
        const int xx = x0 + hx - 1, yy = y0 + hy - 1;
        oks[u] = i < total && xx >= 0 && xx < W && yy >= 0 && yy < H;
        pixs[u] = pix; c8s[u] = c8;
        v[u] = make_uint4(0u, 0u, 0u, 0u);
        if (oks[u]) v[u] = __ldg(reinterpret_cast<const uint4*>(ximg + ((size_t)yy * W + xx) * C + c8 * 8));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (i0 + u * (int)blockDim.x >= total) break;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (oks[u]) {
          const uint32_t wv[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
          uint32_t r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 f = unpack_f16x2(wv[q]);
            const float4 ab = *reinterpret_cast<const float4*>(csm + c8s[u] * 8 + 2 * q);   // (a0, b0, a1, b1)
            float y0v = fmaf(ab.x, f.x, ab.y), y1v = fmaf(ab.z, f.y, ab.w);
            if (act) { y0v = silu_tanh(y0v); y1v = silu_tanh(y1v); }
            r[q] = pack_bf16x2(y0v, y1v);
          }
          o = make_uint4(r[0], r[1], r[2], r[3]);
        }
        *reinterpret_cast<uint4*>(halo + (size_t)pixs[u] * Cp + c8s[u] * 8) = o;
      }
    }
  }
  for (int i = tid; i < 9 * 8 * C8; i += blockDim.x) {
    const int c8 = i % C8, n = (i / C8) % 8, tap = i / (8 * C8);
    cp16_zfill(wsm + (size_t)(tap * 8 + n) * Cp + c8 * 8, n < N ? w + ((size_t)n * 9 + tap) * C + c8 * 8 : w, n < N);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- warp = 16 consecutive pixels of one image row
  const int p0 = warp * 16;
  const int ry = p0 / TW, rx = p0 - ry * TW;
  float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};   // two chains: even / odd k-steps
  const int a_row = lane & 15, a_k = (lane >> 4) * 8;       // ldmatrix.x4 row / k-half supplied by this lane
  const int b_n = lane >> 2, b_k = (lane & 3) * 2;
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap - dy * 3;              // halo coordinates already include the -1
    const __nv_bfloat16* arow = halo + ((size_t)(ry + dy) * Wp + rx + dx + a_row) * Cp + a_k;
    const __nv_bfloat16* brow = wsm + (size_t)(tap * 8 + b_n) * Cp + b_k;
#pragma unroll 2
    for (int k0 = 0; k0 < C; k0 += 32) {
      uint32_t a[4], a2[4];
      const uint32_t addr = (uint32_t)__cvta_generic_to_shared(arow + k0);
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
      asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                   : "=r"(a2[0]), "=r"(a2[1]), "=r"(a2[2]), "=r"(a2[3]) : "r"(addr + 32u));
      const uint32_t b0 = *reinterpret_cast<const uint32_t*>(brow + k0);
      const uint32_t b1 = *reinterpret_cast<const uint32_t*>(brow + k0 + 8);
      const uint32_t b2 = *reinterpret_cast<const uint32_t*>(brow + k0 + 16);
      const uint32_t b3 = *reinterpret_cast<const uint32_t*>(brow + k0 + 24);
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
          "{%0, %1, %2, %3};"
          : "+f"(acc[0]), "+f"(acc[1]), "+f"(acc[2]), "+f"(acc[3])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      asm volatile(
          "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
          "{%0, %1, %2, %3};"
          : "+f"(acc2[0]), "+f"(acc2[1]), "+f"(acc2[2]), "+f"(acc2[3])
          : "r"(a2[0]), "r"(a2[1]), "r"(a2[2]), "r"(a2[3]), "r"(b2), "r"(b3));
    }
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) acc[e] += acc2[e];
  // accumulator layout: acc[0,1] = (row lane/4, cols 2*(lane%4) + {0,1}); acc[2,3] = row + 8
  const int HW = H * W;
  const int pix = (y0 + ry) * W + x0 + rx + (lane >> 2);
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int n = (lane & 3) * 2 + e;
    if (n < N) {
      const float bv = bias ? bias[n] : 0.f;
      float* o = out + ((size_t)img * N + n) * HW + pix;
      o[0] = acc[e] + bv;
      o[8] = acc[2 + e] + bv;
    }
  }
}

}  // namespace

// Returns -100 if the shape is not handled here (the caller falls back to the tensor-core
// GEMM), 0 on success.
int conv3x3_small_n(const vdm_gemm_args* a, cudaStream_t stream) {
  const int W = a->W, H = a->H, C = a->C1, N = a->N;
  // tile = (128 / TW) rows x TW columns: narrow strips of many rows re-read the least halo (TW 16: 1.41x, 32: 1.59x,
  // 64: 2.06x; measured 0.124 / 0.130 / 0.161 ms on 160 x 64 x 64 x 128) -- the narrowest strip the image height allows
  int TW = W;
  for (int tw = 16; tw <= 64; tw *= 2)
    if (W % tw == 0 && H % (TILE_PIX / tw) == 0) {
      TW = tw;
      break;
    }
  const bool xf = a->a1_coef != nullptr;
  if (!(a->taps == 9 && a->a1_mode == 0 && a->C2 == 0 && a->out_nchw && N <= 8 && C % 32 == 0 && W >= 16 &&
        W % TW == 0 && TILE_PIX % TW == 0 && H % (TILE_PIX / TW) == 0 && a->out_f32 && !a->out_bf16 && !a->residual &&
        !a->rowbias && !a->stats_out && (!xf || a->a1_raw_dtype == VDM_F16) && (xf || a->a1_raw_dtype == 0)))
    return -100;
  const int R = TILE_PIX / TW;
  const size_t smem = ((size_t)(R + 2) * (TW + 2) + 9 * 8) * (C + PAD) * sizeof(__nv_bfloat16) +
                      (xf ? (size_t)C * sizeof(float2) : 0);
  if (smem > 200 * 1024) return -100;
  static PerDevice<size_t> configured[2];
  if (smem > configured[xf].get()) {
    cudaError_t e = xf ? cudaFuncSetAttribute(conv3x3_small_n_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                       : cudaFuncSetAttribute(conv3x3_small_n_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3x3_small_n: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured[xf].get() = smem;
  }
  const int grid = a->n_img * (H * W / TILE_PIX);
  if (xf)
    launch_kernel(conv3x3_small_n_kernel<true>, grid, 256, smem, (cudaStream_t)stream, 1,
                  reinterpret_cast<const uint16_t*>(a->a1), reinterpret_cast<const __nv_bfloat16*>(a->w), a->bias, H, W, TW,
                  C, N, reinterpret_cast<const float2*>(a->a1_coef), (int)a->a1_act, a->out_f32);
  else
    launch_kernel(conv3x3_small_n_kernel<false>, grid, 256, smem, (cudaStream_t)stream, 1,
                  reinterpret_cast<const uint16_t*>(a->a1), reinterpret_cast<const __nv_bfloat16*>(a->w), a->bias, H, W, TW,
                  C, N, static_cast<const float2*>(nullptr), 0, a->out_f32);
  VDM_AFTER_LAUNCH("conv3x3_small_n");
  return 0;
}

}  // namespace vdm
