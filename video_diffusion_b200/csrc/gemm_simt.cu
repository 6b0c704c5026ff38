// fp32 SIMT implicit-GEMM convolution / linear: the reference-accuracy path (fp32 multiply,
// fp32 accumulate; no tensor cores) with the same fused epilogue as gemm_tc.cu.  Also used for
// the small fp32 GEMMs of the timestep-embedding path in every mode.
#include "common.cuh"

namespace vdm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256;

struct SimtParams {
  int M, N, K;
  int taps, C1, C2, mode;
  int H, W, HW;       // output geometry
  int srcH, srcW;     // A1 geometry
  const float* a1;
  const float* a2;
  const float* w;
  const float* bias;
  const float* rowbias; int ld_rowbias;
  const float* residual; int ld_res;
  float* out_f32; __nv_bfloat16* out_bf16; float* out_silu;
  int ld_out, ld_out_bf16, out_nchw;
};

__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(const SimtParams p) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // loader mapping: 4 consecutive k of one row
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const int am = m0 + lrow;
  int img = 0, y = 0, x = 0;
  const bool a_row_ok = am < p.M;
  if (a_row_ok) {
    img = am / p.HW;
    const int rem = am - img * p.HW;
    y = rem / p.W;
    x = rem - y * p.W;
  }
  const int bn = n0 + lrow;
  const bool b_row_ok = bn < p.N;
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4] = {};
  const int k1 = p.taps * p.C1;
  for (int k0 = 0; k0 < p.K; k0 += BK) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_row_ok) {
      const int k = k0 + lk;
      if (k < k1) {
        const int tap = k / p.C1;
        const int c = k - tap * p.C1;
        int iy = y, ix = x;
        bool ok = true;
        if (p.taps == 9) {
          const int r = tap / 3, s = tap - r * 3;
          if (p.mode == 0) {
            iy = y + r - 1; ix = x + s - 1;
            ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
          } else if (p.mode == 1) {
            iy = 2 * y + r - 1; ix = 2 * x + s - 1;
            ok = iy >= 0 && iy < p.srcH && ix >= 0 && ix < p.srcW;
          } else {
            const int uy = y + r - 1, ux = x + s - 1;
            ok = uy >= 0 && uy < p.H && ux >= 0 && ux < p.W;
            iy = uy >> 1; ix = ux >> 1;
          }
        }
        if (ok) av = *reinterpret_cast<const float4*>(p.a1 + ((size_t)(img * p.srcH + iy) * p.srcW + ix) * p.C1 + c);
      } else {
        av = *reinterpret_cast<const float4*>(p.a2 + (size_t)am * p.C2 + (k - k1));
      }
    }
    float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (b_row_ok) bv = *reinterpret_cast<const float4*>(p.w + (size_t)bn * p.K + k0 + lk);
    __syncthreads();
    As[lk + 0][lrow] = av.x; As[lk + 1][lrow] = av.y; As[lk + 2][lrow] = av.z; As[lk + 3][lrow] = av.w;
    Bs[lk + 0][lrow] = bv.x; Bs[lk + 1][lrow] = bv.y; Bs[lk + 2][lrow] = bv.z; Bs[lk + 3][lrow] = bv.w;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const float br[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
    const int mi = m / p.HW, mp = m - mi * p.HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      if (p.rowbias) v += p.rowbias[(size_t)mi * p.ld_rowbias + n];
      if (p.residual) v += p.residual[(size_t)m * p.ld_res + n];
      if (p.out_nchw) {
        p.out_f32[((size_t)mi * p.N + n) * p.HW + mp] = v;
      } else {
        if (p.out_f32) p.out_f32[(size_t)m * p.ld_out + n] = v;
        if (p.out_bf16) p.out_bf16[(size_t)m * p.ld_out_bf16 + n] = __float2bfloat16_rn(v);
        if (p.out_silu) p.out_silu[(size_t)m * p.ld_out + n] = silu_precise(v);
      }
    }
  }
}

}  // namespace

int gemm_simt(const vdm_gemm_args* a, cudaStream_t stream) {
  VDM_REQUIRE(a->taps == 1 || a->taps == 9, "gemm_simt: taps must be 1 or 9");
  VDM_REQUIRE(a->C1 > 0 && a->C1 % BK == 0 && a->C2 % BK == 0, "gemm_simt: C1=%d, C2=%d must be multiples of 16",
              a->C1, a->C2);
  VDM_REQUIRE(a->a1_mode >= 0 && a->a1_mode <= 2, "gemm_simt: bad a1_mode");
  VDM_REQUIRE(a->C2 == 0 || a->a2 != nullptr, "gemm_simt: a2 is NULL");
  VDM_REQUIRE(a->stats_out == nullptr, "gemm_simt: stats_out is only produced by the bf16 tensor-core kernel");
  VDM_REQUIRE(a->lda1 == 0 && a->w_group_tiles == 0, "gemm_simt: lda1 / grouped weights are bf16-kernel features");
  SimtParams p{};
  p.M = a->n_img * a->H * a->W;
  p.N = a->N;
  p.K = a->taps * a->C1 + a->C2;
  p.taps = a->taps; p.C1 = a->C1; p.C2 = a->C2; p.mode = a->taps == 9 ? a->a1_mode : 0;
  p.H = a->H; p.W = a->W; p.HW = a->H * a->W;
  p.srcH = a->H; p.srcW = a->W;
  if (p.mode == 1) { p.srcH = 2 * a->H; p.srcW = 2 * a->W; }
  if (p.mode == 2) { p.srcH = a->H / 2; p.srcW = a->W / 2; }
  p.a1 = (const float*)a->a1; p.a2 = (const float*)a->a2; p.w = (const float*)a->w;
  p.bias = a->bias; p.rowbias = a->rowbias; p.ld_rowbias = a->ld_rowbias;
  p.residual = a->residual; p.ld_res = a->ld_res;
  p.out_f32 = a->out_f32; p.out_bf16 = (__nv_bfloat16*)a->out_bf16; p.out_silu = a->out_silu_f32;
  p.ld_out = a->ld_out; p.ld_out_bf16 = a->ld_out_bf16; p.out_nchw = a->out_nchw;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM);
  launch_kernel(gemm_simt_kernel, grid, THREADS, 0, (cudaStream_t)stream, 1, p);
  VDM_AFTER_LAUNCH("gemm_simt");
  return 0;
}

}  // namespace vdm
