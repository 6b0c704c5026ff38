// Spatial attention on tensor cores (bf16 operands, fp32 softmax and accumulation): a fused
// flash-style kernel, softmax(q k^T / sqrt(hd)) v per (image, head) with no mask and no RPE
// (unet.py:258-266 -> 477-536 with rpe_* = None).  One CTA = 64 queries x one head x one image,
// 4 warps x 16 query rows; keys/values stream through shared memory in tiles of 64; the
// probabilities never leave registers (accumulator fragments are re-used as the A operand of
// the P.V product).  Sequence lengths here are 64..256 with head_dim 32..128 -- 0.8 % of the
// model's FLOPs -- so this uses warp-level mma.sync rather than a TMEM pipeline.
#include "common.cuh"

namespace vdm {
namespace {

constexpr int BQ = 64, BKV = 64;

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async_16(void* smem, const void* gmem, bool valid) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
  const int bytes = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(a), "l"(gmem), "r"(bytes) : "memory");
}

template <int HD, typename OutT>
__global__ void __launch_bounds__(128) attn_spatial_mma_kernel(const __nv_bfloat16* __restrict__ qkv, int L, int heads,
                                                               OutT* __restrict__ out, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  constexpr int LDS = HD + 8;  // padded row (elements): 16 B skew keeps ldmatrix conflict-free
  constexpr int CH = HD / 8;   // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sKV = sQ + BQ * LDS;   // two buffers of [K tile | V tile]: tile t+1 streams in while tile t is used
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, n = blockIdx.z;
  const int C = heads * HD;
  const size_t row_stride = (size_t)3 * C;
  const __nv_bfloat16* base = qkv + (size_t)n * L * row_stride + h * HD;

  auto load_kv = [&](int k0, int buf) {
    __nv_bfloat16* dK = sKV + (size_t)buf * 2 * BKV * LDS;
    __nv_bfloat16* dV = dK + BKV * LDS;
    for (int idx = tid; idx < BKV * CH; idx += 128) {
      const int r = idx / CH, c = idx - r * CH;
      const bool ok = k0 + r < L;
      const __nv_bfloat16* src = base + (size_t)(ok ? k0 + r : 0) * row_stride + c * 8;
      cp_async_16(dK + r * LDS + c * 8, src + C, ok);
      cp_async_16(dV + r * LDS + c * 8, src + 2 * C, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int idx = tid; idx < BQ * CH; idx += 128) {
    const int r = idx / CH, c = idx - r * CH;
    const bool ok = q0 + r < L;
    cp_async_16(sQ + r * LDS + c * 8, base + (size_t)(ok ? q0 + r : 0) * row_stride + c * 8, ok);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  load_kv(0, 0);
  asm volatile("cp.async.wait_group 1;" ::: "memory");   // Q has landed (the first K/V tile may still be in flight)
  __syncthreads();

  uint32_t qf[HD / 16][4];
#pragma unroll
  for (int kk = 0; kk < HD / 16; ++kk)
    ldmatrix_x4(qf[kk], sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + kk * 16 + (lane >> 4) * 8);

  float o[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  int buf = 0;
  for (int k0 = 0; k0 < L; k0 += BKV, buf ^= 1) {
    if (k0 + BKV < L) {
      load_kv(k0 + BKV, buf ^ 1);   // that buffer was released by the barrier at the end of the previous iteration
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* sK = sKV + (size_t)buf * 2 * BKV * LDS;
    const __nv_bfloat16* sV = sK + BKV * LDS;

    // S = Q K^T for this warp's 16 rows x 64 keys
    float s[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < HD / 16; ++kk) {
#pragma unroll
      for (int nt = 0; nt < BKV / 8; nt += 2) {
        uint32_t b[4];
        ldmatrix_x4(b, sK + (nt * 8 + (lane & 7) + (lane >> 4) * 8) * LDS + kk * 16 + ((lane >> 3) & 1) * 8);
        mma_bf16(s[nt], qf[kk], b[0], b[1]);
        mma_bf16(s[nt + 1], qf[kk], b[2], b[3]);
      }
    }
    if (k0 + BKV > L) {  // ragged last tile: keys beyond L do not exist
#pragma unroll
      for (int nt = 0; nt < BKV / 8; ++nt) {
        const int col = k0 + nt * 8 + 2 * (lane & 3);
        if (col >= L) s[nt][0] = s[nt][2] = -INFINITY;
        if (col + 1 >= L) s[nt][1] = s[nt][3] = -INFINITY;
      }
    }
    // online softmax (rows lane/4 and lane/4 + 8), base-2 exponent with the scale folded in
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float alpha[2], m_new[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      m_new[r] = fmaxf(m_run[r], mx[r] * scale_log2);
      alpha[r] = exp2f(m_run[r] - m_new[r]);
      m_run[r] = m_new[r];
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int nt = 0; nt < BKV / 8; ++nt) {
      s[nt][0] = exp2f(fmaf(s[nt][0], scale_log2, -m_new[0]));
      s[nt][1] = exp2f(fmaf(s[nt][1], scale_log2, -m_new[0]));
      s[nt][2] = exp2f(fmaf(s[nt][2], scale_log2, -m_new[1]));
      s[nt][3] = exp2f(fmaf(s[nt][3], scale_log2, -m_new[1]));
      l_run[0] += s[nt][0] + s[nt][1];
      l_run[1] += s[nt][2] + s[nt][3];
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o[i][0] *= alpha[0]; o[i][1] *= alpha[0];
      o[i][2] *= alpha[1]; o[i][3] *= alpha[1];
    }
    // O += P V : accumulator fragments of S become the A operand
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nt = 0; nt < HD / 8; nt += 2) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, sV + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + nt * 8 + (lane >> 4) * 8);
        mma_bf16(o[nt], pa, b[0], b[1]);
        mma_bf16(o[nt + 1], pa, b[2], b[3]);
      }
    }
    __syncthreads();   // all warps are done with this K/V buffer before the next prefetch overwrites it
  }
  // finalize: row sums live as per-lane partials within each quad
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  OutT* obase = out + (size_t)n * L * C + h * HD + 2 * (lane & 3);
#pragma unroll
  for (int nt = 0; nt < HD / 8; ++nt) {
    if (r0 < L) {
      OutT* p = obase + (size_t)r0 * C + nt * 8;
      if constexpr (sizeof(OutT) == 2) *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(o[nt][0] * inv0, o[nt][1] * inv0);
      else *reinterpret_cast<float2*>(p) = make_float2(o[nt][0] * inv0, o[nt][1] * inv0);
    }
    if (r1 < L) {
      OutT* p = obase + (size_t)r1 * C + nt * 8;
      if constexpr (sizeof(OutT) == 2) *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(o[nt][2] * inv1, o[nt][3] * inv1);
      else *reinterpret_cast<float2*>(p) = make_float2(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

template <int HD, typename OutT>
int launch(const void* qkv, int n_img, int L, int heads, void* out, cudaStream_t stream) {
  const size_t smem = (size_t)(BQ + 4 * BKV) * (HD + 8) * sizeof(__nv_bfloat16);
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_mma_kernel<HD, OutT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_error("attn_spatial_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  dim3 grid((L + BQ - 1) / BQ, heads, n_img);
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  launch_kernel(attn_spatial_mma_kernel<HD, OutT>, grid, 128, smem, (cudaStream_t)stream, 1, (const __nv_bfloat16*)qkv, L, heads, (OutT*)out,
                                                                scale_log2);
  VDM_AFTER_LAUNCH("attn_spatial_tc");
  return 0;
}

template <typename OutT>
int dispatch(const void* qkv, int n_img, int L, int heads, int hd, void* out, cudaStream_t stream) {
  switch (hd) {
    case 32: return launch<32, OutT>(qkv, n_img, L, heads, out, stream);
    case 64: return launch<64, OutT>(qkv, n_img, L, heads, out, stream);
    case 96: return launch<96, OutT>(qkv, n_img, L, heads, out, stream);
    case 128: return launch<128, OutT>(qkv, n_img, L, heads, out, stream);
  }
  set_error("attn_spatial_tc: head_dim=%d not in {32, 64, 96, 128}", hd);
  return -1;
}

}  // namespace

int attn_spatial_tc(const void* qkv, int n_img, int L, int heads, int hd, void* out_a, int out_dtype,
                    cudaStream_t stream) {
  if (out_dtype == VDM_BF16) return dispatch<__nv_bfloat16>(qkv, n_img, L, heads, hd, out_a, stream);
  return dispatch<float>(qkv, n_img, L, heads, hd, out_a, stream);
}

}  // namespace vdm
