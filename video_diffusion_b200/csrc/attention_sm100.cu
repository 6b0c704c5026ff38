// Spatial attention on the 5th-generation tensor cores: softmax(q k^T / sqrt(hd)) v per (image, head), no mask and no
// RPE (unet.py:258-266 -> 477-536 with rpe_* = None), as ONE persistent tcgen05 / TMEM / TMA kernel.
//
//   tile        128 query rows x one head.  L = 256: an image is two tiles that share its keys and values;
//               L = 128: one tile = one image; L = 64: one tile = TWO images (block-diagonal key mask).
//   S = Q K^T   tcgen05.mma (kind::f16, bf16 x bf16 -> fp32), Q and K as K-major 128B-swizzled TMA tiles, S in TMEM
//               (128 lanes x KEYS columns): lane = query, column = key.
//   softmax     one thread per query row reads its S row with tcgen05.ld (two passes: max, then exp2 / sum) and
//               writes P back INTO TENSOR MEMORY as packed bf16 pairs (tcgen05.st) over the columns it has consumed:
//               the probabilities never touch shared memory or registers of another thread.
//   O = P V     tcgen05.mma with the A operand read from tensor memory (P) and V as an MN-major 128B-swizzled
//               shared-memory operand -- exactly the image TMA writes for a [key][64 channels] box, no transpose.
//   epilogue    the same thread reads its O row (fp32), scales by 1 / row sum and stores bf16.
//
// Warp roles (320 threads, one CTA per SM): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer, warps 2-5 and
// 6-9 two softmax / epilogue groups.  Tensor memory holds two 256-column buffers (S -> P -> O of one tile each); tile
// t uses buffer t & 1 and softmax group t & 1, so the MMAs of one tile overlap the exponentials of the other (the
// kernel is bound by the 16 ex2 per clock of an SM, not by the tensor pipe).  K is released as soon as the S MMAs
// of its tiles have retired, V after the P V MMAs: the next group's loads overlap the current group's softmax.
#include "sm100_ptx.cuh"

namespace vdm {
namespace {

using namespace ptx;

template <int HD, int L>
struct AttnCfg {
  static_assert(L == 64 || L == 128 || L == 256, "sequence length");
  static_assert(HD % 32 == 0 && HD >= 32 && HD <= 128, "head dim");
  static constexpr int KEYS = L >= 128 ? L : 128;            // keys (= rows) of one K/V group
  static constexpr int TPK = KEYS / 128;                     // query tiles per K/V group
  static constexpr int ATOMS = (HD + 63) / 64;               // 64-column TMA boxes per row (hd 96 over-fetches 32)
  static constexpr int BOX_BYTES = 128 * 128;                // one box: 128 rows x 128 B
  static constexpr int Q_BYTES = ATOMS * BOX_BYTES;
  static constexpr int KV_ATOM_BYTES = KEYS * 128;           // all keys of one 64-column atom
  static constexpr int KV_BYTES = ATOMS * KV_ATOM_BYTES;
  static constexpr int Q_OFF = 0;
  static constexpr int K_OFF = 2 * Q_BYTES;
  static constexpr int V_OFF = K_OFF + KV_BYTES;
  static constexpr int BAR_OFF = V_OFF + KV_BYTES;
  static constexpr int NUM_BARS = 16;
  static constexpr int TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + 1024;
  static constexpr int TMEM_BUF = 256;                       // columns per tile buffer: S [0, KEYS), P [0, KEYS/2), O [128, 128 + HD)
  static constexpr int O_COL = 128;
};

constexpr int kThreads = 320;

template <int HD, int L>
__global__ void __launch_bounds__(kThreads, 1)
    attn_spatial_sm100_kernel(const __grid_constant__ CUtensorMap tm_qkv, __nv_bfloat16* __restrict__ out, int m_rows,
                              int heads, int n_items, float scale_log2) {
  using Cfg = AttnCfg<HD, L>;
  constexpr int KEYS = Cfg::KEYS, TPK = Cfg::TPK, ATOMS = Cfg::ATOMS;
  pdl_launch_dependents_persistent();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + Cfg::BAR_OFF;
  auto q_full = [&](int b) { return bar_base + 8u * b; };
  auto q_empty = [&](int b) { return bar_base + 8u * (2 + b); };
  auto s_full = [&](int b) { return bar_base + 8u * (4 + b); };
  auto p_full = [&](int b) { return bar_base + 8u * (6 + b); };
  auto o_full = [&](int b) { return bar_base + 8u * (8 + b); };
  auto o_empty = [&](int b) { return bar_base + 8u * (10 + b); };
  const uint32_t k_full = bar_base + 8u * 12, k_empty = bar_base + 8u * 13;
  const uint32_t v_full = bar_base + 8u * 14, v_empty = bar_base + 8u * 15;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + Cfg::BAR_OFF + Cfg::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int C = heads * HD;

  if (threadIdx.x == 0) {
    for (int b = 0; b < 2; ++b) {
      mbar_init(q_full(b), 1);
      mbar_init(q_empty(b), 1);
      mbar_init(s_full(b), 1);
      mbar_init(p_full(b), 4);       // one arrive per softmax warp
      mbar_init(o_full(b), 1);
      mbar_init(o_empty(b), 4);
    }
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_qkv)) : "memory");
      int t = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int g = item / heads, h = item - g * heads;
        const int row0 = g * KEYS;
        mbar_wait(k_empty, (uint32_t)(it & 1) ^ 1u, 0);
        mbar_expect_tx(k_full, Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < ATOMS; ++a)
#pragma unroll
          for (int kb = 0; kb < TPK; ++kb)
            tma_load_2d(smem_base + Cfg::K_OFF + a * Cfg::KV_ATOM_BYTES + kb * Cfg::BOX_BYTES, &tm_qkv, k_full,
                        C + h * HD + a * 64, row0 + kb * 128);
        for (int tt = 0; tt < TPK; ++tt, ++t) {
          const int b = t & 1;
          mbar_wait(q_empty(b), (uint32_t)((t >> 1) & 1) ^ 1u, 1);
          mbar_expect_tx(q_full(b), Cfg::Q_BYTES);
#pragma unroll
          for (int a = 0; a < ATOMS; ++a)
            tma_load_2d(smem_base + Cfg::Q_OFF + b * Cfg::Q_BYTES + a * Cfg::BOX_BYTES, &tm_qkv, q_full(b),
                        h * HD + a * 64, row0 + tt * 128);
        }
        mbar_wait(v_empty, (uint32_t)(it & 1) ^ 1u, 2);
        mbar_expect_tx(v_full, Cfg::KV_BYTES);
#pragma unroll
        for (int a = 0; a < ATOMS; ++a)
#pragma unroll
          for (int kb = 0; kb < TPK; ++kb)
            tma_load_2d(smem_base + Cfg::V_OFF + a * Cfg::KV_ATOM_BYTES + kb * Cfg::BOX_BYTES, &tm_qkv, v_full,
                        2 * C + h * HD + a * 64, row0 + kb * 128);
      }
      pdl_trigger_late();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_s = idesc_bf16(128, KEYS);
      constexpr uint32_t idesc_o = idesc_bf16(128, HD, /*b_mn=*/true);
      auto issue_s = [&](int t) {      // S = Q K^T of tile t into TMEM buffer t & 1
        const int b = t & 1;
        const uint32_t u = (uint32_t)(t >> 1) & 1u;
        mbar_wait(q_full(b), u, 4);
        mbar_wait(o_empty(b), u ^ 1u, 5);          // the tile two back has left this TMEM buffer
        tc_fence_after();
        const uint32_t q_addr = smem_base + Cfg::Q_OFF + b * Cfg::Q_BYTES;
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint64_t a_desc = smem_desc_k_sw128(q_addr + (kk >> 2) * Cfg::BOX_BYTES + (kk & 3) * 32);
          const uint64_t b_desc =
              smem_desc_k_sw128(smem_base + Cfg::K_OFF + (kk >> 2) * Cfg::KV_ATOM_BYTES + (kk & 3) * 32);
          umma_ss(tmem_base + b * Cfg::TMEM_BUF, a_desc, b_desc, idesc_s, kk > 0);
        }
        umma_commit(s_full(b));
        umma_commit(q_empty(b));
      };
      auto issue_pv = [&](int t) {     // O = P V of tile t: P from tensor memory, V MN-major from shared memory
        const int b = t & 1;
        const uint32_t u = (uint32_t)(t >> 1) & 1u;
        mbar_wait(p_full(b), u, 7);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < KEYS / 16; ++kk) {
          const uint64_t v_desc = smem_desc_mn_sw128(smem_base + Cfg::V_OFF + kk * 2048, Cfg::KV_ATOM_BYTES);
          umma_ts(tmem_base + b * Cfg::TMEM_BUF + Cfg::O_COL, tmem_base + b * Cfg::TMEM_BUF + kk * 8, v_desc, idesc_o,
                  kk > 0);
        }
        umma_commit(o_full(b));
      };
      if constexpr (TPK == 2) {
        // both tiles of an image: S_a, S_b, then P V of each as its softmax group finishes
        int it = 0;
        for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
          mbar_wait(k_full, (uint32_t)(it & 1), 3);
          issue_s(2 * it);
          issue_s(2 * it + 1);
          umma_commit(k_empty);
          mbar_wait(v_full, (uint32_t)(it & 1), 6);
          issue_pv(2 * it);
          issue_pv(2 * it + 1);
          umma_commit(v_empty);
        }
      } else {
        // one tile per K/V group: S of tile t + 1 is issued before P V of tile t, so it runs under tile t's softmax
        const int n_mine = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        for (int t = 0; t <= n_mine; ++t) {
          if (t < n_mine) {
            mbar_wait(k_full, (uint32_t)(t & 1), 3);
            issue_s(t);
            umma_commit(k_empty);
          }
          if (t > 0) {
            mbar_wait(v_full, (uint32_t)((t - 1) & 1), 6);
            issue_pv(t - 1);
            umma_commit(v_empty);
          }
        }
      }
    }
  } else {
    // ===================== softmax + epilogue: group wg owns the tiles with t & 1 == wg =====================
    const int wg = (warp - 2) >> 2;
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                // query row of the tile
    const uint32_t taddr = tmem_base + (uint32_t)(wg * Cfg::TMEM_BUF) + ((uint32_t)(q * 32) << 16);
    // L = 64: the tile holds two images; a query sees the 64 keys of its own image only (warp-uniform)
    const int c_lo = (L < 128) ? (row / L) * (L / 32) : 0;
    const int c_hi = (L < 128) ? c_lo + L / 32 : KEYS / 32;
    int t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int g = item / heads, h = item - g * heads;
      for (int tt = 0; tt < TPK; ++tt, ++t) {
        if ((t & 1) != wg) continue;
        const uint32_t u = (uint32_t)(t >> 1) & 1u;
        mbar_wait(s_full(wg), u, 8);
        tc_fence_after();
        // Both passes keep one TMEM load in flight while the previous chunk is consumed (two register buffers).
        // ---- pass 1: row maximum
        float mx = -INFINITY;
        {
          uint32_t va[32], vb[32];
          auto fold = [&](const uint32_t (&v)[32]) {
            float m0 = mx, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              m0 = fmaxf(m0, __uint_as_float(v[i]));
              m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
              m2 = fmaxf(m2, __uint_as_float(v[i + 2]));
              m3 = fmaxf(m3, __uint_as_float(v[i + 3]));
            }
            mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          };
          tmem_ld_32(taddr + c_lo * 32, va);
#pragma unroll 1
          for (int c = c_lo; c < c_hi; c += 2) {
            tmem_ld_wait();
            if (c + 1 < c_hi) tmem_ld_32(taddr + (c + 1) * 32, vb);
            fold(va);
            if (c + 1 < c_hi) {
              tmem_ld_wait();
              if (c + 2 < c_hi) tmem_ld_32(taddr + (c + 2) * 32, va);
              fold(vb);
            }
          }
        }
        const float msc = mx * scale_log2;
        // ---- pass 2: p = 2^(s * scale - max * scale), row sum, P -> TMEM as bf16 pairs (over consumed S columns)
        float sum = 0.f;
        {
          uint32_t va[32], vb[32];
          float s0 = 0.f, s1 = 0.f;
          auto emit = [&](const uint32_t (&v)[32], int c) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * i]), scale_log2, -msc));
              const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * i + 1]), scale_log2, -msc));
              s0 += p0;
              s1 += p1;
              pk[i] = pack_bf16x2(p0, p1);
            }
            tmem_st_16(taddr + c * 16, pk);
          };
          tmem_ld_32(taddr + c_lo * 32, va);
#pragma unroll 1
          for (int c = c_lo; c < c_hi; c += 2) {
            tmem_ld_wait();
            if (c + 1 < c_hi) tmem_ld_32(taddr + (c + 1) * 32, vb);
            emit(va, c);
            if (c + 1 < c_hi) {
              tmem_ld_wait();
              if (c + 2 < c_hi) tmem_ld_32(taddr + (c + 2) * 32, va);
              emit(vb, c + 1);
            }
          }
          sum = s0 + s1;
          if constexpr (L < 128) {       // keys of the other image: P = 0 (after this row's own S columns are consumed)
            uint32_t zero[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) zero[i] = 0u;
            for (int c = 0; c < KEYS / 32; ++c)
              if (c < c_lo || c >= c_hi) tmem_st_16(taddr + c * 16, zero);
          }
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full(wg));
        // ---- epilogue: O row / sum -> bf16
        const float inv = 1.0f / sum;
        const int grow = g * KEYS + tt * 128 + row;
        __nv_bfloat16* const orow = out + (size_t)grow * C + h * HD;
        mbar_wait(o_full(wg), u, 9);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < HD / 32; ++c) {
          uint32_t v[32];
          tmem_ld_32(taddr + Cfg::O_COL + c * 32, v);
          tmem_ld_wait();
          if (grow < m_rows) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v[8 * i]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
              w.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
              w.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
              w.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
              *reinterpret_cast<uint4*>(orow + c * 32 + i * 8) = w;
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(o_empty(wg));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

template <int HD, int L>
int launch(const void* qkv, int n_img, int heads, void* out, cudaStream_t stream) {
  using Cfg = AttnCfg<HD, L>;
  static_assert(Cfg::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(attn_spatial_sm100_kernel<HD, L>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::TOTAL);
    if (e != cudaSuccess) {
      set_error("attn_spatial_sm100: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  const int C = heads * HD;
  const int64_t m_rows = (int64_t)n_img * L;
  CUtensorMap tm;
  if (int rc = encode_2d_bf16(&tm, qkv, (uint64_t)m_rows, (uint64_t)3 * C, (uint64_t)3 * C, 128)) return rc;
  const int n_groups = (int)((m_rows + Cfg::KEYS - 1) / Cfg::KEYS);
  const int n_items = n_groups * heads;
  const int grid = n_items < num_sms() ? n_items : num_sms();
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  launch_kernel(attn_spatial_sm100_kernel<HD, L>, grid, kThreads, Cfg::TOTAL, stream, 1, tm, (__nv_bfloat16*)out,
                (int)m_rows, heads, n_items, scale_log2);
  VDM_AFTER_LAUNCH("attn_spatial_sm100");
  return 0;
}

template <int HD>
int dispatch_l(const void* qkv, int n_img, int L, int heads, void* out, cudaStream_t stream) {
  switch (L) {
    case 64: return launch<HD, 64>(qkv, n_img, heads, out, stream);
    case 128: return launch<HD, 128>(qkv, n_img, heads, out, stream);
    case 256: return launch<HD, 256>(qkv, n_img, heads, out, stream);
  }
  return -100;
}

}  // namespace

// bf16 q, k, v -> bf16 output on the tcgen05 kernel.  Returns -100 when the shape is outside its envelope (the caller
// then uses the mma.sync kernel): L in {64, 128, 256}, head_dim in {32, 64, 96, 128}, 16-byte-aligned rows.
bool attn_spatial_sm100_supported(int L, int heads, int hd) {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("VDM_ATTN_SM100");     // 0: keep the mma.sync kernel (profiles / A-B tests)
    on = (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }
  return on && (L == 64 || L == 128 || L == 256) && (hd == 32 || hd == 64 || hd == 96 || hd == 128) &&
         (heads * hd) % 8 == 0;
}

int attn_spatial_sm100(const void* qkv, int n_img, int L, int heads, int hd, void* out_a, cudaStream_t stream) {
  switch (hd) {
    case 32: return dispatch_l<32>(qkv, n_img, L, heads, out_a, stream);
    case 64: return dispatch_l<64>(qkv, n_img, L, heads, out_a, stream);
    case 96: return dispatch_l<96>(qkv, n_img, L, heads, out_a, stream);
    case 128: return dispatch_l<128>(qkv, n_img, L, heads, out_a, stream);
  }
  return -100;
}

}  // namespace vdm
