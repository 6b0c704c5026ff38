// Temporal RPE attention of one block as ONE kernel (bf16 mode): RPE score terms, q.k^T, mask, softmax, P.V and the
// attn.R_v term -- nothing between qkv and the attention output touches global memory (unet.py:357-378, 471-536).
//
//   attn[t][s] = scale * ( q_t . (k_s + Rk[t,s]) + k_s . Rq[s,t] )          (the reference scales q and k*scale)
//   out[t]     = sum_s softmax(attn)[t][s] * (v_s + Rv[t,s])
//
// The R tables depend on (video, frame pair, head) but not on the pixel, q / k / v depend on the pixel: the q.k^T and
// P.V products are per-pixel T x T GEMMs (rows = frames), the three RPE contractions are per-frame GEMMs over the
// pixels of a tile (table rows x pixels).  A CTA owns PT pixels of one (video, head); both families run on mma.sync
// m16n8k16 (sequences of T <= 32 frames are far below a tcgen05 tile) and meet in shared memory:
//
//   P1a  Sk^T[s][pix]  = Rk[t] . Q_t^T      one (t, 16 keys) unit per warp and round             -> S (fp32, smem)
//   P1b  Sq^T[t][pix]  = Rq[s] . K_s^T      one (s, 16 queries) unit per warp and round          -> S +=
//   P2a  warp = pixel: Q K^T + S  -> registers
//   P2b  mask, fp32 softmax, P -> smem (bf16, over the pixel's own S rows), P.V -> O (fp32, over the dead Q / K tiles)
//   P3   out^T[f][pix] = Rv[t]^T . P_t^T + O   one (t, 16 channels) unit per warp and round       -> global, bf16
//
// In the RPE products the R table is the A operand (m = 16 table rows) and the pixels are the n = 8 dimension, so a
// tile of 8 pixels wastes nothing.  vdm_rpe_pack writes the tables FRAGMENT-MAJOR: the four A registers of a lane for
// one k-step are 16 contiguous bytes and a warp's load is 512 contiguous bytes of L2; the B fragments of two k-steps
// are one 16-byte shared-memory load from the q / k / P row of the lane's pixel (the contraction index is permuted
// identically on both sides).  No register shuffling between the loads and the mma, and the fragments of several
// units are requested together so that a warp waits for the L2 round trip once per group.
#include <cstdio>

#include "common.cuh"

#ifndef VDM_TF_G1
#define VDM_TF_G1 3
#endif
#ifndef VDM_TF_G3
#define VDM_TF_G3 8
#endif

namespace vdm {
namespace {

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2(uint32_t (&r)[2], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t a) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  const long long t0 = clock64();
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && clock64() - t0 > 2000000000LL) {   // ~1 s: a protocol bug must fail loudly, never hang the box
      printf("vdm attn_temporal_fused: bulk copies never completed (block %d %d %d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
      __trap();
    }
  }
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion counted on an mbarrier
__device__ __forceinline__ void bulk_copy(uint32_t smem_addr, const void* gmem, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr),
               "l"(gmem), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f2(uint32_t a, float x, float y) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ float lds_f(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_f(uint32_t a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t x) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory"); }

// Row pitch (bytes) of a staged q / k / v row: == 64 (mod 128), so the 16-byte B-fragment reads of two adjacent
// pixel rows (one quarter-warp of the RPE phases) cover all 32 banks; consecutive frames of a pixel are
// TSTR = PT * pitch + 16 apart, so the eight rows of an ldmatrix of the per-pixel phases (consecutive frames, one pixel)
// fall into eight different 16-byte bank groups.
template <int HD>
struct FusedCfg {
  static constexpr int LDSB = (HD * 2) % 128 == 64 ? HD * 2 : HD * 2 + 64;
};
__host__ __device__ inline int fused_region_stride(int T, int TP) {   // per-pixel S region, == 32 (mod 128)
  const int base = T * TP * 4;
  return base + ((32 - base % 128) + 128) % 128;
}

// ---------------------------------------------------------------- table packing
// Fragment-major A operands of mma.m16n8k16 (16-byte vectors v = 4 registers of one lane for one k-step; lane = 4 g + t4;
// register r: table row 16 mt + g + 8 (r & 1), contraction elements c(kk2, r, e) = 8 t4 + 4 kk2 + 2 (r >> 1) + e):
//   which 0 / 1 (Rq / Rk): out[blk][grp][h][mt < 2][p < hd/32][kk2 < 2][lane]   row = second frame index j (zero for
//                          j >= T), contraction over channels f = 32 p + c:  R[blk][(grp*T + j)][h*hd + f] + bias
//   which 2 (Rv)         : out[blk][grp][h][mt < hd/16][kk2 < 2][lane]          row = channel f, contraction over the
//                          second frame index s = c (zero for s >= T):       Rv[blk][(grp*T + s)][h*hd + f] + bias
__global__ void __launch_bounds__(128) rpe_pack_kernel(const float* __restrict__ r_q, const float* __restrict__ r_k,
                                                        const float* __restrict__ r_v, const float* __restrict__ bias,
                                                        long long r_block_stride, int T, int heads, int hd,
                                                        uint4* __restrict__ oq, uint4* __restrict__ ok,
                                                        uint4* __restrict__ ov, long long qk_block_stride,
                                                        long long v_block_stride) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int grp = blockIdx.x, h = blockIdx.y, which = blockIdx.z % 3, blk = blockIdx.z / 3;
  const int C = heads * hd, KP = hd / 32, MV = hd / 16;
  const float* R = (which == 0 ? r_q : (which == 1 ? r_k : r_v)) + (size_t)blk * r_block_stride + (size_t)grp * T * C + h * hd;
  const float* bs = bias ? bias + ((size_t)blk * 3 + which) * C + h * hd : nullptr;
  const int nvec = (which < 2 ? 2 * KP : MV) * 64;
  uint4* out = which < 2 ? (which == 0 ? oq : ok) + (size_t)blk * (qk_block_stride / 8) + ((size_t)grp * heads + h) * nvec
                         : ov + (size_t)blk * (v_block_stride / 8) + ((size_t)grp * heads + h) * nvec;
  for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
    const int lane = v & 31, kk2 = (v >> 5) & 1, g = lane >> 2, t4 = lane & 3;
    uint32_t reg[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c0 = 8 * t4 + 4 * kk2 + 2 * (r >> 1);
      float x[2] = {0.f, 0.f};
      if (which < 2) {
        const int mp = v >> 6, mt = mp / KP, p = mp - mt * KP;
        const int j = 16 * mt + g + 8 * (r & 1), f = 32 * p + c0;
        if (j < T) {
          const float2 t2 = __ldg(reinterpret_cast<const float2*>(R + (size_t)j * C + f));
          x[0] = t2.x + (bs ? bs[f] : 0.f);
          x[1] = t2.y + (bs ? bs[f + 1] : 0.f);
        }
      } else {
        const int mt = v >> 6;
        const int f = 16 * mt + g + 8 * (r & 1);
        const float bf = bs ? bs[f] : 0.f;
#pragma unroll
        for (int e = 0; e < 2; ++e)
          if (c0 + e < T) x[e] = __ldg(R + (size_t)(c0 + e) * C + f) + bf;
      }
      reg[r] = pack_bf16x2(x[0], x[1]);
    }
    out[v] = make_uint4(reg[0], reg[1], reg[2], reg[3]);
  }
}

// ---------------------------------------------------------------- the fused block
// grid = (heads, HW / PT, B); PT warps (warp = pixel in the per-pixel phases).
template <int HD, int PT, int NT>
__global__ void __launch_bounds__(PT * 32, (PT == 8 && HD <= 96) ? 2 : 1)
attn_temporal_fused_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ rq,
                           const __nv_bfloat16* __restrict__ rk, const __nv_bfloat16* __restrict__ rv,
                           const float* __restrict__ mask, int pad_interact, int T, int D, int heads,
                           __nv_bfloat16* __restrict__ out, unsigned long long* __restrict__ trace) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  constexpr int LDSB = FusedCfg<HD>::LDSB, KP = HD / 32, TP = NT * 8, W = PT, NTH = PT * 32;
  constexpr int TSTR = PT * LDSB + 16;
  constexpr int NG = HD / 16;           // 16-channel output groups of the R_v phase
  constexpr int G1 = VDM_TF_G1, G3 = PT == 8 ? VDM_TF_G3 : VDM_TF_G3 / 2;   // units whose R fragments are requested together
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t4 = lane & 3;
  const int h = blockIdx.x, pix0 = blockIdx.y * PT, b = blockIdx.z;   // the heads of a pixel tile run side by side
  const int C = heads * HD;
  const uint32_t tensor_bytes = (uint32_t)T * TSTR;
  const uint32_t sQ = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t sK = sQ + tensor_bytes, sV = sK + tensor_bytes, sS = sV + tensor_bytes, sO = sQ;
  const int RS = fused_region_stride(T, TP);
  const float scale = rsqrtf((float)HD);
  // diagnostics (vdm_attn_temporal_fused_set_trace): cycles per phase, summed over CTAs by thread 0; the last time stamp
  // lives in shared memory (behind the two barriers), not in a register
  const uint32_t tr_slot = sS + (uint32_t)PT * RS + 16;
  auto tr_mark = [&](int slot) {
    if (trace != nullptr && threadIdx.x == 0) {
      const unsigned long long now = (unsigned long long)clock64();
      unsigned long long prev;
      asm volatile("ld.shared.u64 %0, [%1];" : "=l"(prev) : "r"(tr_slot));
      if (slot >= 0) atomicAdd(trace + slot, now - prev);
      asm volatile("st.shared.u64 [%0], %1;" ::"r"(tr_slot), "l"(now) : "memory");
    }
  };
  tr_mark(-1);

  // ---- stage q, k (barrier 0) and v (barrier 1): one bulk copy per (frame, pixel) row of HD bf16
  const uint32_t bar_qk = sS + (uint32_t)PT * RS, bar_v = bar_qk + 8;
  if (tid == 0) {
    mbar_init(bar_qk, 1);
    mbar_init(bar_v, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_qk, 2u * T * PT * HD * 2);
    mbar_expect_tx(bar_v, (uint32_t)T * PT * HD * 2);
  }
  __syncthreads();
  {
    const int rows = T * PT;
    for (int idx = tid; idx < 3 * rows; idx += NTH) {
      const int which = idx / rows, r = idx - which * rows;
      const int t = r / PT, p = r - t * PT;
      const __nv_bfloat16* src = qkv + ((size_t)(b * T + t) * D + pix0 + p) * (3 * C) + which * C + h * HD;
      bulk_copy(sQ + which * tensor_bytes + t * TSTR + p * LDSB, src, HD * 2, which == 2 ? bar_v : bar_qk);
    }
  }

  // ---- P1: RPE score terms.  Unit u = (frame i, 16-row tile mt of the table): S^T tile [16 frames][pixels].
  constexpr int NPT = PT / 8;                    // 8-pixel column tiles
  const int MT1 = (T + 15) >> 4;                 // 16-row tiles of a table that hold live rows
  const int n1 = T * MT1;
  const uint4* rq4 = reinterpret_cast<const uint4*>(rq);
  const uint4* rk4 = reinterpret_cast<const uint4*>(rk);
  const uint4* rv4 = reinterpret_cast<const uint4*>(rv);
  auto load_a1 = [&](uint4 (&af)[KP][2], const uint4* tab, int u) {
    const int i = u / MT1, mt = u - i * MT1;
    const uint4* src = tab + ((((size_t)(b * T + i) * heads + h) * 2 + mt) * KP) * 64 + lane;
#pragma unroll
    for (int p = 0; p < KP; ++p) {
      af[p][0] = __ldg(src + p * 64);
      af[p][1] = __ldg(src + p * 64 + 32);
    }
  };
  // Jobs 0 .. NG1-1 are the Sk groups, NG1 .. 2 NG1 - 1 the Sq groups; every warp runs the same number of jobs -- units
  // past the end of its list are clamped to the last unit (loaded and multiplied like the others, so the independent
  // mma chains of a group interleave without branches; only the stores are predicated).
  const int NG1 = (n1 + W * G1 - 1) / (W * G1);
  uint4 afA[G1][KP][2];
  auto load_job1 = [&](uint4 (&af)[G1][KP][2], int job) {
    const uint4* tab = job < NG1 ? rk4 : rq4;
    const int k0 = (job < NG1 ? job : job - NG1) * G1;
#pragma unroll
    for (int j = 0; j < G1; ++j) load_a1(af[j], tab, min(warp + (k0 + j) * W, n1 - 1));
  };
  auto run_job1 = [&](const uint4 (&af)[G1][KP][2], int job) {
    const bool second = job >= NG1;           // Sq: B = K tile, read-modify-write of the transposed element
    const int k0 = (second ? job - NG1 : job) * G1;
    const uint32_t b_tensor = second ? sK : sQ;
    float c[G1][NPT][4];
#pragma unroll
    for (int j = 0; j < G1; ++j) {
      const int i = min(warp + (k0 + j) * W, n1 - 1) / MT1;
      const uint32_t brow = b_tensor + i * TSTR + g * LDSB + t4 * 16;       // row of pixel g (the n index of B)
#pragma unroll
      for (int np = 0; np < NPT; ++np) c[j][np][0] = c[j][np][1] = c[j][np][2] = c[j][np][3] = 0.f;
#pragma unroll
      for (int p = 0; p < KP; ++p) {
#pragma unroll
        for (int np = 0; np < NPT; ++np) {
          const uint4 bv = lds128(brow + np * 8 * LDSB + p * 64);
          mma16816(c[j][np], af[j][p][0].x, af[j][p][0].y, af[j][p][0].z, af[j][p][0].w, bv.x, bv.y);
          mma16816(c[j][np], af[j][p][1].x, af[j][p][1].y, af[j][p][1].z, af[j][p][1].w, bv.z, bv.w);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < G1; ++j) {
      const int u = warp + (k0 + j) * W;
      if (u < n1) {
        const int i = u / MT1, mt = u - i * MT1;
#pragma unroll
        for (int np = 0; np < NPT; ++np)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = mt * 16 + g + (r >> 1) * 8;          // key s (Sk) / query t' (Sq)
            const int pix = np * 8 + 2 * t4 + (r & 1);
            if (row < T) {
              if (!second) {      // Sk[pix][i][s]: every element of S is written exactly once
                sts_f(sS + pix * RS + (i * TP + row) * 4, c[j][np][r]);
              } else {            // S[pix][t'][i] += K_i . Rq[i][t']
                const uint32_t a = sS + pix * RS + (row * TP + i) * 4;
                sts_f(a, lds_f(a) + c[j][np][r]);
              }
            }
          }
      }
    }
  };
  load_job1(afA, 0);
  // mask values of this lane's key columns and query rows (P2b), fetched while the copies are in flight
  float m_col[NT * 2], m_row[4];
#pragma unroll
  for (int j = 0; j < NT * 2; ++j) {
    const int col = (j >> 1) * 8 + 2 * t4 + (j & 1);
    m_col[j] = col < T ? __ldg(mask + b * T + col) : 0.f;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = (j >> 1) * 16 + g + (j & 1) * 8;
    m_row[j] = r < T ? __ldg(mask + b * T + r) : 0.f;
  }
  mbar_wait(bar_qk, 0);
  tr_mark(0);
  for (int job = 0; job < 2 * NG1; ++job) {
    if (job) load_job1(afA, job);
    if (job == NG1) {
      tr_mark(1);
      __syncthreads();                                // all of Sk is stored before the first Sq unit adds to it
    }
    run_job1(afA, job);
  }
  tr_mark(2);
  __syncthreads();

  // ---- P2a: per pixel (warp), S = Q K^T + (Sk + Sq^T), both 16-frame query tiles
  const int pix = warp;
  float s[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) s[mt][nt][0] = s[mt][nt][1] = s[mt][nt][2] = s[mt][nt][3] = 0.f;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 < T) {
      const int tq = min(mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, T - 1);
      const uint32_t qrow = sQ + tq * TSTR + pix * LDSB + (lane >> 4) * 16;
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        uint32_t qa[4];
        ldsm_x4(qa, qrow + kk * 32);
#pragma unroll
        for (int nt = 0; nt + 1 < NT; nt += 2) {
          const int ts = min(nt * 8 + (lane & 7) + (lane >> 4) * 8, T - 1);
          uint32_t kb[4];
          ldsm_x4(kb, sK + ts * TSTR + pix * LDSB + kk * 32 + ((lane >> 3) & 1) * 16);
          mma16816(s[mt][nt], qa[0], qa[1], qa[2], qa[3], kb[0], kb[1]);
          mma16816(s[mt][nt + 1], qa[0], qa[1], qa[2], qa[3], kb[2], kb[3]);
        }
        if (NT & 1) {
          const int ts = min((NT - 1) * 8 + (lane & 7), T - 1);
          uint32_t kb[2];
          ldsm_x2(kb, sK + ts * TSTR + pix * LDSB + kk * 32 + ((lane >> 3) & 1) * 16);
          mma16816(s[mt][NT - 1], qa[0], qa[1], qa[2], qa[3], kb[0], kb[1]);
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = mt * 16 + g + half * 8;
        if (r < T) {
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const float2 v = lds_f2(sS + pix * RS + (r * TP + nt * 8 + 2 * t4) * 4);
            s[mt][nt][half * 2] += v.x;
            s[mt][nt][half * 2 + 1] += v.y;
          }
        }
      }
    }
  }
  tr_mark(3);
  __syncthreads();     // nobody reads the Q / K tiles or S any more
  mbar_wait(bar_v, 0);

  // ---- P2b: mask, softmax, P -> smem, O = P V -> smem (fp32, over the Q / K tiles)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    if (mt * 16 < T) {
      float inv[2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = mt * 16 + g + half * 8;
        const bool row_ok = r < T;
        const float m_r = m_row[mt * 2 + half];
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = nt * 8 + 2 * t4 + e;
            const float m_s = m_col[nt * 2 + e];
            float al = m_r * m_s;
            if (pad_interact) al += (1.f - m_r) * (1.f - m_s);
            else if (col == r) al = 1.f;
            const float v = (row_ok && col < T && al != 0.f) ? scale * s[mt][nt][half * 2 + e] : -INFINITY;
            s[mt][nt][half * 2 + e] = v;
            mx = fmaxf(mx, v);
          }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
        if (mx == -INFINITY) mx = 0.f;      // padded query rows: all keys masked
        float sum = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float pexp = __expf(s[mt][nt][half * 2 + e] - mx);
            s[mt][nt][half * 2 + e] = pexp;
            sum += pexp;
          }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        inv[half] = sum > 0.f ? 1.f / sum : 0.f;
      }
      uint32_t pk[4][2];                     // bf16 pairs: [key tile][row half]
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int half = 0; half < 2; ++half)
          pk[nt][half] = nt < NT ? pack_bf16x2(s[mt][nt][half * 2] * inv[half], s[mt][nt][half * 2 + 1] * inv[half]) : 0u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = mt * 16 + g + half * 8;
        if (r < T) {
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) sts_u32(sS + pix * RS + r * 64 + (nt * 8 + 2 * t4) * 2, pk[nt][half]);
        }
      }
      float o[HD / 8][4];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        if (kk * 16 < T) {
          const int ts = min(kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, T - 1);
          const uint32_t vrow = sV + ts * TSTR + pix * LDSB + (lane >> 4) * 16;
#pragma unroll
          for (int nt = 0; nt < HD / 8; nt += 2) {
            uint32_t vb[4];
            ldsm_x4_t(vb, vrow + nt * 16);
            mma16816(o[nt], pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1], vb[0], vb[1]);
            mma16816(o[nt + 1], pk[2 * kk][0], pk[2 * kk][1], pk[2 * kk + 1][0], pk[2 * kk + 1][1], vb[2], vb[3]);
          }
        }
      }
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int r = mt * 16 + g + half * 8;
        if (r < T) {
          const uint32_t orow = sO + (uint32_t)(pix * T + r) * (HD * 4) + 2 * t4 * 4;
          const int sw = (r ^ pix) & 3;
#pragma unroll
          for (int nt = 0; nt < HD / 8; ++nt) sts_f2(orow + ((nt ^ sw) * 8) * 4, o[nt][half * 2], o[nt][half * 2 + 1]);
        }
      }
    }
  }
  // ---- P3: out^T[f][pix] = Rv[t]^T . P_t^T + O.  Unit u = (frame t, 16 channels mt)
  const int n3 = T * NG;
  const int NG3 = (n3 + W * G3 - 1) / (W * G3);
  uint4 a3[G3][2];
  auto load_job3 = [&](int job) {
#pragma unroll
    for (int j = 0; j < G3; ++j) {
      const int u = min(warp + (job * G3 + j) * W, n3 - 1);
      const int t = u / NG, mt = u - t * NG;
      const uint4* src = rv4 + (((size_t)(b * T + t) * heads + h) * NG + mt) * 64 + lane;
      a3[j][0] = __ldg(src);
      a3[j][1] = __ldg(src + 32);
    }
  };
  load_job3(0);            // in flight across the barrier
  tr_mark(4);
  __syncthreads();
  const bool even = (g & 1) == 0;
  for (int job = 0; job < NG3; ++job) {
    if (job) load_job3(job);
    float c[G3][NPT][4];
#pragma unroll
    for (int j = 0; j < G3; ++j) {
      const int t = min(warp + (job * G3 + j) * W, n3 - 1) / NG;
#pragma unroll
      for (int np = 0; np < NPT; ++np) {
        const uint4 bv = lds128(sS + (np * 8 + g) * RS + t * 64 + t4 * 16);     // P row of pixel g, frame t
        c[j][np][0] = c[j][np][1] = c[j][np][2] = c[j][np][3] = 0.f;
        mma16816(c[j][np], a3[j][0].x, a3[j][0].y, a3[j][0].z, a3[j][0].w, bv.x, bv.y);
        mma16816(c[j][np], a3[j][1].x, a3[j][1].y, a3[j][1].z, a3[j][1].w, bv.z, bv.w);
      }
    }
#pragma unroll
    for (int j = 0; j < G3; ++j) {
      const int u = warp + (job * G3 + j) * W;
      const bool live = u < n3;
      const int uu = min(u, n3 - 1);
      const int t = uu / NG, mt = uu - t * NG;
#pragma unroll
      for (int np = 0; np < NPT; ++np) {
        // accumulator: (channel 16 mt + g [+ 8], pixel 2 t4 [+ 1]).  Lanes g and g ^ 1 swap halves, so each ends up with
        // two CHANNEL pairs of one pixel: 4-byte stores, and the addend O comes as two 8-byte shared-memory loads.
        const float v1 = __shfl_xor_sync(0xffffffffu, even ? c[j][np][1] : c[j][np][0], 4);
        const float v2 = __shfl_xor_sync(0xffffffffu, even ? c[j][np][3] : c[j][np][2], 4);
        const int pix = np * 8 + 2 * t4 + (even ? 0 : 1);
        const int fa = mt * 16 + g - (even ? 0 : 1);
        const float lo0 = even ? c[j][np][0] : v1, lo1 = even ? v1 : c[j][np][1];
        const float hi0 = even ? c[j][np][2] : v2, hi1 = even ? v2 : c[j][np][3];
        if (live) {
          const int sw = (t ^ pix) & 3;
          const uint32_t orow = sO + (uint32_t)(pix * T + t) * (HD * 4);
          const float2 o0 = lds_f2(orow + ((((fa >> 3) ^ sw) << 3) + (fa & 7)) * 4);
          const float2 o1 = lds_f2(orow + (((((fa >> 3) + 1) ^ sw) << 3) + (fa & 7)) * 4);
          __nv_bfloat16* dst = out + ((size_t)(b * T + t) * D + pix0 + pix) * C + h * HD + fa;
          *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(lo0 + o0.x, lo1 + o0.y);
          *reinterpret_cast<uint32_t*>(dst + 8) = pack_bf16x2(hi0 + o1.x, hi1 + o1.y);
        }
      }
    }
  }
  tr_mark(5);
  if (trace != nullptr && threadIdx.x == 0) atomicAdd(trace + 7, 1ULL);
}

static unsigned long long* g_tf_trace = nullptr;

template <int HD, int PT, int NT>
int launch_fused(const void* qkv, const void* rq, const void* rk, const void* rv, const float* mask, int pad, int B, int T,
                 int D, int heads, void* out, cudaStream_t stream) {
  constexpr int LDSB = FusedCfg<HD>::LDSB, TSTR = PT * LDSB + 16;
  const size_t smem = (size_t)3 * T * TSTR + (size_t)PT * fused_region_stride(T, NT * 8) + 32;
  if (smem > 227 * 1024) {
    set_error("attn_temporal_fused: %zu bytes of shared memory (T=%d, head_dim=%d, %d pixels per CTA)", smem, T, HD, PT);
    return -1;
  }
  static PerDevice<size_t> configured;
  if (configured.get() < smem) {
    cudaError_t e = cudaFuncSetAttribute(attn_temporal_fused_kernel<HD, PT, NT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("attn_temporal_fused: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = smem;
  }
  launch_kernel(attn_temporal_fused_kernel<HD, PT, NT>, dim3(heads, D / PT, B), PT * 32, smem, stream, 1,
                (const __nv_bfloat16*)qkv, (const __nv_bfloat16*)rq, (const __nv_bfloat16*)rk, (const __nv_bfloat16*)rv,
                mask, pad, T, D, heads, (__nv_bfloat16*)out, g_tf_trace);
  VDM_AFTER_LAUNCH("attn_temporal_fused");
  return 0;
}

template <int HD, int PT>
int launch_fused_nt(int t_pad, const void* qkv, const void* rq, const void* rk, const void* rv, const float* mask, int pad,
                    int B, int T, int D, int heads, void* out, cudaStream_t stream) {
  if (t_pad == 24) return launch_fused<HD, PT, 3>(qkv, rq, rk, rv, mask, pad, B, T, D, heads, out, stream);
  return launch_fused<HD, PT, 4>(qkv, rq, rk, rv, mask, pad, B, T, D, heads, out, stream);
}

}  // namespace
}  // namespace vdm

using namespace vdm;

extern "C" int vdm_rpe_pack(const float* r_q, const float* r_k, const float* r_v, const float* bias, int32_t n_blocks,
                            int64_t r_block_stride, int32_t B, int32_t T, int32_t heads, int32_t hd, void* rq, void* rk,
                            void* rv, int64_t qk_block_stride, int64_t v_block_stride, vdm_stream_t stream) {
  VDM_REQUIRE(r_q && r_k && r_v && rq && rk && rv, "rpe_pack: NULL pointer");
  VDM_REQUIRE(T >= 1 && T <= 32, "rpe_pack: T=%d unsupported", T);
  VDM_REQUIRE(hd % 32 == 0 && n_blocks >= 1, "rpe_pack: head_dim must be a multiple of 32");
  const int64_t G = (int64_t)B * T;
  if (qk_block_stride == 0) qk_block_stride = G * heads * hd * 32;     // 2 row tiles x hd/32 x 2 k-steps x 32 lanes x 8
  if (v_block_stride == 0) v_block_stride = G * heads * hd * 32;       // hd/16 row tiles x 2 k-steps x 32 lanes x 8
  VDM_REQUIRE(qk_block_stride % 8 == 0 && v_block_stride % 8 == 0, "rpe_pack: block strides are multiples of 8 elements");
  launch_kernel(rpe_pack_kernel, dim3((unsigned)G, heads, 3 * n_blocks), 128, 0, (cudaStream_t)stream, 1, r_q, r_k, r_v,
                bias, (long long)r_block_stride, T, heads, hd, (uint4*)rq, (uint4*)rk, (uint4*)rv,
                (long long)qk_block_stride, (long long)v_block_stride);
  VDM_AFTER_LAUNCH("rpe_pack");
  return 0;
}

extern "C" void vdm_attn_temporal_fused_set_trace(void* buf) { g_tf_trace = reinterpret_cast<unsigned long long*>(buf); }

extern "C" int64_t vdm_attn_temporal_fused_smem(int32_t T, int32_t hd, int32_t t_pad, int32_t pixels_per_cta) {
  if (T < 1 || T > 32 || (t_pad != 24 && t_pad != 32) || T > t_pad) return -1;
  if (hd != 32 && hd != 64 && hd != 96 && hd != 128) return -1;
  const int pt = pixels_per_cta == 0 ? 8 : pixels_per_cta;
  if (!(pt == 8 || (pt == 16 && hd == 96))) return -1;
  const int ldsb = (hd * 2) % 128 == 64 ? hd * 2 : hd * 2 + 64;
  return (int64_t)3 * T * (pt * ldsb + 16) + (int64_t)pt * fused_region_stride(T, t_pad) + 32;
}

extern "C" int vdm_attn_temporal_fused(const void* qkv, const void* rq, const void* rk, const void* rv, const float* mask,
                                       int32_t allow_pad_interactions, int32_t B, int32_t T, int32_t HW, int32_t heads,
                                       int32_t hd, int32_t t_pad, int32_t pixels_per_cta, void* out,
                                       vdm_stream_t stream) {
  VDM_REQUIRE(qkv && rq && rk && rv && mask && out, "attn_temporal_fused: NULL pointer");
  VDM_REQUIRE(T >= 1 && T <= 32 && (t_pad == 24 || t_pad == 32) && T <= t_pad,
              "attn_temporal_fused: T=%d, t_pad=%d unsupported", T, t_pad);
  int pt = pixels_per_cta;
  if (pt == 0) pt = 8;
  VDM_REQUIRE((pt == 8 || (pt == 16 && hd == 96)) && HW % pt == 0,
              "attn_temporal_fused: %d pixels per CTA unsupported (HW=%d, head_dim=%d)", pt, HW, hd);
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = allow_pad_interactions;
  if (pt == 16) return launch_fused_nt<96, 16>(t_pad, qkv, rq, rk, rv, mask, pad, B, T, HW, heads, out, st);
  switch (hd) {
    case 32: return launch_fused_nt<32, 8>(t_pad, qkv, rq, rk, rv, mask, pad, B, T, HW, heads, out, st);
    case 64: return launch_fused_nt<64, 8>(t_pad, qkv, rq, rk, rv, mask, pad, B, T, HW, heads, out, st);
    case 96: return launch_fused_nt<96, 8>(t_pad, qkv, rq, rk, rv, mask, pad, B, T, HW, heads, out, st);
    case 128: return launch_fused_nt<128, 8>(t_pad, qkv, rq, rk, rv, mask, pad, B, T, HW, heads, out, st);
  }
  set_error("attn_temporal_fused: head_dim=%d not in {32, 64, 96, 128}", hd);
  return -1;
}
