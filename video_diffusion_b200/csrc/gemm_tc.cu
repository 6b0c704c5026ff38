// Implicit-GEMM convolution / linear on the 5th-gen tensor cores (sm_100a):
//   TMA (cp.async.bulk.tensor, 128B swizzle, OOB zero fill = conv padding) -> shared memory
//   -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM) -> tcgen05.ld epilogue (bias / per-image
//   bias / residual / bf16 or fp32 or NCHW store, GroupNorm statistics of the stored tile).
//
// All kernels are persistent and warp-specialised: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (single
// thread), warps 2..9 epilogue; two accumulator stages in TMEM let the epilogue of one tile overlap the MMAs of the
// next.  Four kernels share the helpers and (except the transposed one) the epilogue below:
//   gemm_tc_kernel              one shifted 5-D TMA box per (tap, 64-channel chunk), then the C2/64 chunks of the
//                               optional second operand (fused 1x1 skip projection, unet.py:172-173,198); single CTA
//                               or CTA pair (cta_group::2): linears, stride-2 convs, small layers
//   gemm_tc_halo_kernel         3x3 stride-1: the three vertical taps share one activation slot (tile rows + halo);
//                               interleaved-tile form for the 8x8 level (two images per tile)
//   gemm_tc_halo_t_kernel       the same with operand roles swapped (weights = M, 256 pixels = N), weight tiles
//                               multicast across a 2-CTA cluster, staging-free epilogue: 128- and 384-channel levels
//   gemm_tc_upfold_halo_kernel  nearest-x2 upsample folded into 2x2 parity convs, both vertical parities per tile
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace vdm {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + 32 * EPI_WARPS;
// Kernels with the fused-normalisation stage (XF) add four transform warps (threads NUM_THREADS ..): they rewrite
// every activation slot in shared memory, between the TMA load and the MMA, to act(a * x + b).
constexpr int XF_WARPS = 4;
constexpr int NUM_THREADS_XF = NUM_THREADS + 32 * XF_WARPS;
// Diagnostics build (make TRACE=1): per-role cycle counters (vdm_gemm_set_trace) and the VDM_GEMM_DEBUG timing
// experiments.  Off by default: the single MMA-issuing thread is the critical path and must stay branch-free.
#ifdef VDM_GEMM_TRACE
constexpr bool kTrace = true;
#else
constexpr bool kTrace = false;
#endif

// The second operand range (fused 1x1 skip projection / identity residual) may come from two tensors concatenated
// along channels (the U-Net skip concat): chunks [0, c2a_chunks) from `a`, the rest from `b`.
struct A2Maps {
  CUtensorMap a, b;
};

struct TcParams {
  int M, N;
  int taps, c1_chunks, c2_chunks;
  int c2a_chunks;     // chunks of the second range that come from its first tensor (== c2_chunks for one tensor)
  int a1_f16;         // the first range (activations AND its weight columns) is IEEE half (linears: vdm_gemm dtype VDM_F16)
  int a2_f16;         // the second range (activations AND its weight columns) is IEEE half: those MMAs use the f16 descriptor
  int a1_mode;    // 0 stride-1 / linear, 1 stride-2 parity planes, 3 nearest-x2 upsample folded into 2x2 taps
  int w_group_tiles;  // grouped weights: 128-row tile m reads weight rows (m / w_group_tiles) * N + n
  int tiles_per_par;  // > 0: tile index = par * tiles_per_par + ...; par = output parity (a1_mode 3) or problem of a batch
  int n_par;          // number of parities / problems (1 when tiles_per_par == 0)
  int par_w_rows;     // weight rows per par (added to the weight row coordinate)
  int par_a_cols;     // batch of problems: A1 column offset per problem
  long long par_out_stride;   // batch of problems: output element offset per problem
  int is_linear;  // A1 addressed as 2-D [M][C1]
  int H, W, HW;
  const float* bias;
  const float* rowbias;
  int ld_rowbias;
  const float* residual;
  int ld_res;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ld_out, ld_out_bf16;
  int out_nchw;
  int64_t* stats_out;
  int io_f16;                  // out_f32 / residual point to fp16 tensors (EPI bit 6 on the lean paths)
  const float2* xf_coef;       // XF kernels: per-(image, A1 channel) (a, b) of the fused GroupNorm-apply
  int xf_act;                  // XF kernels: 1 = SiLU after the affine
  int stats_via_smem;          // fold the GroupNorm partial sums of a tile in shared memory
  int dbg;                     // diagnostics (VDM_GEMM_DEBUG): bit 0 skip the TMA loads, bit 1 skip the MMAs (results are garbage)
  unsigned long long* trace;   // diagnostics (vdm_gemm_set_trace): per-CTA wait / busy cycle counters, else NULL
  unsigned int* img_done;      // per-image completion counters (vdm_gemm_args.img_done), transposed-role kernels only
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must fail loudly, never hang the GPU box.  On a timeout the kernel
// reports which barrier stalled, raises a device-wide abort flag (every later wait returns at once,
// so the grid drains and the message is flushed) and the host sees the flag after the launch.
__device__ int g_gemm_abort = 0;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (*reinterpret_cast<volatile int*>(&g_gemm_abort)) return;
    if (clock64() - t0 > 600000000LL) {  // ~0.3 s
      printf("vdm gemm_tc: mbarrier timeout tag=%d parity=%u (block %d thread %d)\n", tag, parity, blockIdx.x,
             threadIdx.x);
      atomicExch(&g_gemm_abort, 1);
      return;
    }
  }
}

__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
      "%7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                       // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset: 8 rows x 128 B
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// ---- 2-CTA (cta_group::2) forms: the CTA pair of a cluster executes one M=256 MMA; the even CTA (leader)
// issues it, operands come from BOTH CTAs' shared memory at identical offsets, TMA loads issued by either
// CTA signal the LEADER's mbarrier (cluster address with the peer bit cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on CTA `rank`'s copy of a barrier (remote arrive through the cluster address space)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}

// The same without release semantics: handing a drained TMEM accumulator stage back to the leader's MMA thread publishes
// no generic-proxy data (the tcgen05.ld's are complete -- tcgen05.wait::ld -- and ordered by
// tcgen05.fence::before_thread_sync), and the release form costs a MEMBAR.ALL.CTA + ERRBAR per warp and tile that waits for
// the lane's staging stores (ncu: 17 % of the samples of the short-K linears sat on it).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}

// ---- cluster multicast forms used by the transposed halo kernel (two independent cta_group::1 CTAs that share
// every weight tile): a TMA load lands at the same shared-memory offset in every CTA of the mask and signals the
// mbarrier at the same offset there; a commit arrives on the barrier of every CTA of the mask.
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// clears the A / B format fields (1 = bf16) of an instruction descriptor: both operands IEEE half (0)
constexpr uint32_t kIdescF16Mask = ~((1u << 7) | (1u << 10));

template <int BLOCK_N, int UMMA_M = BLOCK_M>
constexpr uint32_t instr_desc() {
  return (1u << 4)                          // accumulator fp32
         | (1u << 7) | (1u << 10)           // A, B = bf16
         | ((uint32_t)(BLOCK_N >> 3) << 17) // N
         | ((uint32_t)(UMMA_M >> 4) << 24);  // M (256 = one MMA across the CTA pair)
}

template <int BLOCK_N, int M_SUB>
constexpr int tmem_cols() {   // two accumulator stages, rounded up to the power of two tcgen05.alloc wants
  int need = 2 * M_SUB * BLOCK_N, c = 32;
  while (c < need) c *= 2;
  return c;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ===================== transform role (XF kernels): fused GroupNorm-apply + SiLU on the A operand =====================
// One activation slot, in place.  The slot holds `rows` pixel rows of 128 B (64 bf16 channels), 128B-swizzled exactly as
// TMA wrote them: the 16-byte column j of row r sits at r * 128 + ((j ^ (r & 7)) << 4).  Thread tid of the 128 owns the
// logical column j = tid & 7 (channels 8j .. 8j+7 of the chunk: its eight (a, b) pairs stay in registers) and the rows
// r = (tid >> 3) + 16 i, for which r & 7 -- hence the physical column -- is constant.  Row r is pixel
// (yy, xx) = (r / W, r % W) of the halo box at image position (y_first + yy, xx + x_off); pixels outside the image are
// the convolution's zero padding (TMA out-of-bounds fill) and must stay zero, so they are skipped.
// ILV (8x8 level): box rows are (y, image, x)-ordered, r -> (r >> 4, r & 7); a thread's rows all belong to one image.
template <bool ILV>
__device__ __forceinline__ void xf_transform_slot(uint8_t* slot, int rows, int log2w, int W, int H, int y_first,
                                                  int x_off, const float (&a)[8], const float (&b)[8], int act,
                                                  int tid) {
  const int j = tid & 7, r0 = tid >> 3;
  uint8_t* const base = slot + ((j ^ (r0 & 7)) << 4);
#pragma unroll 4
  for (int r = r0; r < rows; r += 16) {
    const int yy = ILV ? (r >> 4) : (r >> log2w);
    const int xx = ILV ? (r & 7) : (r & (W - 1));
    if ((unsigned)(y_first + yy) < (unsigned)H && (unsigned)(xx + x_off) < (unsigned)W) {
      uint4* const ptr = reinterpret_cast<uint4*>(base + r * 128);
      const uint4 v = *ptr;
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
      uint32_t o[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {       // bf16 -> fp32 is a 16-bit shift
        float y0 = fmaf(__uint_as_float(w[i] << 16), a[2 * i], b[2 * i]);
        float y1 = fmaf(__uint_as_float(w[i] & 0xffff0000u), a[2 * i + 1], b[2 * i + 1]);
        if (act) {
          y0 = silu_tanh(y0);
          y1 = silu_tanh(y1);
        }
        o[i] = pack_bf16x2(y0, y1);
      }
      *ptr = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}
// the thread's eight (a, b) pairs of one 64-channel chunk: coef = table row of the image + chunk * 64 + 8 * (tid & 7)
__device__ __forceinline__ void xf_load_coef(const float2* coef, float (&a)[8], float (&b)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(coef) + i);
    a[2 * i] = t.x; b[2 * i] = t.y; a[2 * i + 1] = t.z; b[2 * i + 1] = t.w;
  }
}
// generic-proxy writes -> visible to the tensor core's async-proxy reads, then one arrive per warp
__device__ __forceinline__ void xf_publish_fence() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
}

// TS_ESIZE > 0: the TMA-store epilogue (no statistics, no residual): the staging area holds the CTA's whole 128 x BLOCK_N
// output tile (elements of TS_ESIZE bytes) as 128-byte-swizzled slabs of 128 rows x 128 B.
template <int BLOCK_N, int M_SUB, int STAGES, bool CTA2 = false, int TS_ESIZE = 0>
struct SmemLayout {
  static constexpr int CHUNK = (BLOCK_N < 32 || M_SUB > 1) ? 16 : 32;
  static constexpr int A_SUB_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int A_BYTES = M_SUB * A_SUB_BYTES;
  static constexpr int B_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * BLOCK_K * 2;   // 2-CTA: each CTA holds half of N
  static constexpr int B_STRIDE = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int STAGE_BYTES = A_BYTES + B_STRIDE;
  static constexpr int STG_OFFSET = STAGES * STAGE_BYTES;                    // epilogue staging
  static constexpr int TOTAL_CHUNKS = M_SUB * (BLOCK_N / CHUNK);
  static constexpr int WARP_STG_FLOATS = 32 * (CHUNK + 4);
  static constexpr int STG_BYTES = TS_ESIZE ? BLOCK_M * BLOCK_N * TS_ESIZE : EPI_WARPS * WARP_STG_FLOATS * 4;
  static constexpr int STAT_IMGS = 2;                                         // images a CTA tile may span (H*W >= 64)
  static constexpr int STAT_OFFSET = STG_OFFSET + STG_BYTES;                  // GroupNorm partial sums per lane quarter
  static constexpr int STAT_BYTES = TS_ESIZE ? 0 : 4 * STAT_IMGS * 2 * BLOCK_N * 4;
  static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 4;
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;         // + alignment slack
};

// ===================== epilogue role (warps 2..9), shared by the kernels below =====================
// tmem_full_bar0 / tmem_empty_bar0: shared-memory addresses of the two-entry barrier arrays of the accumulator stages.
// stab: zeroed shared-memory table [4 lane quarters][stat_imgs][2][BLOCK_N] of fp32 GroupNorm partial sums.
// ILV (8x8 level of the halo kernel): the CTA's 128 tile rows are two whole images interleaved by image row,
// tile row l = (y, image, x) with x fastest, i.e. global row = first + (l>>3 & 1)*64 + (l>>4)*8 + (l&7).
// UPF (folded-upsample halo kernel): the M_SUB = 2 accumulators of a CTA are the two vertical output parities
// (a = 0, 1) of the SAME 128 low-resolution pixels; the tile's `par` is the horizontal parity b.
template <int BLOCK_N, int M_SUB, int CHUNK, int EPI, bool CTA2, bool ILV = false, bool UPF = false>
__device__ __forceinline__ void epilogue_role(const TcParams& p, float* stg_base, float* stab,
                                              int stat_imgs, uint32_t tmem_base,
                                              uint32_t tmem_full_bar0, uint32_t tmem_empty_bar0, int n_tiles,
                                              int n_tiles_n, int work_id0, int work_step, uint32_t cta_rank, int warp,
                                              int lane) {
  constexpr int CTA_M = BLOCK_M * (UPF ? 1 : M_SUB);          // distinct A rows per CTA
  constexpr int TILE_M = CTA_M * (CTA2 ? 2 : 1);
  static_assert(!UPF || (M_SUB == 2 && (EPI & 1) == 0), "folded upsample: two parity accumulators, no residual");
  constexpr bool HAS_RES = (EPI & 1) != 0, BF16_OUT = (EPI & 2) != 0, GEN = (EPI & 8) != 0;
  constexpr bool STATS = GEN || (EPI & 4) != 0;
  // fp16 stream: `out_f32` / `residual` are half tensors.  Compile-time on the lean paths (EPI bit 6), a uniform
  // run-time flag on the generic one.
  constexpr bool F16IO = (EPI & 64) != 0;
  const bool f16io = GEN ? (p.io_f16 != 0) : F16IO;
  const int io_shift = f16io ? 1 : 2;                 // log2 of the stream element size
  static_assert(!ILV || (!GEN && M_SUB == 1 && CHUNK == 32), "interleaved 8x8 tiles: lean epilogue, 32-column chunks");
  auto ilv_row = [](int l) { return ((l >> 3) & 1) * 64 + (l >> 4) * 8 + (l & 7); };   // tile row -> row offset
  auto tmem_full_bar = [&](int a) { return tmem_full_bar0 + 8u * a; };
  auto tmem_empty_bar = [&](int a) { return tmem_empty_bar0 + 8u * a; };
  // TMEM hands each lane one accumulator ROW; writing rows straight to global memory would touch
  // 32 cache lines per instruction.  Each warp stages its 32 x CHUNK block in shared memory (row
  // stride CHUNK+4 floats keeps 128-bit accesses conflict-free both ways) and re-reads it so that
  // CHUNK/4 lanes cover one contiguous row segment: fully coalesced residual loads and stores.
  constexpr int N_CHUNKS = BLOCK_N / CHUNK;
  constexpr int STG_LD = CHUNK + 4;
  constexpr int LPR = CHUNK / 4;   // lanes per row
  constexpr int RPI = 32 / LPR;    // rows per instruction
  constexpr int NRES = 32 / RPI;   // float4 per lane per chunk
  const int ew = warp - 2;
  const int q = warp & 3;          // TMEM lane quarter this warp may access
  const int half = ew >> 2;        // the two warps of a quarter split the column chunks
  float* stg = stg_base + ew * (32 * (CHUNK + 4));
  const int r_sub = lane / LPR, c4 = (lane % LPR) * 4;
  int it = 0;
  long long tr_wait = 0;
  const long long tr_start = (kTrace && p.trace) ? clock64() : 0;
  for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
    int par = 0, tl = tile;
    if (p.tiles_per_par > 0) {
      par = tile / p.tiles_per_par;
      tl = tile - par * p.tiles_per_par;
    }
    const int n0 = (tl % n_tiles_n) * BLOCK_N;
    const int mt0 = (tl / n_tiles_n) * TILE_M + (int)cta_rank * CTA_M;
    // GroupNorm statistics go through a shared-memory table (one global atomic per (image, channel) and tile
    // instead of one per 32 rows) whenever this CTA's rows span at most stat_imgs images.  Every (quarter, column)
    // entry has exactly one writer warp, so plain read-modify-writes in a fixed order: deterministic.
    const int img_first = mt0 / p.HW;
    const bool use_tab = STATS && p.stats_out != nullptr && p.stats_via_smem &&
                         (mt0 + CTA_M - 1) / p.HW - img_first < stat_imgs;
    float* const tabq = stab + (size_t)q * stat_imgs * 2 * BLOCK_N;
    // UPF, lean path: high-resolution output row of each of this lane's row groups for vertical parity 0 (parity 1
    // is one output image row = 2W rows further); the same for every chunk of the tile
    int up_row[UPF ? NRES : 1];
    if constexpr (UPF && !GEN) {
#pragma unroll
      for (int i = 0; i < NRES; ++i) {
        const int orow = mt0 + q * 32 + i * RPI + r_sub;
        const int img = orow / p.HW, rem = orow - img * p.HW;
        const int yy = rem / p.W, xx = rem - yy * p.W;
        up_row[i] = (img * 2 * p.H + 2 * yy) * (2 * p.W) + 2 * xx + par;
      }
    }
    const int as = it & 1;
    const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
    float4 res_cur[NRES];
    // chunk index jj enumerates (sub-tile, column chunk): jj = sub * N_CHUNKS + j
    auto load_residual = [&](float4 (&dst)[NRES], int jj) {
      if constexpr (HAS_RES) {
        const int m0 = mt0 + (jj / N_CHUNKS) * BLOCK_M;
        const int n = n0 + (jj % N_CHUNKS) * CHUNK + c4;
#pragma unroll
        for (int i = 0; i < NRES; ++i) {
          const int orow = ILV ? m0 + ilv_row(q * 32 + i * RPI + r_sub) : m0 + q * 32 + i * RPI + r_sub;
          dst[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (orow < p.M && n < p.N) {
            if (f16io) {
              // RAW bits only: a conversion right behind each load would make the in-order warp wait for that load
              // before issuing the next one (measured: the fp16 residual read twice as slow as the fp32 one)
              const uint2 h = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.residual) +
                                                                   (size_t)orow * p.ld_res + n));
              dst[i] = make_float4(__uint_as_float(h.x), __uint_as_float(h.y), 0.f, 0.f);
            } else {
              dst[i] = __ldg(reinterpret_cast<const float4*>(p.residual + (size_t)orow * p.ld_res + n));
            }
          }
        }
      }
    };
    auto res_value = [&](const float4& raw) {     // what load_residual fetched, as fp32
      if (f16io) {
        const float2 lo = unpack_f16x2(__float_as_uint(raw.x)), hi = unpack_f16x2(__float_as_uint(raw.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      return raw;
    };
    // The residual tile (up to 128 KB) is far more than the one-chunk-ahead register prefetch below keeps in
    // flight, which made the residual read latency-bound (~1.6 TB/s).  Pull the whole tile into L2 now, while
    // this tile's MMAs are still running; the LDGs below then hit L2.
    if constexpr (HAS_RES) {
      const int lines_per_row = (BLOCK_N << io_shift) / 128, cols_per_line = 128 >> io_shift;   // 128-byte lines
      for (int i = ew * 32 + lane; i < M_SUB * BLOCK_M * lines_per_row; i += EPI_WARPS * 32) {
        const int row = mt0 + i / lines_per_row, n = n0 + (i % lines_per_row) * cols_per_line;
        if (row < p.M && n < p.N)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.residual) +
                                                        (((size_t)row * p.ld_res + n) << io_shift)));
      }
    }
    // fetched one chunk ahead -- the first one while this tile's MMAs are still running
    constexpr int TOTAL_CHUNKS = M_SUB * N_CHUNKS;
    if (half < TOTAL_CHUNKS) load_residual(res_cur, half);
    // the bias quad of the next chunk is fetched one chunk ahead as well: the streaming stores keep evicting it
    // from L1, and waiting ~500 cycles for it at its first use was a quarter of the per-chunk time
    auto load_bias = [&](int jj) {
      const int n = n0 + (jj % N_CHUNKS) * CHUNK + c4;
      return (p.bias && n < p.N) ? __ldg(reinterpret_cast<const float4*>(p.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    float4 bias_cur = make_float4(0.f, 0.f, 0.f, 0.f);
    if (half < TOTAL_CHUNKS) bias_cur = load_bias(half);
    if (kTrace && p.trace) {
      const long long t0 = clock64();
      mbar_wait(tmem_full_bar(as), aphase, 2);
      tr_wait += clock64() - t0;
    } else {
      mbar_wait(tmem_full_bar(as), aphase, 2);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int jj = half; jj < TOTAL_CHUNKS; jj += 2) {
      const int sub = jj / N_CHUNKS, j = jj - sub * N_CHUNKS;
      const int m0 = mt0 + (UPF ? 0 : sub * BLOCK_M);
      const int opar = UPF ? sub * 2 + par : par;   // output parity (a, b) of this accumulator
      const int row = m0 + q * 32 + lane;
      uint32_t acc[CHUNK];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) +
                             (uint32_t)(as * M_SUB * BLOCK_N + sub * BLOCK_N + j * CHUNK);
      if constexpr (CHUNK == 32) tmem_ld_32(taddr, acc);
      else tmem_ld_16(taddr, acc);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int nb = n0 + j * CHUNK;
      if (GEN && p.out_nchw) {  // lanes = consecutive pixels: already coalesced per channel
        if (row < p.M) {
          const int img = row / p.HW, pix = row - img * p.HW;
          for (int i = 0; i < CHUNK; ++i) {
            const int n = nb + i;
            if (n < p.N) {
              float v = __uint_as_float(acc[i]);
              if (p.bias) v += p.bias[n];
              p.out_f32[((size_t)img * p.N + n) * p.HW + pix] = v;
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int i = 0; i < CHUNK; i += 4)
        *reinterpret_cast<float4*>(stg + lane * STG_LD + i) =
            make_float4(__uint_as_float(acc[i]), __uint_as_float(acc[i + 1]), __uint_as_float(acc[i + 2]),
                        __uint_as_float(acc[i + 3]));
      __syncwarp();
      float4 res_next[NRES];
      float4 bias_next = make_float4(0.f, 0.f, 0.f, 0.f);
      if (jj + 2 < TOTAL_CHUNKS) {
        load_residual(res_next, jj + 2);
        bias_next = load_bias(jj + 2);
      }
      const int n = nb + c4;
      float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = make_float4(0.f, 0.f, 0.f, 0.f);
      float4 ssum1 = make_float4(0.f, 0.f, 0.f, 0.f), ssq1 = make_float4(0.f, 0.f, 0.f, 0.f);   // ILV: second image
      if constexpr (!GEN) {
        // lean path: N is a multiple of BLOCK_N, exactly one output, no per-image bias / row remap
        const float4 bv = bias_cur;
        const int row_first = m0 + q * 32 + r_sub;
        const bool full = m0 + BLOCK_M <= p.M;
        const float* sp = stg + r_sub * STG_LD + c4;
        // all staged rows first (independent LDS in flight), then the arithmetic and the stores: left to the
        // compiler the loop came out as a serial load -> convert -> store chain of ~130 cycles per row group
        float4 vv[NRES];
#pragma unroll
        for (int i = 0; i < NRES; ++i) vv[i] = *reinterpret_cast<const float4*>(sp + i * RPI * STG_LD);
        __syncwarp();   // keeps ptxas from sinking the loads back next to their uses
        // the epilogue is instruction-issue bound: one running output pointer (no per-row 64-bit multiply) and,
        // for full tiles, no per-row bounds test
        auto finish_row = [&](int i, auto* optr) {
          float4 v = vv[i];
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
          if constexpr (HAS_RES) {
            const float4 r = res_value(res_cur[i]);
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
          }
          if constexpr (STATS) {
            if (ILV && (((i * RPI) >> 3) & 1)) {   // rows 8..15, 24..31 of the warp belong to the second image
              ssum1.x += v.x; ssum1.y += v.y; ssum1.z += v.z; ssum1.w += v.w;
              ssq1.x = fmaf(v.x, v.x, ssq1.x); ssq1.y = fmaf(v.y, v.y, ssq1.y);
              ssq1.z = fmaf(v.z, v.z, ssq1.z); ssq1.w = fmaf(v.w, v.w, ssq1.w);
            } else {
              ssum.x += v.x; ssum.y += v.y; ssum.z += v.z; ssum.w += v.w;
              ssq.x = fmaf(v.x, v.x, ssq.x); ssq.y = fmaf(v.y, v.y, ssq.y);
              ssq.z = fmaf(v.z, v.z, ssq.z); ssq.w = fmaf(v.w, v.w, ssq.w);
            }
          }
          if constexpr (BF16_OUT) {
            uint2 pk;
            pk.x = pack_bf16x2(v.x, v.y);
            pk.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(optr) = pk;
          } else if constexpr (F16IO) {
            uint2 pk;
            pk.x = pack_f16x2(v.x, v.y);
            pk.y = pack_f16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(optr) = pk;
          } else {
            *reinterpret_cast<float4*>(optr) = v;
          }
        };
        auto store_rows = [&](auto* optr, const size_t step) {
          if constexpr (UPF) {   // optr = the output's first row; rows are scattered over the high-resolution image
            const size_t ld = step / RPI;
#pragma unroll
            for (int i = 0; i < NRES; ++i)
              if (full || row_first + i * RPI < p.M) finish_row(i, optr + (size_t)(up_row[i] + sub * 2 * p.W) * ld);
          } else if constexpr (ILV) {   // optr = the tile's first row; every row group has its own offset
            const size_t ld = step / RPI;
#pragma unroll
            for (int i = 0; i < NRES; ++i) {
              const int off = ilv_row(q * 32 + i * RPI + r_sub);
              if (full || m0 + off < p.M) finish_row(i, optr + (size_t)off * ld);
            }
          } else if (full) {
#pragma unroll
            for (int i = 0; i < NRES; ++i, optr += step) finish_row(i, optr);
          } else {
#pragma unroll
            for (int i = 0; i < NRES; ++i, optr += step)
              if (row_first + i * RPI < p.M) finish_row(i, optr);
          }
        };
        const int row_base = UPF ? 0 : (ILV ? m0 : row_first);
        if constexpr (BF16_OUT)
          store_rows(p.out_bf16 + par * p.par_out_stride + (size_t)row_base * p.ld_out_bf16 + n, (size_t)RPI * p.ld_out_bf16);
        else if constexpr (F16IO)
          store_rows(reinterpret_cast<__half*>(p.out_f32) + par * p.par_out_stride + (size_t)row_base * p.ld_out + n,
                     (size_t)RPI * p.ld_out);
        else
          store_rows(p.out_f32 + par * p.par_out_stride + (size_t)row_base * p.ld_out + n, (size_t)RPI * p.ld_out);
      } else if (n < p.N) {  // N is a multiple of 4 on this path
        const float4 bv = bias_cur;
#pragma unroll
        for (int rr = 0; rr < 32; rr += RPI) {
          const int rl = rr + r_sub;
          const int orow = m0 + q * 32 + rl;
          if (orow < p.M) {
            float4 v = *reinterpret_cast<const float4*>(stg + rl * STG_LD + c4);
            v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
            if (p.rowbias) {
              const float4 b = *reinterpret_cast<const float4*>(p.rowbias + (size_t)(orow / p.HW) * p.ld_rowbias + n);
              v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
            }
            if constexpr (HAS_RES) {
              const float4 r = res_value(res_cur[rr / RPI]);
              v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
            }
            ssum.x += v.x; ssum.y += v.y; ssum.z += v.z; ssum.w += v.w;
            ssq.x = fmaf(v.x, v.x, ssq.x); ssq.y = fmaf(v.y, v.y, ssq.y);
            ssq.z = fmaf(v.z, v.z, ssq.z); ssq.w = fmaf(v.w, v.w, ssq.w);
            size_t drow = (size_t)orow;
            if (p.a1_mode == 3) {   // low-res pixel (img, y, x) of parity (a, b) -> high-res row
              const int img = orow / p.HW, rem = orow - img * p.HW;
              const int yy = rem / p.W, xx = rem - yy * p.W;
              drow = ((size_t)img * 2 * p.H + 2 * yy + (opar >> 1)) * (2 * p.W) + 2 * xx + (opar & 1);
            }
            if (p.out_f32) {
              if (f16io) {
                uint2 pk;
                pk.x = pack_f16x2(v.x, v.y);
                pk.y = pack_f16x2(v.z, v.w);
                *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out_f32) + drow * p.ld_out + n) = pk;
              } else {
                *reinterpret_cast<float4*>(p.out_f32 + drow * p.ld_out + n) = v;
              }
            }
            if (p.out_bf16) {
              uint2 pk;
              pk.x = pack_bf16x2(v.x, v.y);
              pk.y = pack_bf16x2(v.z, v.w);
              *reinterpret_cast<uint2*>(p.out_bf16 + drow * p.ld_out_bf16 + n) = pk;
            }
          }
        }
      }
      if (STATS && p.stats_out != nullptr) {
        const int row0 = m0 + q * 32;      // the warp's 32 rows lie in one image (H*W % 32 == 0)
        if (use_tab) {
          // transpose the lanes' partial sums through the (now idle) staging block so that a lane owns a column
          __syncwarp();
          float* part = stg;               // [RPI row groups][sum, sum of squares][CHUNK]
          *reinterpret_cast<float4*>(part + (r_sub * 2 + 0) * CHUNK + c4) = ssum;
          *reinterpret_cast<float4*>(part + (r_sub * 2 + 1) * CHUNK + c4) = ssq;
          __syncwarp();
          float* trow = tabq + (size_t)(ILV ? 0 : row0 / p.HW - img_first) * 2 * BLOCK_N + (nb - n0);
#pragma unroll
          for (int idx = lane; idx < 2 * CHUNK; idx += 32) {
            const int k = idx / CHUNK, col = idx - k * CHUNK;
            float sacc = 0.f;
#pragma unroll
            for (int r = 0; r < RPI; ++r) sacc += part[(r * 2 + k) * CHUNK + col];
            if (ILV || row0 < p.M) trow[k * BLOCK_N + col] += sacc;   // ILV: every warp holds rows of the first image
          }
          if constexpr (ILV) {   // the second image's partial sums, table row 1
            __syncwarp();
            *reinterpret_cast<float4*>(part + (r_sub * 2 + 0) * CHUNK + c4) = ssum1;
            *reinterpret_cast<float4*>(part + (r_sub * 2 + 1) * CHUNK + c4) = ssq1;
            __syncwarp();
#pragma unroll
            for (int idx = lane; idx < 2 * CHUNK; idx += 32) {
              const int k = idx / CHUNK, col = idx - k * CHUNK;
              float sacc = 0.f;
#pragma unroll
              for (int r = 0; r < RPI; ++r) sacc += part[(r * 2 + k) * CHUNK + col];
              if (mt0 + 64 < p.M) trow[(2 + k) * BLOCK_N + col] += sacc;
            }
          }
        } else {
          // direct route: fold the lanes that share a column quad, then add this warp's 32-row partial sums to the
          // per-(image, channel) table with 64-bit fixed-point atomics (integer addition is associative, so the
          // statistics and everything downstream are bit-reproducible from run to run)
#pragma unroll
          for (int o = LPR; o < 32; o <<= 1) {
            ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, o); ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, o);
            ssum.z += __shfl_xor_sync(0xffffffffu, ssum.z, o); ssum.w += __shfl_xor_sync(0xffffffffu, ssum.w, o);
            ssq.x += __shfl_xor_sync(0xffffffffu, ssq.x, o); ssq.y += __shfl_xor_sync(0xffffffffu, ssq.y, o);
            ssq.z += __shfl_xor_sync(0xffffffffu, ssq.z, o); ssq.w += __shfl_xor_sync(0xffffffffu, ssq.w, o);
          }
          if (r_sub == 0 && n < p.N && row0 < p.M) {
            const float sv[8] = {ssum.x, ssum.y, ssum.z, ssum.w, ssq.x, ssq.y, ssq.z, ssq.w};
            unsigned long long* tab = reinterpret_cast<unsigned long long*>(p.stats_out) +
                                      (size_t)(row0 / p.HW) * 2 * p.N + n;
#pragma unroll
            for (int e = 0; e < 8; ++e)   // 2^24 scaling of a float is exact
              atomicAdd(tab + (e >> 2) * p.N + (e & 3), (unsigned long long)__float2ll_rn(sv[e] * 16777216.0f));
          }
        }
      }
      if constexpr (HAS_RES) {
        if (jj + 2 < TOTAL_CHUNKS) {
#pragma unroll
          for (int i = 0; i < NRES; ++i) res_cur[i] = res_next[i];
        }
      }
      bias_cur = bias_next;
      __syncwarp();
    }
    // all of this warp's TMEM reads of the stage are complete: hand it back to the MMA warp
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if constexpr (CTA2) mbar_arrive_cluster_relaxed(tmem_empty_bar(as), 0);   // the leader's MMA thread owns the wait
      else mbar_arrive(tmem_empty_bar(as));
    }
    if (use_tab) {
      // all eight epilogue warps have added their partial sums: one thread per (image, statistic, channel) folds
      // the four quarters in fixed point (order-free) and issues the only global atomic of this tile for it
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      const int per_q = stat_imgs * 2 * BLOCK_N;
      for (int i = ew * 32 + lane; i < per_q; i += EPI_WARPS * 32) {
        long long fx = 0;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          fx += __float2ll_rn(stab[qq * per_q + i] * 16777216.0f);
          stab[qq * per_q + i] = 0.f;
        }
        if (fx != 0) {
          const int img = img_first + i / (2 * BLOCK_N), k = (i / BLOCK_N) & 1, col = n0 + i % BLOCK_N;
          atomicAdd(reinterpret_cast<unsigned long long*>(p.stats_out) + ((size_t)img * 2 + k) * p.N + col,
                    (unsigned long long)fx);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");   // table is clean before the next tile adds
    }
  }
  if (kTrace && p.trace && ew == 0 && lane == 0) {
    p.trace[blockIdx.x * 8 + 4] = (unsigned long long)tr_wait;
    p.trace[blockIdx.x * 8 + 5] = (unsigned long long)(clock64() - tr_start);
  }
}

// ===================== epilogue role with TMA stores (plain linears: bias only, one output) =====================
// The short-K linears of the attention blocks (qkv, the RPE score GEMMs) are bound by the epilogue above: per 32-column
// chunk it stages, re-reads and issues eight row-quad stores per lane.  Here a lane keeps its accumulator ROW: it adds
// the bias, converts and writes the row segment into a shared-memory image of the output tile (128-byte-swizzled slabs
// of 128 rows x 128 B, the layout a SWIZZLE_128B tensor map expects; conflict-free 16-byte stores), and one elected
// thread hands the finished tile to the TMA engine (cp.async.bulk.tensor store, fully coalesced, asynchronous: it
// drains while the next tile's accumulator is converted).  Rows / columns beyond M / N are clipped by the tensor map.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
template <int BLOCK_N, bool BF16, bool CTA2>
__device__ __forceinline__ void epilogue_ts_role(const TcParams& p, const CUtensorMap* tm_out, uint8_t* stg,
                                                 uint32_t stg_addr, uint32_t tmem_base, uint32_t tmem_full_bar0,
                                                 uint32_t tmem_empty_bar0, int n_tiles, int n_tiles_n, int work_id0,
                                                 int work_step, uint32_t cta_rank, int warp, int lane) {
  constexpr int TILE_M = BLOCK_M * (CTA2 ? 2 : 1);
  constexpr int ESZ = BF16 ? 2 : 4;
  constexpr int SLAB_COLS = 128 / ESZ;                 // output columns per 128-byte slab row
  constexpr int N_SLABS = BLOCK_N / SLAB_COLS;
  constexpr int SLAB_BYTES = BLOCK_M * 128;
  constexpr int N_CHUNKS = BLOCK_N / 32;               // accumulator columns are read 32 at a time
  const int ew = warp - 2, q = warp & 3, half = ew >> 2;
  const int row = q * 32 + lane;                       // tile row of this lane = its TMEM lane
  uint8_t* const rowp = stg + row * 128;
  const int sw = row & 7;
  int it = 0;
  for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
    int par = 0, tl = tile;
    if (p.tiles_per_par > 0) {
      par = tile / p.tiles_per_par;
      tl = tile - par * p.tiles_per_par;
    }
    const int n0 = (tl % n_tiles_n) * BLOCK_N;
    const int m0 = (tl / n_tiles_n) * TILE_M + (int)cta_rank * BLOCK_M;
    const int as = it & 1;
    const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
    mbar_wait(tmem_full_bar0 + 8u * as, aphase, 2);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int jj = half; jj < N_CHUNKS; jj += 2) {
      uint32_t acc[32];
      tmem_ld_32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * BLOCK_N + jj * 32), acc);
      float bv[32];
      if (p.bias != nullptr) {      // the chunk's 32 bias values: uniform addresses, broadcast loads
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + jj * 32 + i));
          bv[i] = t.x; bv[i + 1] = t.y; bv[i + 2] = t.z; bv[i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) bv[i] = 0.f;
      }
      if (jj == half && it > 0) {
        // the previous tile's stores must have READ the staging area before it is overwritten -- waited for here, under
        // the TMEM and bias loads of this warp's first chunk, not in front of them
        if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if constexpr (BF16) {         // 32 columns = 64 B = four 16-byte pieces of slab jj / 2
        uint8_t* const sp = rowp + (jj >> 1) * SLAB_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint4 v;
          v.x = pack_bf16x2(__uint_as_float(acc[8 * k + 0]) + bv[8 * k + 0], __uint_as_float(acc[8 * k + 1]) + bv[8 * k + 1]);
          v.y = pack_bf16x2(__uint_as_float(acc[8 * k + 2]) + bv[8 * k + 2], __uint_as_float(acc[8 * k + 3]) + bv[8 * k + 3]);
          v.z = pack_bf16x2(__uint_as_float(acc[8 * k + 4]) + bv[8 * k + 4], __uint_as_float(acc[8 * k + 5]) + bv[8 * k + 5]);
          v.w = pack_bf16x2(__uint_as_float(acc[8 * k + 6]) + bv[8 * k + 6], __uint_as_float(acc[8 * k + 7]) + bv[8 * k + 7]);
          const int j = (jj & 1) * 4 + k;
          *reinterpret_cast<uint4*>(sp + ((j ^ sw) << 4)) = v;
        }
      } else {                      // 32 columns = 128 B = the whole row of slab jj
        uint8_t* const sp = rowp + jj * SLAB_BYTES;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float4 v = make_float4(__uint_as_float(acc[4 * k + 0]) + bv[4 * k + 0], __uint_as_float(acc[4 * k + 1]) + bv[4 * k + 1],
                                       __uint_as_float(acc[4 * k + 2]) + bv[4 * k + 2], __uint_as_float(acc[4 * k + 3]) + bv[4 * k + 3]);
          *reinterpret_cast<float4*>(sp + ((k ^ sw) << 4)) = v;
        }
      }
    }
    // accumulator stage drained: hand it back to the MMA warp; then publish the tile image to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
      if constexpr (CTA2) mbar_arrive_cluster_relaxed(tmem_empty_bar0 + 8u * as, 0);
      else mbar_arrive(tmem_empty_bar0 + 8u * as);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    if (ew == 0 && lane == 0) {
#pragma unroll
      for (int sl = 0; sl < N_SLABS; ++sl) tma_store_3d(tm_out, stg_addr + sl * SLAB_BYTES, n0 + sl * SLAB_COLS, m0, par);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
  if (ew == 0 && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
}

// Persistent, warp-specialised kernel: grid = min(#tiles, #SMs); every role loops over the CTA's
// tiles (tile = blockIdx.x + i*gridDim.x, N-tile fastest so CTAs running together share A in L2).
//   warp 0      : TMA producer (one lane), STAGES-deep smem ring
//   warp 1      : TMEM allocator + tcgen05.mma issuer (one lane), two accumulator stages in TMEM
//   warps 2..9  : epilogue; the accumulator of tile i drains while the MMAs of tile i+1 run
// EPI selects the epilogue at compile time (the epilogue is instruction-issue bound, so every option that
// is a runtime branch inside its row loop costs throughput): bit 0 residual add, bit 1 bf16 output (else
// fp32), bit 2 GroupNorm statistics; bit 3 = generic path with every option decided at run time (per-image
// bias, both outputs, NCHW store, folded-upsample row remap, ragged N).
// bit 5 (with bit 1 = bf16): the TMA-store epilogue above (bias only, no residual / statistics), M_SUB == 1.
template <int BLOCK_N, int M_SUB, int STAGES, int EPI, bool CTA2 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(const __grid_constant__ CUtensorMap tm_a1,
                                                                 const __grid_constant__ A2Maps tm_a2,
                                                                 const __grid_constant__ CUtensorMap tm_w,
                                                                 const __grid_constant__ CUtensorMap tm_out,
                                                                 const TcParams p) {
  pdl_launch_dependents_persistent();   // the next kernel's prologue may overlap this kernel's tail (common.cuh)
  constexpr bool TS = (EPI & 32) != 0;
  static_assert(!TS || M_SUB == 1, "TMA-store epilogue: one 128-row sub-tile per CTA");
  using L = SmemLayout<BLOCK_N, M_SUB, STAGES, CTA2, TS ? ((EPI & 2) ? 2 : 4) : 0>;
  constexpr int TILE_M = BLOCK_M * M_SUB * (CTA2 ? 2 : 1);     // rows per work tile (CTA or CTA pair)
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;     // 0 = leader
  const int work_id0 = CTA2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int work_step = CTA2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  constexpr bool HAS_RES = (EPI & 1) != 0, BF16_OUT = (EPI & 2) != 0, GEN = (EPI & 8) != 0;
  constexpr bool STATS = GEN || (EPI & 4) != 0;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.taps * p.c1_chunks + p.c2_chunks;
  const int n_tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int n_tiles = p.n_par * n_tiles_n * ((p.M + TILE_M - 1) / TILE_M);

  for (int i = threadIdx.x; i < L::STAT_BYTES / 4; i += NUM_THREADS)
    reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS * (CTA2 ? 2 : 1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if constexpr (CTA2) {   // the same warp of BOTH CTAs allocates the pair's columns
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                   "n"(tmem_cols<BLOCK_N, M_SUB>())
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                   "n"(tmem_cols<BLOCK_N, M_SUB>())
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CTA2) cluster_sync_all();   // peer barriers must be initialised before any remote arrive
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // barriers, TMEM and shared-memory tables are set up: from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      const int k1 = p.taps * p.c1_chunks;
      int stage = 0;
      uint32_t phase = 0;
      long long tr_wait = 0;
      for (int tile = work_id0; tile < n_tiles; tile += work_step) {
        int par = 0, tl = tile;
        if (p.tiles_per_par > 0) {
          par = tile / p.tiles_per_par;
          tl = tile - par * p.tiles_per_par;
        }
        const int n0 = (tl % n_tiles_n) * BLOCK_N;
        const int m0 = (tl / n_tiles_n) * TILE_M + (int)cta_rank * (BLOCK_M * M_SUB);   // a CTA's rows are contiguous
        int img0[M_SUB], y0[M_SUB], x0[M_SUB];
#pragma unroll
        for (int sub = 0; sub < M_SUB; ++sub) {
          img0[sub] = y0[sub] = x0[sub] = 0;
          if (!p.is_linear) {
            const int ms = m0 + sub * BLOCK_M;
            img0[sub] = ms / p.HW;
            const int rem = ms - img0[sub] * p.HW;
            y0[sub] = rem / p.W;
            x0[sub] = rem - y0[sub] * p.W;
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          if (kTrace && p.trace) {
            const long long t0 = clock64();
            mbar_wait(empty_bar(stage), phase ^ 1u, 0);
            tr_wait += clock64() - t0;
          } else {
            mbar_wait(empty_bar(stage), phase ^ 1u, 0);
          }
          const uint32_t a_dst = smem_base + stage * L::STAGE_BYTES;
          const uint32_t b_dst = a_dst + L::A_BYTES;
          if (kTrace && (p.dbg & 1)) {        // timing experiment: operands are whatever the slot holds
            if (cta_rank == 0) mbar_arrive(full_bar(stage));
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
            continue;
          }
          if constexpr (CTA2) {   // the leader's barrier counts the bytes landing in both CTAs
            if (cta_rank == 0) mbar_expect_tx(full_bar(stage), 2 * (L::A_BYTES + L::B_BYTES));
          } else {
            mbar_expect_tx(full_bar(stage), L::A_BYTES + L::B_BYTES);
          }
          int dy = 0, dx = 0, plane = 0, c0 = 0;
          const CUtensorMap* a2_map = &tm_a2.a;
          const bool first_range = kb < k1;
          if (first_range) {
            const int tap = kb / p.c1_chunks;
            c0 = (kb - tap * p.c1_chunks) * BLOCK_K;
            if (p.a1_mode == 3) {   // output parity (a, b): rows {y-1, y} / {y, y+1}, same for columns
              const int i = tap >> 1, j = tap & 1;
              dy = (par >> 1) ? i : i - 1;
              dx = (par & 1) ? j : j - 1;
            } else if (p.taps == 9) {
              const int r = tap / 3, s = tap - r * 3;
              if (p.a1_mode == 0) {
                dy = r - 1;
                dx = s - 1;
              } else {  // stride 2 on parity planes: r -> (parity, offset) = (1,-1), (0,0), (1,0)
                const int py = (r != 1), px = (s != 1);
                dy = (r == 0) ? -1 : 0;
                dx = (s == 0) ? -1 : 0;
                plane = py * 2 + px;
              }
            }
          } else {
            c0 = (kb - k1) * BLOCK_K;
            if (kb - k1 >= p.c2a_chunks) {     // second tensor of a concatenated second range
              c0 -= p.c2a_chunks * BLOCK_K;
              a2_map = &tm_a2.b;
            }
          }
#pragma unroll
          for (int sub = 0; sub < M_SUB; ++sub) {
            const uint32_t dst = a_dst + sub * L::A_SUB_BYTES;
            if constexpr (CTA2) {
              if (!first_range) tma_load_5d_2cta(dst, a2_map, full_bar(stage), c0, m0 + sub * BLOCK_M, 0, 0, 0);
              else if (p.is_linear) tma_load_5d_2cta(dst, &tm_a1, full_bar(stage), c0 + par * p.par_a_cols, m0 + sub * BLOCK_M, 0, 0, 0);
              else tma_load_5d_2cta(dst, &tm_a1, full_bar(stage), c0, x0[sub] + dx, y0[sub] + dy, plane, img0[sub]);
            } else {
              if (!first_range) tma_load_5d(dst, a2_map, full_bar(stage), c0, m0 + sub * BLOCK_M, 0, 0, 0);
              else if (p.is_linear) tma_load_5d(dst, &tm_a1, full_bar(stage), c0 + par * p.par_a_cols, m0 + sub * BLOCK_M, 0, 0, 0);
              else tma_load_5d(dst, &tm_a1, full_bar(stage), c0, x0[sub] + dx, y0[sub] + dy, plane, img0[sub]);
            }
          }
          if constexpr (CTA2)   // this CTA's half of the weight tile
            tma_load_2d_2cta(b_dst, &tm_w, full_bar(stage), kb * BLOCK_K, n0 + par * p.par_w_rows + (int)cta_rank * (BLOCK_N / 2));
          else
            tma_load_2d(b_dst, &tm_w, full_bar(stage), kb * BLOCK_K,
                        n0 + par * p.par_w_rows + (p.w_group_tiles ? ((m0 / BLOCK_M) / p.w_group_tiles) * p.N : 0));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (kTrace && p.trace) p.trace[blockIdx.x * 8 + 0] = (unsigned long long)tr_wait;
      pdl_trigger_late();   // every load of this CTA is issued: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && cta_rank == 0) {   // 2-CTA: only the leader issues
      constexpr uint32_t idesc_c = instr_desc<BLOCK_N, CTA2 ? 256 : BLOCK_M>();
      const uint32_t idesc_bf16 = p.a1_f16 ? (idesc_c & kIdescF16Mask) : idesc_c;     // descriptor of the first range
      const uint32_t idesc_a2 = p.a2_f16 ? (idesc_c & kIdescF16Mask) : idesc_c;
      const int k1_mma = p.taps * p.c1_chunks;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      long long tr_acc = 0, tr_full = 0;
      const long long tr_start = (kTrace && p.trace) ? clock64() : 0;
      for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        if (kTrace && p.trace) {
          const long long t0 = clock64();
          mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);
          tr_acc += clock64() - t0;
        } else {
          mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);   // epilogue has drained this accumulator stage
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * M_SUB * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          if (kTrace && p.trace) {
            const long long t0 = clock64();
            mbar_wait(full_bar(stage), phase, 1);
            tr_full += clock64() - t0;
          } else {
            mbar_wait(full_bar(stage), phase, 1);
          }
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_addr = smem_base + stage * L::STAGE_BYTES;
          const uint64_t b_desc = make_smem_desc(a_addr + L::A_BYTES);
          const uint32_t idesc = kb < k1_mma ? idesc_bf16 : idesc_a2;
#pragma unroll
          for (int sub = 0; sub < M_SUB; ++sub) {
            if (kTrace && (p.dbg & 2)) break;
            const uint64_t a_desc = make_smem_desc(a_addr + sub * L::A_SUB_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 32 B (16 bf16) inside the 128 B swizzle row: +2 in the (addr >> 4) field
              if constexpr (CTA2) umma_bf16_2cta(tmem_acc + (uint32_t)(sub * BLOCK_N), a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
              else umma_bf16(tmem_acc + (uint32_t)(sub * BLOCK_N), a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
            }
          }
          // frees the smem slot (in both CTAs of a pair) when these MMAs retire
          if constexpr (CTA2) umma_commit_2cta(empty_bar(stage));
          else umma_commit(empty_bar(stage));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if constexpr (CTA2) umma_commit_2cta(tmem_full_bar(as));   // accumulator complete (both CTAs' epilogues)
        else umma_commit(tmem_full_bar(as));
      }
      if (kTrace && p.trace) {
        p.trace[blockIdx.x * 8 + 1] = (unsigned long long)tr_acc;
        p.trace[blockIdx.x * 8 + 2] = (unsigned long long)tr_full;
        p.trace[blockIdx.x * 8 + 3] = (unsigned long long)(clock64() - tr_start);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    if constexpr (TS)
      epilogue_ts_role<BLOCK_N, (EPI & 2) != 0, CTA2>(p, &tm_out, smem_gen + L::STG_OFFSET, smem_base + L::STG_OFFSET,
                                                      tmem_base, tmem_full_bar(0), tmem_empty_bar(0), n_tiles, n_tiles_n,
                                                      work_id0, work_step, cta_rank, warp, lane);
    else
      epilogue_role<BLOCK_N, M_SUB, L::CHUNK, EPI, CTA2>(p, reinterpret_cast<float*>(smem_gen + L::STG_OFFSET),
                                             reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET),
                                             L::STAT_IMGS, tmem_base,
                                             tmem_full_bar(0), tmem_empty_bar(0), n_tiles, n_tiles_n, work_id0,
                                             work_step, cta_rank, warp, lane);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if constexpr (CTA2) cluster_sync_all();   // neither CTA may exit (or free TMEM) while its peer still uses it
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if constexpr (CTA2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(tmem_cols<BLOCK_N, M_SUB>())
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                   "n"(tmem_cols<BLOCK_N, M_SUB>())
                   : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Short-K linears (attention qkv / proj_out: K = 384 or 512, N = 3..8 tiles), "A-stationary" variant, CTA pair.
//
// The plain kernel re-fetches a row tile of A once per output-column tile; with six k-blocks per tile these linears
// are bound by exactly that L2 -> shared-memory operand delivery (the TMA-store epilogue made them no faster).  Here a
// worker (CTA pair) owns a CONTIGUOUS run of tiles with the column tile fastest, keeps the whole K extent of its
// current row tile resident (KB_MAX x 16 KB per CTA) and streams only weight tiles through the ring: A is read once
// per row tile instead of once per (row tile, column tile).  Roles, accumulator staging and both epilogue
// implementations are those of gemm_tc_kernel (they only see a tile range).
template <int BLOCK_N, int KB_MAX, int SB, int TS_ESIZE>
struct AStatLayout {
  static constexpr int CHUNK = 32;
  static constexpr int A_SUB_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int A_BYTES = KB_MAX * A_SUB_BYTES;
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;          // this CTA's half of the weight tile
  static constexpr int B_SLOT = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_OFFSET = A_BYTES;
  static constexpr int STG_OFFSET = B_OFFSET + SB * B_SLOT;
  static constexpr int STG_BYTES = TS_ESIZE ? BLOCK_M * BLOCK_N * TS_ESIZE : EPI_WARPS * 32 * (CHUNK + 4) * 4;
  static constexpr int STAT_IMGS = 2;
  static constexpr int STAT_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int STAT_BYTES = TS_ESIZE ? 0 : 4 * STAT_IMGS * 2 * BLOCK_N * 4;
  static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
  static constexpr int SBX = SB + 4;                 // weight-ring slots incl. those carved from unused A slots (K < 64 KB_MAX)
  static constexpr int NUM_BARS = 2 + 2 * SBX + 4;
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};

template <int BLOCK_N, int KB_MAX, int SB, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_astat_kernel(const __grid_constant__ CUtensorMap tm_a1,
                                                                       const __grid_constant__ CUtensorMap tm_w,
                                                                       const __grid_constant__ CUtensorMap tm_out,
                                                                       const TcParams p, const int tiles_per_worker) {
  pdl_launch_dependents_persistent();
  constexpr bool TS = (EPI & 32) != 0;
  using L = AStatLayout<BLOCK_N, KB_MAX, SB, TS ? ((EPI & 2) ? 2 : 4) : 0>;
  constexpr int TILE_M = 2 * BLOCK_M;
  const uint32_t cta_rank = cluster_ctarank();
  const int worker = (int)(blockIdx.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  const uint32_t a_full = bar_base, a_empty = bar_base + 8u;
  constexpr int SBX = L::SBX;
  auto b_full = [&](int s) { return bar_base + 8u * (2 + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 + SBX + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 + 2 * SBX + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 + 2 * SBX + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = p.c1_chunks;
  // K shorter than the reserved A region (KB_MAX k-blocks): the unused A slots become extra weight-ring slots -- the
  // ring depth, not the tensor core, is what the MMA thread of these short-K linears waits for
  const int sb_n = SB + min(SBX - SB, ((KB_MAX - num_kb) * L::A_SUB_BYTES) / L::B_SLOT);
  auto b_slot = [&](int s) {
    return s < SB ? smem_base + L::B_OFFSET + s * L::B_SLOT
                  : smem_base + (uint32_t)num_kb * L::A_SUB_BYTES + (uint32_t)(s - SB) * L::B_SLOT;
  };
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles = n_tiles_n * ((p.M + TILE_M - 1) / TILE_M);
  const int t0 = worker * tiles_per_worker;
  const int t_end = min(n_tiles, t0 + tiles_per_worker);

  for (int i = threadIdx.x; i < L::STAT_BYTES / 4; i += NUM_THREADS)
    reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 + 2 * SBX; ++s) mbar_init(bar_base + 8u * s, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS * 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(tmem_cols<BLOCK_N, 1>())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_a1)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      int cur_mt = -1, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = t0; tile < t_end; ++tile) {
        const int mt = tile / n_tiles_n;
        const int n0 = (tile - mt * n_tiles_n) * BLOCK_N;
        const int m0 = mt * TILE_M + (int)cta_rank * BLOCK_M;
        if (mt != cur_mt) {     // new row tile: its whole K extent, once
          mbar_wait(a_empty, pa ^ 1u, 0);
          if (cta_rank == 0) mbar_expect_tx(a_full, 2u * (uint32_t)num_kb * L::A_SUB_BYTES);
          for (int kb = 0; kb < num_kb; ++kb)
            tma_load_5d_2cta(smem_base + kb * L::A_SUB_BYTES, &tm_a1, a_full, kb * BLOCK_K, m0, 0, 0, 0);
          pa ^= 1u;
          cur_mt = mt;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_empty(sb), pb ^ 1u, 4);
          if (cta_rank == 0) mbar_expect_tx(b_full(sb), 2 * L::B_BYTES);
          tma_load_2d_2cta(b_slot(sb), &tm_w, b_full(sb), kb * BLOCK_K, n0 + (int)cta_rank * (BLOCK_N / 2));
          if (++sb == sb_n) {
            sb = 0;
            pb ^= 1u;
          }
        }
      }
      pdl_trigger_late();   // every load of this CTA is issued: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      const uint32_t idesc = p.a1_f16 ? (instr_desc<BLOCK_N, 256>() & kIdescF16Mask) : instr_desc<BLOCK_N, 256>();
      int cur_mt = -1, sb = 0, it = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = t0; tile < t_end; ++tile, ++it) {
        const int mt = tile / n_tiles_n;
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);
        if (mt != cur_mt) {
          mbar_wait(a_full, pa, 1);
          pa ^= 1u;
          cur_mt = mt;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(b_full(sb), pb, 5);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t a_desc = make_smem_desc(smem_base + kb * L::A_SUB_BYTES);
          const uint64_t b_desc = make_smem_desc(b_slot(sb));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16_2cta(tmem_acc, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
          umma_commit_2cta(b_empty(sb));
          if (++sb == sb_n) {
            sb = 0;
            pb ^= 1u;
          }
        }
        umma_commit_2cta(tmem_full_bar(as));
        // last tile of this row tile: the A region is free once these MMAs have retired
        if (tile + 1 >= t_end || (tile + 1) / n_tiles_n != mt) umma_commit_2cta(a_empty);
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): the tile range [t0, t_end), step 1 =====================
    if constexpr (TS)
      epilogue_ts_role<BLOCK_N, (EPI & 2) != 0, true>(p, &tm_out, smem_gen + L::STG_OFFSET, smem_base + L::STG_OFFSET,
                                                      tmem_base, tmem_full_bar(0), tmem_empty_bar(0), t_end, n_tiles_n, t0, 1,
                                                      cta_rank, warp, lane);
    else
      epilogue_role<BLOCK_N, 1, L::CHUNK, EPI, true>(p, reinterpret_cast<float*>(smem_gen + L::STG_OFFSET),
                                                     reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET), L::STAT_IMGS,
                                                     tmem_base, tmem_full_bar(0), tmem_empty_bar(0), t_end, n_tiles_n, t0, 1,
                                                     cta_rank, warp, lane);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(tmem_cols<BLOCK_N, 1>())
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// 3x3 stride-1 convolution, "halo" variant (always a CTA pair).
//
// The plain kernel fetches one shifted activation box per tap: nine L2 -> shared-memory copies of the same
// pixels, and TMA operand delivery (~10 TB/s over the chip) is what bounds it.  Here one TMA box per
// (64-channel chunk, horizontal shift dx) brings the CTA's R image rows PLUS the row above and below
// ((R + 2) x W pixels, out-of-image rows / columns zero-filled = the padding).  The three vertical taps then
// are the same shared-memory slot read at a start offset of dy * W pixel rows (W * 128 B, a multiple of the
// 1024 B swizzle atom for W >= 8), so A traffic drops from 9 to 3 * (R + 2) / R tile loads per chunk.
// Activation slots and weight tiles travel in two independent mbarrier rings (SA x A_SLOT, SB x B_SLOT);
// an optional second operand (fused 1x1 skip projection) appends plain 128-row tiles to the same rings.
template <int BLOCK_N, int M_SUB, int SA, int SB, bool ILV = false>
struct HaloLayout {
  static constexpr int CHUNK = 32;   // columns per epilogue step
  static constexpr int A_SUB_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int A_SLOT = (M_SUB + 1) * A_SUB_BYTES;   // (R + 2) * W <= M_SUB * 128 + 128 pixels for W <= 64
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the weight tile
  static constexpr int B_SLOT = (B_BYTES + 1023) / 1024 * 1024;
  static constexpr int B_OFFSET = SA * A_SLOT;
  static constexpr int STG_OFFSET = B_OFFSET + SB * B_SLOT;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * (CHUNK + 4) * 4;
  static constexpr int STAT_IMGS = ILV ? 2 : 1;                         // a halo CTA tile lies inside one image (ILV: two 8x8 images)
  static constexpr int STAT_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int STAT_BYTES = 4 * STAT_IMGS * 2 * BLOCK_N * 4;
  static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
  static constexpr int NUM_BARS = 3 * SA + 2 * SB + 4;               // + SA "slot transformed" barriers (XF)
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};

// XF: fused GroupNorm-apply.  Each CTA's activation box then signals that CTA's OWN a_full barrier (plain TMA load), its
// four transform warps rewrite the slot and arrive on the LEADER's a_ready barrier (2 x 4 arrivals), which is what the
// MMA thread waits for instead of a_full.
template <int BLOCK_N, int M_SUB, int SA, int SB, int EPI, bool ILV = false, bool XF = false>
__global__ void __launch_bounds__(XF ? NUM_THREADS_XF : NUM_THREADS, 1)
    gemm_tc_halo_kernel(const __grid_constant__ CUtensorMap tm_halo, const __grid_constant__ A2Maps tm_a2,
                        const __grid_constant__ CUtensorMap tm_w, const TcParams p) {
  pdl_launch_dependents_persistent();   // the next kernel's prologue may overlap this kernel's tail (common.cuh)
  using L = HaloLayout<BLOCK_N, M_SUB, SA, SB, ILV>;
  static_assert(!ILV || M_SUB == 1, "interleaved 8x8 tiles are 128 rows");
  constexpr int CTA_ROWS = BLOCK_M * M_SUB;
  constexpr int TILE_M = 2 * CTA_ROWS;
  const uint32_t cta_rank = cluster_ctarank();
  const int work_id0 = (int)(blockIdx.x >> 1), work_step = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  auto a_ready = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SB + 4 + s); };   // XF: slot transformed (leader's copy)
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles = n_tiles_n * ((p.M + TILE_M - 1) / TILE_M);
  // ILV: the tile is two whole 8x8 images; the box is (64 ch, 8 x, 2 images, 8 + 2 rows), so a tap's 128 rows are
  // again one contiguous run of the slot, (image row, image, x)-ordered
  const int halo_rows = ILV ? p.H + 2 : CTA_ROWS / p.W + 2;
  const uint32_t halo_bytes = (uint32_t)(halo_rows * (ILV ? 2 * p.W : p.W)) * (BLOCK_K * 2);

  for (int i = threadIdx.x; i < L::STAT_BYTES / 4; i += blockDim.x)
    reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * SA + 2 * SB; ++s) mbar_init(bar_base + 8u * s, 1);
    for (int s = 0; s < SA; ++s) mbar_init(a_ready(s), 2 * XF_WARPS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS * 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(tmem_cols<BLOCK_N, M_SUB>())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // barriers, TMEM and shared-memory tables are set up: from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      auto load_b = [&](int k_elem, int n0) {
        mbar_wait(b_empty(sb), pb ^ 1u, 4);
        if (cta_rank == 0) mbar_expect_tx(b_full(sb), 2 * L::B_BYTES);
        tma_load_2d_2cta(smem_base + L::B_OFFSET + sb * L::B_SLOT, &tm_w, b_full(sb), k_elem,
                         n0 + (int)cta_rank * (BLOCK_N / 2));
        if (++sb == SB) {
          sb = 0;
          pb ^= 1u;
        }
      };
      for (int tile = work_id0; tile < n_tiles; tile += work_step) {
        const int n0 = (tile % n_tiles_n) * BLOCK_N;
        const int m0 = (tile / n_tiles_n) * TILE_M + (int)cta_rank * CTA_ROWS;
        const int img = m0 / p.HW;
        const int y0 = (m0 - img * p.HW) / p.W;
        for (int chunk = 0; chunk < p.c1_chunks; ++chunk) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(a_empty(sa), pa ^ 1u, 0);
            if constexpr (XF) {   // this CTA's box signals this CTA's barrier: its transform warps go first
              mbar_expect_tx(a_full(sa), halo_bytes);
              if constexpr (ILV) tma_load_5d(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, dx - 1, img, -1, 0);
              else tma_load_5d(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, dx - 1, y0 - 1, 0, img);
            } else {
              if (cta_rank == 0) mbar_expect_tx(a_full(sa), 2 * halo_bytes);
              if constexpr (ILV) tma_load_5d_2cta(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, dx - 1, img, -1, 0);
              else tma_load_5d_2cta(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, dx - 1, y0 - 1, 0, img);
            }
            if (++sa == SA) {
              sa = 0;
              pa ^= 1u;
            }
            for (int dy = 0; dy < 3; ++dy) load_b(((dy * 3 + dx) * p.c1_chunks + chunk) * BLOCK_K, n0);
          }
        }
        for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {
          const bool second = chunk >= p.c2a_chunks;        // second tensor of a concatenated second range
          const CUtensorMap* a2_map = second ? &tm_a2.b : &tm_a2.a;
          const int a2_c0 = (second ? chunk - p.c2a_chunks : chunk) * BLOCK_K;
          mbar_wait(a_empty(sa), pa ^ 1u, 0);
          if constexpr (XF) {
            mbar_expect_tx(a_full(sa), M_SUB * L::A_SUB_BYTES);
            if constexpr (ILV) {
              tma_load_5d(smem_base + sa * L::A_SLOT, a2_map, a_full(sa), a2_c0, 0, img, 0, 0);
            } else {
#pragma unroll
              for (int sub = 0; sub < M_SUB; ++sub)
                tma_load_5d(smem_base + sa * L::A_SLOT + sub * L::A_SUB_BYTES, a2_map, a_full(sa), a2_c0,
                            m0 + sub * BLOCK_M, 0, 0, 0);
            }
          } else {
            if (cta_rank == 0) mbar_expect_tx(a_full(sa), 2 * M_SUB * L::A_SUB_BYTES);
            if constexpr (ILV) {   // tm_a2 is the (C2, x, image, y) view: rows land in the tile's (y, image, x) order
              tma_load_5d_2cta(smem_base + sa * L::A_SLOT, a2_map, a_full(sa), a2_c0, 0, img, 0, 0);
            } else {
#pragma unroll
              for (int sub = 0; sub < M_SUB; ++sub)
                tma_load_5d_2cta(smem_base + sa * L::A_SLOT + sub * L::A_SUB_BYTES, a2_map, a_full(sa), a2_c0,
                                 m0 + sub * BLOCK_M, 0, 0, 0);
            }
          }
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
          load_b((9 * p.c1_chunks + chunk) * BLOCK_K, n0);
        }
      }
      pdl_trigger_late();   // every load of this CTA is issued: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc_bf16 = instr_desc<BLOCK_N, 256>();
      const uint32_t idesc_a2 = p.a2_f16 ? (idesc_bf16 & kIdescF16Mask) : idesc_bf16;   // second range in IEEE half
      uint32_t idesc = idesc_bf16;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const uint32_t dy_bytes = (uint32_t)(ILV ? 2 * p.W : p.W) * (BLOCK_K * 2);   // one image row of the halo slot
      for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * M_SUB * BLOCK_N);
        uint32_t accumulate = 0;
        // one weight tile against the activation rows starting at a_addr (per sub-tile: + 128 pixel rows)
        auto mma_block = [&](uint32_t a_addr) {
          mbar_wait(b_full(sb), pb, 5);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t b_desc = make_smem_desc(smem_base + L::B_OFFSET + sb * L::B_SLOT);
#pragma unroll
          for (int sub = 0; sub < M_SUB; ++sub) {
            const uint64_t a_desc = make_smem_desc(a_addr + sub * L::A_SUB_BYTES);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_2cta(tmem_acc + (uint32_t)(sub * BLOCK_N), a_desc + 2u * k, b_desc + 2u * k, idesc,
                             accumulate | (uint32_t)k);   // each sub-tile's first MMA of the tile overwrites
          }
          accumulate = 1;
          umma_commit_2cta(b_empty(sb));
          if (++sb == SB) {
            sb = 0;
            pb ^= 1u;
          }
        };
        const int n_units = 3 * p.c1_chunks;
        idesc = idesc_bf16;
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(XF ? a_ready(sa) : a_full(sa), pa, 1);
          const uint32_t a_slot = smem_base + sa * L::A_SLOT;
          for (int dy = 0; dy < 3; ++dy) mma_block(a_slot + dy * dy_bytes);
          umma_commit_2cta(a_empty(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        idesc = idesc_a2;
        for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {
          mbar_wait(XF ? a_ready(sa) : a_full(sa), pa, 1);
          mma_block(smem_base + sa * L::A_SLOT);
          umma_commit_2cta(a_empty(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        umma_commit_2cta(tmem_full_bar(as));
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    epilogue_role<BLOCK_N, M_SUB, L::CHUNK, EPI, true, ILV>(p, reinterpret_cast<float*>(smem_gen + L::STG_OFFSET),
                                             reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET),
                                             L::STAT_IMGS, tmem_base,
                                             tmem_full_bar(0), tmem_empty_bar(0), n_tiles, n_tiles_n, work_id0,
                                             work_step, cta_rank, warp, lane);
  } else if constexpr (XF) {
    // ===================== transform warps: GroupNorm-apply (+ SiLU) on this CTA's activation slots =====================
    const int tid = (int)threadIdx.x - NUM_THREADS;
    const int log2w = 31 - __clz(p.W);
    const int slot_rows = halo_rows * (ILV ? 2 * p.W : p.W);
    const int C1 = p.c1_chunks * BLOCK_K;
    const int n_img = p.M / p.HW;
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = work_id0; tile < n_tiles; tile += work_step) {
      const int m0 = (tile / n_tiles_n) * TILE_M + (int)cta_rank * CTA_ROWS;
      // ILV: the tile is two whole images and a thread's rows all lie in image (tid >> 6) & 1 of them
      const int img = m0 / p.HW + (ILV ? ((tid >> 6) & 1) : 0);
      const int y0 = ILV ? 0 : (m0 - (m0 / p.HW) * p.HW) / p.W;
      const bool live = img < n_img;            // past the last image the box is zero fill: leave it
      const float2* const crow = p.xf_coef + (size_t)(live ? img : 0) * C1 + (tid & 7) * 8;
      for (int chunk = 0; chunk < p.c1_chunks; ++chunk) {
        float ca[8], cb[8];
        xf_load_coef(crow + chunk * BLOCK_K, ca, cb);
        for (int dx = 0; dx < 3; ++dx) {
          mbar_wait(a_full(sa), pa, 6);
          if (live)
            xf_transform_slot<ILV>(smem_gen + sa * L::A_SLOT, slot_rows, log2w, p.W, p.H, y0 - 1, dx - 1, ca, cb, p.xf_act,
                                   tid);
          xf_publish_fence();
          if (lane == 0) mbar_arrive_cluster(a_ready(sa), 0);
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
      }
      for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {   // the raw 1x1 skip operand passes through untouched
        mbar_wait(a_full(sa), pa, 6);
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_ready(sa), 0);
        if (++sa == SA) {
          sa = 0;
          pa ^= 1u;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(tmem_cols<BLOCK_N, M_SUB>())
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// Folded nearest-x2 upsample + 3x3 conv (four 2x2 parity convs on the low-resolution input), halo variant.
// A CTA pair owns 256 low-res pixels x 128 output channels x ONE horizontal parity b and keeps BOTH vertical
// parities a = 0, 1 as two accumulators: per (64-channel chunk, horizontal tap j) one halo box (the CTA's R rows
// plus the row above and below) serves all four (a, vertical tap i) combinations -- the tap of parity a reads the
// slot at a start offset of (a + i) image rows.  A traffic: 2 boxes per chunk instead of 8 tile loads.
template <int SA, int SB>
struct UpfoldHaloLayout {
  static constexpr int BLOCK_N = 128, CHUNK = 32;
  static constexpr int A_SLOT = 2 * BLOCK_M * BLOCK_K * 2;     // (R + 2) * W <= 256 pixel rows for W <= 64
  static constexpr int B_BYTES = (BLOCK_N / 2) * BLOCK_K * 2;  // this CTA's half of the weight tile
  static constexpr int B_SLOT = B_BYTES;
  static constexpr int B_OFFSET = SA * A_SLOT;
  static constexpr int STG_OFFSET = B_OFFSET + SB * B_SLOT;
  static constexpr int STG_BYTES = EPI_WARPS * 32 * (CHUNK + 4) * 4;
  static constexpr int STAT_IMGS = 1;
  static constexpr int STAT_OFFSET = STG_OFFSET + STG_BYTES;
  static constexpr int STAT_BYTES = 4 * STAT_IMGS * 2 * BLOCK_N * 4;
  static constexpr int BAR_OFFSET = STAT_OFFSET + STAT_BYTES;
  static constexpr int NUM_BARS = 2 * SA + 2 * SB + 4;
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};

template <int SA, int SB, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_upfold_halo_kernel(const __grid_constant__ CUtensorMap tm_halo,
                                                                             const __grid_constant__ CUtensorMap tm_w,
                                                                             const TcParams p) {
  pdl_launch_dependents_persistent();   // the next kernel's prologue may overlap this kernel's tail (common.cuh)
  using L = UpfoldHaloLayout<SA, SB>;
  constexpr int BLOCK_N = L::BLOCK_N, CTA_ROWS = BLOCK_M, TILE_M = 2 * CTA_ROWS;
  const uint32_t cta_rank = cluster_ctarank();
  const int work_id0 = (int)(blockIdx.x >> 1), work_step = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BLOCK_N;
  const int n_tiles = 2 * p.tiles_per_par;                    // tile = b * tiles_per_par + m_tile * n_tiles_n + n_tile
  const int halo_rows = CTA_ROWS / p.W + 2;
  const uint32_t halo_bytes = (uint32_t)(halo_rows * p.W) * (BLOCK_K * 2);

  for (int i = threadIdx.x; i < L::STAT_BYTES / 4; i += NUM_THREADS)
    reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET)[i] = 0.f;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2 * SA + 2 * SB; ++s) mbar_init(bar_base + 8u * s, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS * 2);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(const_cast<uint32_t*>(tmem_ptr_smem))),
                 "n"(tmem_cols<BLOCK_N, 2>())
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // barriers, TMEM and shared-memory tables are set up: from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = work_id0; tile < n_tiles; tile += work_step) {
        const int b = tile / p.tiles_per_par, tl = tile - b * p.tiles_per_par;
        const int n0 = (tl % n_tiles_n) * BLOCK_N;
        const int m0 = (tl / n_tiles_n) * TILE_M + (int)cta_rank * CTA_ROWS;
        const int img = m0 / p.HW;
        const int y0 = (m0 - img * p.HW) / p.W;
        for (int chunk = 0; chunk < p.c1_chunks; ++chunk) {
          for (int j = 0; j < 2; ++j) {
            mbar_wait(a_empty(sa), pa ^ 1u, 0);
            if (cta_rank == 0) mbar_expect_tx(a_full(sa), 2 * halo_bytes);
            tma_load_5d_2cta(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, b ? j : j - 1, y0 - 1, 0,
                             img);
            if (++sa == SA) {
              sa = 0;
              pa ^= 1u;
            }
            for (int ai = 0; ai < 4; ++ai) {   // (a, i) = (0,0), (0,1), (1,0), (1,1): the MMA thread's order
              const int a = ai >> 1, i = ai & 1;
              mbar_wait(b_empty(sb), pb ^ 1u, 4);
              if (cta_rank == 0) mbar_expect_tx(b_full(sb), 2 * L::B_BYTES);
              tma_load_2d_2cta(smem_base + L::B_OFFSET + sb * L::B_SLOT, &tm_w, b_full(sb),
                               ((i * 2 + j) * p.c1_chunks + chunk) * BLOCK_K,
                               (a * 2 + b) * p.N + n0 + (int)cta_rank * (BLOCK_N / 2));
              if (++sb == SB) {
                sb = 0;
                pb ^= 1u;
              }
            }
          }
        }
      }
      pdl_trigger_late();   // every load of this CTA is issued: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && cta_rank == 0) {
      constexpr uint32_t idesc = instr_desc<BLOCK_N, 256>();
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const uint32_t dy_bytes = (uint32_t)p.W * (BLOCK_K * 2);
      for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * 2 * BLOCK_N);
        uint32_t started[2] = {0u, 0u};     // per vertical parity: has its accumulator been written in this tile?
        const int n_units = 2 * p.c1_chunks;
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(a_full(sa), pa, 1);
          const uint32_t a_slot = smem_base + sa * L::A_SLOT;
#pragma unroll
          for (int ai = 0; ai < 4; ++ai) {
            const int a = ai >> 1, i = ai & 1;
            mbar_wait(b_full(sb), pb, 5);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint64_t a_desc = make_smem_desc(a_slot + (uint32_t)(a + i) * dy_bytes);
            const uint64_t b_desc = make_smem_desc(smem_base + L::B_OFFSET + sb * L::B_SLOT);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
              umma_bf16_2cta(tmem_acc + (uint32_t)(a * BLOCK_N), a_desc + 2u * k, b_desc + 2u * k, idesc,
                             started[a] | (uint32_t)k);
            started[a] = 1u;
            umma_commit_2cta(b_empty(sb));
            if (++sb == SB) {
              sb = 0;
              pb ^= 1u;
            }
          }
          umma_commit_2cta(a_empty(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        umma_commit_2cta(tmem_full_bar(as));
      }
    }
  } else {
    epilogue_role<BLOCK_N, 2, L::CHUNK, EPI, true, false, true>(
        p, reinterpret_cast<float*>(smem_gen + L::STG_OFFSET), reinterpret_cast<float*>(smem_gen + L::STAT_OFFSET),
        L::STAT_IMGS, tmem_base, tmem_full_bar(0), tmem_empty_bar(0), n_tiles, n_tiles_n, work_id0, work_step, cta_rank,
        warp, lane);
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(tmem_cols<BLOCK_N, 2>())
                 : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// 3x3 stride-1 convolution with 128 output channels, operand roles swapped ("transposed" halo kernel).
//
// With both operands in shared memory an N=128 MMA retires in ~112 cycles instead of 64, an N=256 one in ~165
// instead of 128 -- and the 64x64 level only has 128 output channels.  So here the WEIGHT tile is the M=128
// operand and 256 pixels are the N operand: D^T[cout][pixel] = W[cout][k] * X[pixel][k]^T, one 128 x 256 MMA per
// K step.  Both operands are K-major tiles exactly as before, so the halo slot (the CTA's R image rows plus the
// row above and below, the vertical tap = a start offset of dy * W pixel rows) simply becomes the B operand.
// A cluster of two CTAs works on 2 x 256 pixels with independent cta_group::1 MMAs and shares every weight tile:
// each CTA loads one half of it and multicasts it to both (w_empty collects the commits of both CTAs).
// The accumulator comes out transposed -- TMEM lane = output channel, column = pixel -- which makes the epilogue
// simple: a warp's 32 lanes are 32 consecutive channels of one pixel (one coalesced 128-byte store, no staging
// through shared memory) and the GroupNorm statistics are plain per-lane running sums.
template <int SA, int SB, bool WIDE = false>
struct HaloTLayout {
  static constexpr int PIX = 256;                                   // pixels per CTA tile = MMA N
  // (R + 2) * W pixel rows: <= 384 for W <= 64; the WIDE variant holds the 4 x 128 rows of a 128-pixel-wide image
  static constexpr int A_SLOT = (PIX / BLOCK_M + (WIDE ? 2 : 1)) * BLOCK_M * BLOCK_K * 2;
  static constexpr int W_BYTES = BLOCK_M * BLOCK_K * 2;              // 128 output channels x 64 k
  static constexpr int W_OFFSET = SA * A_SLOT;
  static constexpr int BAR_OFFSET = W_OFFSET + SB * W_BYTES;
  static constexpr int NUM_BARS = 3 * SA + 2 * SB + 4;               // + SA "slot transformed" barriers (XF)
  static constexpr int TOTAL = BAR_OFFSET + NUM_BARS * 8 + 16 + 1024;
};

template <int SA, int SB, int EPI, bool WIDE = false, bool XF = false>
__global__ void __launch_bounds__(XF ? NUM_THREADS_XF : NUM_THREADS, 1)
    gemm_tc_halo_t_kernel(const __grid_constant__ CUtensorMap tm_halo, const __grid_constant__ A2Maps tm_a2,
                          const __grid_constant__ CUtensorMap tm_w, const TcParams p) {
  pdl_launch_dependents_persistent();   // the next kernel's prologue may overlap this kernel's tail (common.cuh)
  // image-pipelined GroupNorm-apply: the dependent polls img_done, so it must be released now, while this grid's CTAs
  // are all resident and can only make progress (vdm_gemm_args.img_done)
  if (p.img_done != nullptr) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  using L = HaloTLayout<SA, SB, WIDE>;
  constexpr int PIX = L::PIX, TILE_M = 2 * PIX;
  constexpr bool HAS_RES = (EPI & 1) != 0, BF16_OUT = (EPI & 2) != 0, STATS = (EPI & 4) != 0;
  constexpr bool F16IO = (EPI & 64) != 0;     // out_f32 / residual are fp16 tensors (the model's residual stream)
  const uint32_t cta_rank = cluster_ctarank();
  const int work_id0 = (int)(blockIdx.x >> 1), work_step = (int)(gridDim.x >> 1);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + L::BAR_OFFSET;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (SA + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto tmem_full_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto tmem_empty_bar = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  auto a_ready = [&](int s) { return bar_base + 8u * (2 * SA + 2 * SB + 4 + s); };   // XF: slot transformed
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(smem_gen + L::BAR_OFFSET + L::NUM_BARS * 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = p.N / BLOCK_M;                       // 128 output channels per tile; n fastest, so the
  const int n_tiles = n_tiles_n * ((p.M + TILE_M - 1) / TILE_M);   // CTAs of a wave share the activation rows in L2
  const int halo_rows = PIX / p.W + 2;
  const uint32_t halo_bytes = (uint32_t)(halo_rows * p.W) * (BLOCK_K * 2);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(w_full(s), 1);
      mbar_init(w_empty(s), 2);       // both CTAs of the cluster must have consumed the tile
    }
    for (int s = 0; s < SA; ++s) mbar_init(a_ready(s), XF_WARPS);
    for (int a = 0; a < 2; ++a) {
      mbar_init(tmem_full_bar(a), 1);
      mbar_init(tmem_empty_bar(a), EPI_WARPS);
    }
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
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();       // peer barriers exist before any multicast load / commit can reach them
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();   // barriers, TMEM and shared-memory tables are set up: from here on global memory is touched

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_halo)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tm_w)) : "memory");
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int n0 = 0;
      auto load_w = [&](int k_elem) {     // this CTA's 64 output channels of the tile, to both CTAs
        mbar_wait(w_empty(sb), pb ^ 1u, 4);
        mbar_expect_tx(w_full(sb), L::W_BYTES);
        tma_load_2d_multicast(smem_base + L::W_OFFSET + sb * L::W_BYTES + cta_rank * (L::W_BYTES / 2), &tm_w, w_full(sb),
                              k_elem, n0 + (int)cta_rank * (BLOCK_M / 2), (uint16_t)3);
        if (++sb == SB) {
          sb = 0;
          pb ^= 1u;
        }
      };
      for (int tile = work_id0; tile < n_tiles; tile += work_step) {
        n0 = (tile % n_tiles_n) * BLOCK_M;
        const int m0 = (tile / n_tiles_n) * TILE_M + (int)cta_rank * PIX;
        const int img = m0 / p.HW;
        const int y0 = (m0 - img * p.HW) / p.W;
        for (int chunk = 0; chunk < p.c1_chunks; ++chunk) {
          for (int dx = 0; dx < 3; ++dx) {
            mbar_wait(a_empty(sa), pa ^ 1u, 0);
            mbar_expect_tx(a_full(sa), halo_bytes);
            tma_load_5d(smem_base + sa * L::A_SLOT, &tm_halo, a_full(sa), chunk * BLOCK_K, dx - 1, y0 - 1, 0, img);
            if (++sa == SA) {
              sa = 0;
              pa ^= 1u;
            }
            for (int dy = 0; dy < 3; ++dy) load_w(((dy * 3 + dx) * p.c1_chunks + chunk) * BLOCK_K);
          }
        }
        for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {
          const bool second = chunk >= p.c2a_chunks;        // second tensor of a concatenated second range
          const CUtensorMap* a2_map = second ? &tm_a2.b : &tm_a2.a;
          const int a2_c0 = (second ? chunk - p.c2a_chunks : chunk) * BLOCK_K;
          mbar_wait(a_empty(sa), pa ^ 1u, 0);
          mbar_expect_tx(a_full(sa), PIX * BLOCK_K * 2);
#pragma unroll
          for (int sub = 0; sub < PIX / BLOCK_M; ++sub)
            tma_load_5d(smem_base + sa * L::A_SLOT + sub * (BLOCK_M * BLOCK_K * 2), a2_map, a_full(sa), a2_c0,
                        m0 + sub * BLOCK_M, 0, 0, 0);
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
          load_w((9 * p.c1_chunks + chunk) * BLOCK_K);
        }
      }
      pdl_trigger_late();   // every load of this CTA is issued: the next kernel may start launching
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc_bf16 = instr_desc<PIX, BLOCK_M>();    // M = 128 output channels, N = 256 pixels
      const uint32_t idesc_a2 = p.a2_f16 ? (idesc_bf16 & kIdescF16Mask) : idesc_bf16;   // second range in IEEE half
      uint32_t idesc = idesc_bf16;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const uint32_t dy_bytes = (uint32_t)p.W * (BLOCK_K * 2);
      for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(tmem_empty_bar(as), aphase ^ 1u, 3);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_acc = tmem_base + (uint32_t)(as * PIX);
        uint32_t accumulate = 0;
        auto mma_block = [&](uint32_t pix_addr) {
          mbar_wait(w_full(sb), pb, 5);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t w_desc = make_smem_desc(smem_base + L::W_OFFSET + sb * L::W_BYTES);
          const uint64_t x_desc = make_smem_desc(pix_addr);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
            umma_bf16(tmem_acc, w_desc + 2u * k, x_desc + 2u * k, idesc, accumulate | (uint32_t)k);
          accumulate = 1;
          umma_commit_multicast(w_empty(sb), (uint16_t)3);
          if (++sb == SB) {
            sb = 0;
            pb ^= 1u;
          }
        };
        const int n_units = 3 * p.c1_chunks;
        idesc = idesc_bf16;
        for (int u = 0; u < n_units; ++u) {
          mbar_wait(XF ? a_ready(sa) : a_full(sa), pa, 1);   // XF: the transform warps hand the slot over
          const uint32_t a_slot = smem_base + sa * L::A_SLOT;
          for (int dy = 0; dy < 3; ++dy) mma_block(a_slot + dy * dy_bytes);
          umma_commit(a_empty(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        idesc = idesc_a2;
        for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {
          mbar_wait(XF ? a_ready(sa) : a_full(sa), pa, 1);   // XF: the transform warps hand the slot over
          mma_block(smem_base + sa * L::A_SLOT);
          umma_commit(a_empty(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
        umma_commit(tmem_full_bar(as));
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ===================== epilogue (warps 2..9): lane = output channel, TMEM column = pixel =====================
    const int ew = warp - 2;
    const int q = warp & 3;                 // TMEM lane quarter: channels q*32 .. q*32+31
    const int half = ew >> 2;               // pixel columns half*128 .. +127
    int it = 0;
    for (int tile = work_id0; tile < n_tiles; tile += work_step, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int c = (tile % n_tiles_n) * BLOCK_M + q * 32 + lane;
      const float bias_c = p.bias ? __ldg(p.bias + c) : 0.f;
      const int row0 = (tile / n_tiles_n) * TILE_M + (int)cta_rank * PIX + half * 128;   // first row of this warp
      float rsum = 0.f, rsq = 0.f;
      float res_cur[32];
      auto load_res = [&](float (&dst)[32], int chunk) {
        if constexpr (HAS_RES && F16IO) {
          const uint16_t* rp = reinterpret_cast<const uint16_t*>(p.residual) + (size_t)(row0 + chunk * 32) * p.ld_res + c;
#pragma unroll
          for (int i = 0; i < 32; ++i)     // raw bits; converted at the use (keeps all 32 loads in flight)
            dst[i] = __uint_as_float((row0 + chunk * 32 + i < p.M) ? (uint32_t)__ldg(rp + (size_t)i * p.ld_res) : 0u);
        } else if constexpr (HAS_RES) {
          const float* rp = p.residual + (size_t)(row0 + chunk * 32) * p.ld_res + c;
#pragma unroll
          for (int i = 0; i < 32; ++i) dst[i] = (row0 + chunk * 32 + i < p.M) ? __ldg(rp + (size_t)i * p.ld_res) : 0.f;
        }
      };
      if constexpr (HAS_RES) {   // this warp's residual block (128 rows x 128 B) into L2 while the MMAs still run
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = row0 + j * 32 + lane;
          if (r < p.M)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.residual) +
                                                          ((size_t)r * p.ld_res + c - lane) * (F16IO ? 2 : 4)));
        }
      }
      load_res(res_cur, 0);
      mbar_wait(tmem_full_bar(as), aphase, 2);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int chunk = 0; chunk < 4; ++chunk) {
        uint32_t acc[32];
        tmem_ld_32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * PIX + half * 128 + chunk * 32), acc);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float res_next[32];
        if (chunk + 1 < 4) load_res(res_next, chunk + 1);
        const int r = row0 + chunk * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (r + i < p.M) {
            float v = __uint_as_float(acc[i]) + bias_c;
            if constexpr (HAS_RES && F16IO) v += f16_bits_to_f32((uint16_t)__float_as_uint(res_cur[i]));
            else if constexpr (HAS_RES) v += res_cur[i];
            if constexpr (STATS) {
              rsum += v;
              rsq = fmaf(v, v, rsq);
            }
            if constexpr (BF16_OUT) p.out_bf16[(size_t)(r + i) * p.ld_out_bf16 + c] = __float2bfloat16_rn(v);
            else if constexpr (F16IO) reinterpret_cast<uint16_t*>(p.out_f32)[(size_t)(r + i) * p.ld_out + c] = f32_to_f16_bits(v);
            else p.out_f32[(size_t)(r + i) * p.ld_out + c] = v;
          }
        }
        if constexpr (HAS_RES) {
          if (chunk + 1 < 4) {
#pragma unroll
            for (int i = 0; i < 32; ++i) res_cur[i] = res_next[i];
          }
        }
        if constexpr (STATS) {
          // 64-pixel images (8x8 level, linears only): the warp's first 64 pixels are an image of their own
          if (chunk == 1 && p.HW == 64) {
            if (p.stats_out != nullptr && row0 < p.M) {
              unsigned long long* tab = reinterpret_cast<unsigned long long*>(p.stats_out) + (size_t)(row0 / 64) * 2 * p.N + c;
              atomicAdd(tab, (unsigned long long)__float2ll_rn(rsum * 16777216.0f));
              atomicAdd(tab + p.N, (unsigned long long)__float2ll_rn(rsq * 16777216.0f));
            }
            rsum = 0.f;
            rsq = 0.f;
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(as));
      if constexpr (STATS) {
        // this warp's 128 pixels (or, with 64-pixel images, its last 64) lie in one image: two fixed-point atomics per channel
        const int srow = p.HW == 64 ? row0 + 64 : row0;
        if (p.stats_out != nullptr && srow < p.M) {
          unsigned long long* tab = reinterpret_cast<unsigned long long*>(p.stats_out) + (size_t)(srow / p.HW) * 2 * p.N + c;
          atomicAdd(tab, (unsigned long long)__float2ll_rn(rsum * 16777216.0f));
          atomicAdd(tab + p.N, (unsigned long long)__float2ll_rn(rsq * 16777216.0f));
        }
      }
      if (p.img_done != nullptr && row0 < p.M) {
        // this warp's 128 rows x 32 channels (one image: H*W % 128 == 0 here) and their statistics are out: publish
        __threadfence();
        __syncwarp();
        if (lane == 0) atomicAdd(p.img_done + row0 / p.HW, (unsigned int)(min(128, p.M - row0) * 32));
      }
    }
  } else if constexpr (XF) {
    // ===================== transform warps: GroupNorm-apply (+ SiLU) on every activation slot =====================
    const int tid = (int)threadIdx.x - NUM_THREADS;
    const int log2w = 31 - __clz(p.W);
    const int slot_rows = halo_rows * p.W;
    const int C1 = p.c1_chunks * BLOCK_K;
    int sa = 0;
    uint32_t pa = 0;
    for (int tile = work_id0; tile < n_tiles; tile += work_step) {
      const int m0 = (tile / n_tiles_n) * TILE_M + (int)cta_rank * PIX;
      const int img = m0 / p.HW;
      const int y0 = (m0 - img * p.HW) / p.W;
      const float2* const crow = p.xf_coef + (size_t)min(img, p.M / p.HW - 1) * C1 + (tid & 7) * 8;
      for (int chunk = 0; chunk < p.c1_chunks; ++chunk) {
        float ca[8], cb[8];
        xf_load_coef(crow + chunk * BLOCK_K, ca, cb);
        for (int dx = 0; dx < 3; ++dx) {
          mbar_wait(a_full(sa), pa, 6);
          xf_transform_slot<false>(smem_gen + sa * L::A_SLOT, slot_rows, log2w, p.W, p.H, y0 - 1, dx - 1, ca, cb, p.xf_act,
                                   tid);
          xf_publish_fence();
          if (lane == 0) mbar_arrive(a_ready(sa));
          if (++sa == SA) {
            sa = 0;
            pa ^= 1u;
          }
        }
      }
      for (int chunk = 0; chunk < p.c2_chunks; ++chunk) {   // the raw 1x1 skip operand passes through untouched
        mbar_wait(a_full(sa), pa, 6);
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready(sa));
        if (++sa == SA) {
          sa = 0;
          pa ^= 1u;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();       // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// ---------------------------------------------------------------------------------------
// host side: tensor maps

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  });
  return fn;
}

// bf16 tensor of up to 5 dims (innermost first), 128B swizzle, zero OOB fill.
int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
               const uint32_t* box, CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16) {
  auto fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return -2;
  }
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstrides[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdims, gstrides, gbox, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rank=%d dims=[%llu,%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              (unsigned long long)(rank > 4 ? dims[4] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0,
              rank > 3 ? box[3] : 0, rank > 4 ? box[4] : 0);
    return -3;
  }
  return 0;
}

// [rows][C] matrix (row stride ld elements) viewed as 5-D (C, rows, 1, 1, 1) with a (64, 128, 1, 1, 1) box.
int encode_rows_map(CUtensorMap* map, const void* base, int64_t rows, int64_t C, int64_t ld = 0) {
  if (ld == 0) ld = C;
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)rows, 1, 1, 1};
  uint64_t st[5] = {2, (uint64_t)ld * 2, (uint64_t)ld * 2 * rows, (uint64_t)ld * 2 * rows, (uint64_t)ld * 2 * rows};
  uint32_t box[5] = {BLOCK_K, BLOCK_M, 1, 1, 1};
  return encode_map(map, base, 5, dims, st, box);
}

template <int BLOCK_N, int M_SUB, int STAGES, int EPI, bool CTA2 = false>
int launch_inst(const CUtensorMap& ma1, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                cudaStream_t stream, const CUtensorMap* mout = nullptr) {
  constexpr bool TS = (EPI & 32) != 0;
  using L = SmemLayout<BLOCK_N, M_SUB, STAGES, CTA2, TS ? ((EPI & 2) ? 2 : 4) : 0>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BLOCK_N, M_SUB, STAGES, EPI, CTA2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gemm_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  constexpr int TILE_ROWS = BLOCK_M * M_SUB * (CTA2 ? 2 : 1);
  const int tiles = p.n_par * ((p.N + BLOCK_N - 1) / BLOCK_N) * ((p.M + TILE_ROWS - 1) / TILE_ROWS);
  const CUtensorMap& mo = mout ? *mout : mw;      // only read by the TMA-store epilogue
  if constexpr (CTA2) {
    // one CTA pair (a 2-CTA cluster on one TPC) per work tile; persistent over pairs
    const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
    launch_kernel(gemm_tc_kernel<BLOCK_N, M_SUB, STAGES, EPI, CTA2>, 2 * pairs, NUM_THREADS, L::TOTAL, stream, 2, ma1, ma2,
                  mw, mo, p);
  } else {
    const int grid = tiles < num_sms() ? tiles : num_sms();
    launch_kernel(gemm_tc_kernel<BLOCK_N, M_SUB, STAGES, EPI, CTA2>, grid, NUM_THREADS, L::TOTAL, stream, 1, ma1, ma2, mw, mo,
                  p);
  }
  VDM_AFTER_LAUNCH("gemm_tc");
  return 0;
}

// epilogue variant for this call (see the kernel's EPI parameter)
int epilogue_variant(const TcParams& p, int block_n) {
  const bool one_out = (p.out_f32 != nullptr) != (p.out_bf16 != nullptr);
  const bool fast = !p.out_nchw && !p.rowbias && p.a1_mode != 3 && one_out && p.N % block_n == 0;
  if (!fast) return 8 | (p.residual ? 1 : 0);
  // fp16 stream IO (bit 6): lean instances exist for the fp16-output combinations 64, 65, 68, 69
  return (p.residual ? 1 : 0) | (p.out_bf16 ? 2 : 0) | (p.stats_out ? 4 : 0) | (p.io_f16 ? 64 : 0);
}

template <int BLOCK_N, int KB_MAX, int SB, int EPI>
int launch_astat(const CUtensorMap& ma1, const CUtensorMap& mw, const CUtensorMap& mo, const TcParams& p,
                 cudaStream_t stream) {
  constexpr bool TS = (EPI & 32) != 0;
  using L = AStatLayout<BLOCK_N, KB_MAX, SB, TS ? ((EPI & 2) ? 2 : 4) : 0>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_astat_kernel<BLOCK_N, KB_MAX, SB, EPI>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gemm_tc (A-stationary): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  const int n_tiles = (p.N / BLOCK_N) * ((p.M + 2 * BLOCK_M - 1) / (2 * BLOCK_M));
  int workers = n_tiles < num_sms() / 2 ? n_tiles : num_sms() / 2;
  const int tpw = (n_tiles + workers - 1) / workers;      // contiguous tiles per worker, column tile fastest
  workers = (n_tiles + tpw - 1) / tpw;
  launch_kernel(gemm_tc_astat_kernel<BLOCK_N, KB_MAX, SB, EPI>, 2 * workers, NUM_THREADS, L::TOTAL, stream, 2, ma1, mw, mo, p,
                tpw);
  VDM_AFTER_LAUNCH("gemm_tc_astat");
  return 0;
}

int bad_variant(int v) {
  set_error("gemm_tc: epilogue combination %d is not built (fp16 stream IO goes with an fp16 output: variants 64, 65, 68, 69)", v);
  return -1;
}

// The TMA-store epilogue (variants 32 / 34) replaces the lean bias-only variants 0 / 2 wherever its output-tile staging
// fits next to the operand ring; VDM_GEMM_TS=0 switches it off (tests compare the two).
template <int BLOCK_N, int M_SUB, int STAGES, bool CTA2, int ESZ>
constexpr bool ts_fits() {
  return M_SUB == 1 && BLOCK_N >= 64 && SmemLayout<BLOCK_N, M_SUB, STAGES, CTA2, ESZ>::TOTAL <= 232448;
}
template <int BLOCK_N, int ESZ>
int encode_out_map(CUtensorMap* mo, const TcParams& p) {
  void* base = ESZ == 2 ? (void*)p.out_bf16 : (void*)p.out_f32;
  const uint64_t ld = ESZ == 2 ? p.ld_out_bf16 : p.ld_out;
  if (ld * ESZ % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return 1;   // not TMA-addressable
  const uint64_t par_stride = p.n_par > 1 ? (uint64_t)p.par_out_stride : (uint64_t)p.M * ld;
  if (par_stride * ESZ % 16 != 0) return 1;
  uint64_t dims[3] = {(uint64_t)p.N, (uint64_t)p.M, (uint64_t)p.n_par};
  uint64_t st[3] = {(uint64_t)ESZ, ld * ESZ, par_stride * ESZ};
  uint32_t box[3] = {128 / ESZ, BLOCK_M, 1};
  return encode_map(mo, base, 3, dims, st, box, ESZ == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
}

template <int BLOCK_N, int M_SUB, int STAGES, bool CTA2 = false>
int launch(const CUtensorMap& ma1, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
           cudaStream_t stream) {
  const int v = epilogue_variant(p, BLOCK_N);
  const char* ts_env = getenv("VDM_GEMM_TS");
  const bool ts_on = ts_env == nullptr || atoi(ts_env) != 0;
  if (ts_on && v == 2) {
    if constexpr (ts_fits<BLOCK_N, M_SUB, STAGES, CTA2, 2>()) {
      CUtensorMap mo;
      const int rc = encode_out_map<BLOCK_N, 2>(&mo, p);
      if (rc < 0) return rc;
      if (rc == 0) return launch_inst<BLOCK_N, M_SUB, STAGES, 34, CTA2>(ma1, ma2, mw, p, stream, &mo);
    }
  }
  if (ts_on && v == 0) {
    if constexpr (ts_fits<BLOCK_N, M_SUB, STAGES, CTA2, 4>()) {
      CUtensorMap mo;
      const int rc = encode_out_map<BLOCK_N, 4>(&mo, p);
      if (rc < 0) return rc;
      if (rc == 0) return launch_inst<BLOCK_N, M_SUB, STAGES, 32, CTA2>(ma1, ma2, mw, p, stream, &mo);
    }
  }
  switch (v) {
    case 0: return launch_inst<BLOCK_N, M_SUB, STAGES, 0, CTA2>(ma1, ma2, mw, p, stream);
    case 1: return launch_inst<BLOCK_N, M_SUB, STAGES, 1, CTA2>(ma1, ma2, mw, p, stream);
    case 2: return launch_inst<BLOCK_N, M_SUB, STAGES, 2, CTA2>(ma1, ma2, mw, p, stream);
    case 3: return launch_inst<BLOCK_N, M_SUB, STAGES, 3, CTA2>(ma1, ma2, mw, p, stream);
    case 4: return launch_inst<BLOCK_N, M_SUB, STAGES, 4, CTA2>(ma1, ma2, mw, p, stream);
    case 5: return launch_inst<BLOCK_N, M_SUB, STAGES, 5, CTA2>(ma1, ma2, mw, p, stream);
    case 6: return launch_inst<BLOCK_N, M_SUB, STAGES, 6, CTA2>(ma1, ma2, mw, p, stream);
    case 7: return launch_inst<BLOCK_N, M_SUB, STAGES, 7, CTA2>(ma1, ma2, mw, p, stream);
    case 9: return launch_inst<BLOCK_N, M_SUB, STAGES, 9, CTA2>(ma1, ma2, mw, p, stream);
    case 64: return launch_inst<BLOCK_N, M_SUB, STAGES, 64, CTA2>(ma1, ma2, mw, p, stream);
    case 65: return launch_inst<BLOCK_N, M_SUB, STAGES, 65, CTA2>(ma1, ma2, mw, p, stream);
    case 68: return launch_inst<BLOCK_N, M_SUB, STAGES, 68, CTA2>(ma1, ma2, mw, p, stream);
    case 69: return launch_inst<BLOCK_N, M_SUB, STAGES, 69, CTA2>(ma1, ma2, mw, p, stream);
    case 8: return launch_inst<BLOCK_N, M_SUB, STAGES, 8, CTA2>(ma1, ma2, mw, p, stream);
    default: return bad_variant(v);
  }
}

template <int BLOCK_N, int M_SUB, int SA, int SB, int EPI, bool ILV = false, bool XF = false>
int launch_halo_inst(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                     cudaStream_t stream) {
  using L = HaloLayout<BLOCK_N, M_SUB, SA, SB, ILV>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_halo_kernel<BLOCK_N, M_SUB, SA, SB, EPI, ILV, XF>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gemm_tc (halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  constexpr int TILE_ROWS = 2 * BLOCK_M * M_SUB;
  const int tiles = (p.N / BLOCK_N) * ((p.M + TILE_ROWS - 1) / TILE_ROWS);
  const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(XF ? NUM_THREADS_XF : NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {     // programmatic dependent launch (common.cuh)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_halo_kernel<BLOCK_N, M_SUB, SA, SB, EPI, ILV, XF>, mh, ma2, mw, p);
  if (e != cudaSuccess) {
    set_error("gemm_tc (halo): launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  VDM_AFTER_LAUNCH("gemm_tc_halo");
  return 0;
}

// fused-normalisation instances: one output + GroupNorm statistics, with or without residual (variants 4..7)
template <int BLOCK_N, int M_SUB, int SA, int SB, bool ILV>
int launch_halo_xf(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                   cudaStream_t stream) {
  switch (epilogue_variant(p, BLOCK_N)) {
    case 4: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 4, ILV, true>(mh, ma2, mw, p, stream);
    case 5: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 5, ILV, true>(mh, ma2, mw, p, stream);
    case 6: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 6, ILV, true>(mh, ma2, mw, p, stream);
    case 7: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 7, ILV, true>(mh, ma2, mw, p, stream);
    default: set_error("gemm_tc: fused normalisation needs one output and stats_out"); return -1;
  }
}

// interleaved 8x8 tiles: lean epilogue variants only (the caller checked epilogue_variant < 8)
template <int BLOCK_N, int SA, int SB>
int launch_halo_ilv(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                    cudaStream_t stream) {
  if (p.xf_coef != nullptr) return launch_halo_xf<BLOCK_N, 1, SA, SB, true>(mh, ma2, mw, p, stream);
  switch (epilogue_variant(p, BLOCK_N)) {
    case 0: return launch_halo_inst<BLOCK_N, 1, SA, SB, 0, true>(mh, ma2, mw, p, stream);
    case 1: return launch_halo_inst<BLOCK_N, 1, SA, SB, 1, true>(mh, ma2, mw, p, stream);
    case 2: return launch_halo_inst<BLOCK_N, 1, SA, SB, 2, true>(mh, ma2, mw, p, stream);
    case 3: return launch_halo_inst<BLOCK_N, 1, SA, SB, 3, true>(mh, ma2, mw, p, stream);
    case 4: return launch_halo_inst<BLOCK_N, 1, SA, SB, 4, true>(mh, ma2, mw, p, stream);
    case 5: return launch_halo_inst<BLOCK_N, 1, SA, SB, 5, true>(mh, ma2, mw, p, stream);
    case 6: return launch_halo_inst<BLOCK_N, 1, SA, SB, 6, true>(mh, ma2, mw, p, stream);
    case 7: return launch_halo_inst<BLOCK_N, 1, SA, SB, 7, true>(mh, ma2, mw, p, stream);
    case 64: return launch_halo_inst<BLOCK_N, 1, SA, SB, 64, true>(mh, ma2, mw, p, stream);
    case 65: return launch_halo_inst<BLOCK_N, 1, SA, SB, 65, true>(mh, ma2, mw, p, stream);
    case 68: return launch_halo_inst<BLOCK_N, 1, SA, SB, 68, true>(mh, ma2, mw, p, stream);
    case 69: return launch_halo_inst<BLOCK_N, 1, SA, SB, 69, true>(mh, ma2, mw, p, stream);
    default: return bad_variant(epilogue_variant(p, BLOCK_N));
  }
}

template <int BLOCK_N, int M_SUB, int SA, int SB>
int launch_halo(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                cudaStream_t stream) {
  if (p.xf_coef != nullptr) return launch_halo_xf<BLOCK_N, M_SUB, SA, SB, false>(mh, ma2, mw, p, stream);
  switch (epilogue_variant(p, BLOCK_N)) {
    case 0: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 0>(mh, ma2, mw, p, stream);
    case 1: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 1>(mh, ma2, mw, p, stream);
    case 2: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 2>(mh, ma2, mw, p, stream);
    case 3: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 3>(mh, ma2, mw, p, stream);
    case 4: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 4>(mh, ma2, mw, p, stream);
    case 5: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 5>(mh, ma2, mw, p, stream);
    case 6: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 6>(mh, ma2, mw, p, stream);
    case 7: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 7>(mh, ma2, mw, p, stream);
    case 9: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 9>(mh, ma2, mw, p, stream);
    case 64: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 64>(mh, ma2, mw, p, stream);
    case 65: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 65>(mh, ma2, mw, p, stream);
    case 68: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 68>(mh, ma2, mw, p, stream);
    case 69: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 69>(mh, ma2, mw, p, stream);
    case 8: return launch_halo_inst<BLOCK_N, M_SUB, SA, SB, 8>(mh, ma2, mw, p, stream);
    default: return bad_variant(epilogue_variant(p, BLOCK_N));
  }
}

template <int SA, int SB, int EPI>
int launch_upfold_halo_inst(const CUtensorMap& mh, const CUtensorMap& mw, const TcParams& p, cudaStream_t stream) {
  using L = UpfoldHaloLayout<SA, SB>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_upfold_halo_kernel<SA, SB, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gemm_tc (upfold halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  const int tiles = 2 * p.tiles_per_par;
  const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {     // programmatic dependent launch (common.cuh)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_upfold_halo_kernel<SA, SB, EPI>, mh, mw, p);
  if (e != cudaSuccess) {
    set_error("gemm_tc (upfold halo): launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  VDM_AFTER_LAUNCH("gemm_tc_upfold_halo");
  return 0;
}

// lean epilogues (one output, N a multiple of 128) when possible, else the generic one
template <int SA, int SB>
int launch_upfold_halo(const CUtensorMap& mh, const CUtensorMap& mw, const TcParams& p, cudaStream_t stream) {
  const bool one_out = (p.out_f32 != nullptr) != (p.out_bf16 != nullptr);
  if (one_out && !p.rowbias && !p.residual) {
    const int v = (p.out_bf16 ? 2 : 0) | (p.stats_out ? 4 : 0) | ((p.io_f16 && !p.out_bf16) ? 64 : 0);
    switch (v) {
      case 0: return launch_upfold_halo_inst<SA, SB, 0>(mh, mw, p, stream);
      case 2: return launch_upfold_halo_inst<SA, SB, 2>(mh, mw, p, stream);
      case 4: return launch_upfold_halo_inst<SA, SB, 4>(mh, mw, p, stream);
      case 64: return launch_upfold_halo_inst<SA, SB, 64>(mh, mw, p, stream);
      case 68: return launch_upfold_halo_inst<SA, SB, 68>(mh, mw, p, stream);
      default: return launch_upfold_halo_inst<SA, SB, 6>(mh, mw, p, stream);
    }
  }
  return launch_upfold_halo_inst<SA, SB, 8>(mh, mw, p, stream);
}

template <int SA, int SB, int EPI, bool WIDE, bool XF = false>
int launch_halo_t_inst(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                       cudaStream_t stream) {
  using L = HaloTLayout<SA, SB, WIDE>;
  static_assert(L::TOTAL <= 232448, "shared memory budget exceeded");
  static PerDevice<bool> configured;
  if (!configured.get()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_halo_t_kernel<SA, SB, EPI, WIDE, XF>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, L::TOTAL);
    if (e != cudaSuccess) {
      set_error("gemm_tc (transposed halo): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return (int)e;
    }
    configured.get() = true;
  }
  const int tiles = (p.N / BLOCK_M) * ((p.M + 2 * L::PIX - 1) / (2 * L::PIX));
  const int pairs = tiles < num_sms() / 2 ? tiles : num_sms() / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(XF ? NUM_THREADS_XF : NUM_THREADS);
  cfg.dynamicSmemBytes = L::TOTAL;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (pdl_enabled()) {     // programmatic dependent launch (common.cuh)
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.numAttrs = 2;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_halo_t_kernel<SA, SB, EPI, WIDE, XF>, mh, ma2, mw, p);
  if (e != cudaSuccess) {
    set_error("gemm_tc (transposed halo): launch failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  VDM_AFTER_LAUNCH("gemm_tc_halo_t");
  return 0;
}

// epilogue variants the fused-normalisation (XF) kernels are built for: one output + GroupNorm statistics, with or
// without a residual -- what the ResBlock convs ask for (other combinations keep the standalone GroupNorm-apply)
bool xf_variant_ok(int v) { return v >= 4 && v <= 7; }

template <int SA, int SB, bool WIDE = false>
int launch_halo_t(const CUtensorMap& mh, const A2Maps& ma2, const CUtensorMap& mw, const TcParams& p,
                  cudaStream_t stream) {
  const int v = epilogue_variant(p, 128);
  if (p.xf_coef != nullptr) {
    switch (v) {
      case 4: return launch_halo_t_inst<SA, SB, 4, WIDE, true>(mh, ma2, mw, p, stream);
      case 5: return launch_halo_t_inst<SA, SB, 5, WIDE, true>(mh, ma2, mw, p, stream);
      case 6: return launch_halo_t_inst<SA, SB, 6, WIDE, true>(mh, ma2, mw, p, stream);
      case 7: return launch_halo_t_inst<SA, SB, 7, WIDE, true>(mh, ma2, mw, p, stream);
      default: set_error("gemm_tc: fused normalisation needs one output and stats_out"); return -1;
    }
  }
  switch (v) {
    case 0: return launch_halo_t_inst<SA, SB, 0, WIDE>(mh, ma2, mw, p, stream);
    case 1: return launch_halo_t_inst<SA, SB, 1, WIDE>(mh, ma2, mw, p, stream);
    case 2: return launch_halo_t_inst<SA, SB, 2, WIDE>(mh, ma2, mw, p, stream);
    case 3: return launch_halo_t_inst<SA, SB, 3, WIDE>(mh, ma2, mw, p, stream);
    case 4: return launch_halo_t_inst<SA, SB, 4, WIDE>(mh, ma2, mw, p, stream);
    case 5: return launch_halo_t_inst<SA, SB, 5, WIDE>(mh, ma2, mw, p, stream);
    case 6: return launch_halo_t_inst<SA, SB, 6, WIDE>(mh, ma2, mw, p, stream);
    case 7: return launch_halo_t_inst<SA, SB, 7, WIDE>(mh, ma2, mw, p, stream);
    case 64: return launch_halo_t_inst<SA, SB, 64, WIDE>(mh, ma2, mw, p, stream);
    case 65: return launch_halo_t_inst<SA, SB, 65, WIDE>(mh, ma2, mw, p, stream);
    case 68: return launch_halo_t_inst<SA, SB, 68, WIDE>(mh, ma2, mw, p, stream);
    case 69: return launch_halo_t_inst<SA, SB, 69, WIDE>(mh, ma2, mw, p, stream);
    default: return bad_variant(v);
  }
}

}  // namespace

int gemm_tc_upfold(const vdm_gemm_args* a, cudaStream_t stream);
int conv3x3_small_n(const vdm_gemm_args* a, cudaStream_t stream);   // conv_small_n.cu; -100 = shape not handled

static unsigned long long* g_trace_buf = nullptr;
void gemm_tc_set_trace(void* buf) { g_trace_buf = reinterpret_cast<unsigned long long*>(buf); }

// probe != nullptr: do not launch; *probe = 1 if the call would run on a halo kernel with a lean epilogue, i.e. on a
// kernel that has the fused-normalisation transform stage (vdm_gemm_fused_norm_supported)
int gemm_tc(const vdm_gemm_args* a, cudaStream_t stream, int* probe) {
  const int64_t M = (int64_t)a->n_img * a->H * a->W;
  if (probe) *probe = 0;
  const bool xf = a->a1_coef != nullptr;
  VDM_REQUIRE(!xf || (a->taps == 9 && a->a1_mode == 0 && (!a->out_nchw || a->N <= 8)),
              "gemm_tc: fused normalisation (a1_coef) takes 3x3 stride-1 convolutions only");
  VDM_REQUIRE(a->a1_raw_dtype == 0 || (xf && a->out_nchw && a->a1_raw_dtype == VDM_F16),
              "gemm_tc: a1_raw_dtype = VDM_F16 belongs to the small-N output head with a1_coef");
  VDM_REQUIRE(a->img_done == nullptr || (a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && !xf),
              "gemm_tc: img_done is maintained by the transposed-role 3x3 conv kernels only");
  if (a->a1_mode == 3) return probe ? 0 : gemm_tc_upfold(a, stream);
  if (!probe && a->out_nchw && a->N <= 8 && !getenv("VDM_NO_SMALL_N")) {
    const int rc = conv3x3_small_n(a, stream);
    if (rc != -100) return rc;
  }
  VDM_REQUIRE(!(xf && a->out_nchw), "gemm_tc: this output-head shape has no fused-normalisation kernel");
  VDM_REQUIRE(a->taps == 1 || a->taps == 9, "gemm_tc: taps must be 1 or 9");
  VDM_REQUIRE(a->C1 > 0 && a->C1 % BLOCK_K == 0, "gemm_tc: C1=%d must be a multiple of 64", a->C1);
  VDM_REQUIRE(a->C2 % BLOCK_K == 0 && a->C2b % BLOCK_K == 0, "gemm_tc: C2=%d / C2b=%d must be multiples of 64", a->C2,
              a->C2b);
  VDM_REQUIRE(a->C2b == 0 || (a->C2 > 0 && a->a2b != nullptr), "gemm_tc: a2b needs a2 (the first tensor of the range)");
  VDM_REQUIRE(a->a2_dtype == VDM_BF16 || a->a2_dtype == VDM_F16 || a->a2_dtype == 0,
              "gemm_tc: a2_dtype must be VDM_BF16 or VDM_F16");
  VDM_REQUIRE(a->a1_mode == 0 || (a->a1_mode == 1 && a->taps == 9), "gemm_tc: unsupported a1_mode %d", a->a1_mode);
  VDM_REQUIRE(a->out_nchw || a->N % 8 == 0, "gemm_tc: N=%d must be a multiple of 8", a->N);
  VDM_REQUIRE(a->out_silu_f32 == nullptr, "gemm_tc: out_silu_f32 is only supported by the fp32 kernel");
  VDM_REQUIRE(a->stats_out == nullptr || (!a->out_nchw && (a->H * a->W) % 32 == 0),
              "gemm_tc: stats_out needs H*W %% 32 == 0 and a channels-last output");
  VDM_REQUIRE(!a->out_nchw || (a->out_f32 && !a->residual && !a->rowbias && !a->out_bf16),
              "gemm_tc: NCHW output supports bias only");
  const int HW = a->H * a->W;
  const bool is_linear = (a->taps == 1);
  TcParams p{};
  p.M = (int)M;
  p.N = a->N;
  p.taps = a->taps;
  p.c1_chunks = a->C1 / BLOCK_K;
  p.c2_chunks = (a->C2 + a->C2b) / BLOCK_K;
  p.c2a_chunks = a->C2 / BLOCK_K;
  p.a2_f16 = a->a2_dtype == VDM_F16 ? 1 : 0;
  p.a1_f16 = a->dtype == VDM_F16 ? 1 : 0;
  VDM_REQUIRE(!p.a1_f16 || (a->taps == 1 && a->C2 == 0 && !xf && a->w_group_tiles == 0 && !a->out_nchw),
              "gemm_tc: dtype VDM_F16 (fp16 activations and weights) takes plain linears only");
  p.a1_mode = a->a1_mode;
  p.is_linear = is_linear;
  p.H = a->H; p.W = a->W; p.HW = HW;
  p.bias = a->bias; p.rowbias = a->rowbias; p.ld_rowbias = a->ld_rowbias;
  p.residual = a->residual; p.ld_res = a->ld_res;
  p.out_f32 = a->out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.ld_out = a->ld_out; p.ld_out_bf16 = a->ld_out_bf16; p.out_nchw = a->out_nchw;
  p.stats_out = a->stats_out;
  p.stats_via_smem = 1;
  p.xf_coef = reinterpret_cast<const float2*>(a->a1_coef);
  p.xf_act = a->a1_act;
  VDM_REQUIRE(a->io_dtype == VDM_F32 || a->io_dtype == VDM_F16, "gemm_tc: io_dtype must be VDM_F32 or VDM_F16");
  p.io_f16 = (a->io_dtype == VDM_F16 && (a->out_f32 != nullptr || a->residual != nullptr)) ? 1 : 0;
  VDM_REQUIRE(!p.io_f16 || (!a->out_nchw && a->out_f32 != nullptr), "gemm_tc: fp16 stream IO needs an fp16 `out_f32` output");
  p.w_group_tiles = a->w_group_tiles;
  p.n_par = 1;
  const int n_prob = a->n_prob > 1 ? a->n_prob : 1;
  if (n_prob > 1) {   // several same-shape linears in one launch (column blocks of one A1, stacked weights, stacked outputs)
    VDM_REQUIRE(is_linear && a->C2 == 0 && !a->out_nchw && !a->residual && !a->rowbias && !a->stats_out && !a->bias,
                "gemm_tc: a problem batch takes plain linears only");
    VDM_REQUIRE(a->prob_a_cols % BLOCK_K == 0 && a->lda1 >= a->C1 + (n_prob - 1) * a->prob_a_cols,
                "gemm_tc: problem batch: bad A1 column stride");
    p.n_par = n_prob;
    p.par_a_cols = a->prob_a_cols;
    p.par_w_rows = (int)a->prob_w_rows;
    p.par_out_stride = a->prob_out_stride;
  }
  p.trace = g_trace_buf;
  p.img_done = a->img_done;
  if (const char* e = getenv("VDM_GEMM_DEBUG")) p.dbg = atoi(e);
  VDM_REQUIRE(a->w_group_tiles == 0 || (is_linear && a->C2 == 0 && !a->out_nchw), "gemm_tc: grouped weights need taps == 1");
  VDM_REQUIRE(a->lda1 == 0 || (is_linear && a->lda1 >= a->C1 && a->lda1 % 8 == 0), "gemm_tc: bad lda1");

  CUtensorMap ma1, mw;
  A2Maps ma2;
  int rc;
  const bool xf_epi = xf_variant_ok(epilogue_variant(p, 128));   // what the fused-normalisation instances exist for
  if (probe) {
    // mirror of the halo dispatch below, without descriptors or launches
    const int hmode = getenv("VDM_GEMM_HALO") ? atoi(getenv("VDM_GEMM_HALO")) : 1;
    const int tmode = getenv("VDM_GEMM_HALO_T") ? atoi(getenv("VDM_GEMM_HALO_T")) : 1;
    const int bn = a->N % 256 == 0 ? 256 : (a->N % 192 == 0 ? 192 : (a->N % 128 == 0 ? 128 : 0));
    const int rows = BLOCK_M * (bn == 128 ? 2 : 1);
    const bool base = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 && xf_epi;
    const bool ok8 = base && a->W == 8 && a->H == 8 && a->N % 256 == 0 &&
                     (hmode == 2 || ((M + 255) / 256) * (a->N / 256) >= 40);
    const bool ok = base && bn != 0 && a->W >= 8 && a->W <= 64 && rows % a->W == 0 && HW % rows == 0 &&
                    (hmode == 2 || ((M + 2 * rows - 1) / (2 * rows)) * (a->N / bn) >= 40);
    const bool okw = base && a->N % 128 == 0 && a->W == 128 && HW % 256 == 0 &&
                     (hmode == 2 || ((M + 511) / 512) * (a->N / 128) >= 40);
    // bit 1: the call runs on a transposed-role kernel (which maintains img_done); mirrors the dispatch below
    const bool lean = (epilogue_variant(p, 128) & 8) == 0;
    const bool base_t = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 && lean;
    const bool geo = bn != 0 && a->W >= 8 && a->W <= 64 && rows % a->W == 0 && HW % rows == 0 &&
                     (hmode == 2 || ((M + 2 * rows - 1) / (2 * rows)) * (a->N / bn) >= 40);
    const bool sel8 = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 && a->W == 8 &&
                      a->H == 8 && a->N % 256 == 0 && (epilogue_variant(p, 256) & 8) == 0 &&
                      (hmode == 2 || ((M + 255) / 256) * (a->N / 256) >= 40);
    const bool sel_wide_t = base_t && a->N % 128 == 0 && a->W == 128 && HW % 256 == 0 &&
                            (hmode == 2 || ((M + 511) / 512) * (a->N / 128) >= 40);
    const bool sel_halo_t = base_t && geo && a->N % 128 == 0 && (a->N % 256 != 0 || tmode == 2) && tmode > 0 &&
                            256 % a->W == 0 && HW % 256 == 0;
    *probe = ((ok8 || ok || okw) ? 1 : 0) | ((!sel8 && !xf && (sel_wide_t || sel_halo_t)) ? 2 : 0);
    return 0;
  }
  if (is_linear) {
    rc = encode_rows_map(&ma1, a->a1, M, a->C1 + (n_prob - 1) * (n_prob > 1 ? a->prob_a_cols : 0), a->lda1);
  } else {
    // tile = 128 consecutive pixels of the channels-last image stack
    const int W = a->W, H = a->H;
    VDM_REQUIRE((W <= 128 && 128 % W == 0) || W % 128 == 0, "gemm_tc: unsupported width %d", W);
    uint32_t bw = W < 128 ? W : 128, bh = 1, bn = 1;
    if (W < 128) {
      if (HW >= 128) {
        VDM_REQUIRE(HW % 128 == 0, "gemm_tc: H*W=%d must be a multiple of 128", HW);
        bh = 128 / W;
      } else {
        VDM_REQUIRE(128 % HW == 0, "gemm_tc: H*W=%d must divide 128", HW);
        bh = H;
        bn = 128 / HW;
      }
    }
    const uint64_t C = a->C1;
    const uint64_t planes = a->a1_mode == 1 ? 4 : 1;
    uint64_t dims[5] = {C, (uint64_t)W, (uint64_t)H, planes, (uint64_t)a->n_img};
    uint64_t st[5] = {2, C * 2, C * 2 * W, C * 2 * W * H, C * 2 * W * H * planes};
    uint32_t box[5] = {BLOCK_K, bw, bh, 1, bn};
    rc = encode_map(&ma1, a->a1, 5, dims, st, box);
  }
  if (rc) return rc;
  ma2.a = ma2.b = ma1;
  if (a->C2 > 0) {
    VDM_REQUIRE(a->a2 != nullptr, "gemm_tc: a2 is NULL");
    rc = encode_rows_map(&ma2.a, a->a2, M, a->C2);
    if (rc) return rc;
    if (a->C2b > 0) {
      rc = encode_rows_map(&ma2.b, a->a2b, M, a->C2b);
      if (rc) return rc;
    }
  }
  const int64_t K = (int64_t)a->taps * a->C1 + a->C2 + a->C2b;
  // 3x3 stride-1 layers whose CTA tile is a whole number of image rows inside one image: halo kernel (the three
  // vertical taps share one activation slot); VDM_GEMM_HALO=0 switches it off
  {
    const char* e = getenv("VDM_GEMM_HALO");   // 0 off, 1 heuristic (default), 2 whenever legal (tests)
    const int hmode = e ? atoi(e) : 1;
    const int bn = a->N % 256 == 0 ? 256 : (a->N % 192 == 0 ? 192 : (a->N % 128 == 0 ? 128 : 0));
    const int msub = bn == 128 ? 2 : 1;
    const int rows = BLOCK_M * msub;
    // 8x8 level: a 128-row tile is two whole images -> interleaved halo tiles (lean epilogues, 256-wide only)
    const bool ok8 = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 &&
                     a->W == 8 && a->H == 8 && a->N % 256 == 0 && (epilogue_variant(p, 256) & 8) == 0 &&
                     (hmode == 2 || ((M + 255) / 256) * (a->N / 256) >= 40);
    if (ok8) {
      VDM_REQUIRE(a->img_done == nullptr, "gemm_tc: img_done is maintained by the transposed-role conv kernels only");
      CUtensorMap mh, mw2;
      const uint64_t C = a->C1;
      // dims (C, x, image, y): the image dimension sits between x and y, so the box lands (y, image, x)-ordered
      uint64_t dims[5] = {C, 8, (uint64_t)a->n_img, 8, 1};
      uint64_t st[5] = {2, C * 2, C * 2 * 64, C * 2 * 8, C * 2 * 64 * (uint64_t)a->n_img};
      uint32_t box[5] = {BLOCK_K, 8, 2, 10, 1};
      rc = encode_map(&mh, a->a1, 5, dims, st, box);
      if (rc) return rc;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      uint32_t wbox[2] = {BLOCK_K, 128};
      rc = encode_map(&mw2, a->w, 2, wdims, wst, wbox);
      if (rc) return rc;
      A2Maps ma2p{mh, mh};
      for (int src = 0; src < 2; ++src) {   // the fused 1x1 skip operand(s) through the same (x, image, y) view
        const uint64_t C2 = src ? a->C2b : a->C2;
        if (C2 == 0) continue;
        uint64_t d2[5] = {C2, 8, (uint64_t)a->n_img, 8, 1};
        uint64_t s2[5] = {2, C2 * 2, C2 * 2 * 64, C2 * 2 * 8, C2 * 2 * 64 * (uint64_t)a->n_img};
        uint32_t b2[5] = {BLOCK_K, 8, 2, 8, 1};
        rc = encode_map(src ? &ma2p.b : &ma2p.a, src ? a->a2b : a->a2, 5, d2, s2, b2);
        if (rc) return rc;
      }
      return launch_halo_ilv<256, 3, 4>(mh, ma2p, mw2, p, stream);
    }
    const bool ok = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 &&
                    bn != 0 && a->W >= 8 && a->W <= 64 && rows % a->W == 0 && HW % rows == 0 &&
                    (hmode == 2 || ((M + 2 * rows - 1) / (2 * rows)) * (a->N / bn) >= 40);
    // 128-pixel-wide images (the top level of the 128x128 model): a 256-pixel tile is two image rows, the slot holds
    // four -> the transposed-role kernel with 64 KB slots (two of them) for every N % 128 == 0
    const bool okw = hmode > 0 && a->taps == 9 && a->a1_mode == 0 && !a->out_nchw && a->w_group_tiles == 0 &&
                     a->N % 128 == 0 && a->W == 128 && HW % 256 == 0 && (epilogue_variant(p, 128) & 8) == 0 &&
                     (hmode == 2 || ((M + 511) / 512) * (a->N / 128) >= 40);
    if (okw) {
      CUtensorMap mh, mwt;
      const uint64_t C = a->C1;
      uint64_t dims[5] = {C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->n_img};
      uint64_t st[5] = {2, C * 2, C * 2 * a->W, C * 2 * a->W * a->H, C * 2 * a->W * a->H};
      uint32_t box[5] = {BLOCK_K, (uint32_t)a->W, 4, 1, 1};
      rc = encode_map(&mh, a->a1, 5, dims, st, box);
      if (rc) return rc;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      uint32_t wbox[2] = {BLOCK_K, 64};
      rc = encode_map(&mwt, a->w, 2, wdims, wst, wbox);
      if (rc) return rc;
      return launch_halo_t<2, 6, true>(mh, ma2, mwt, p, stream);
    }
    // transposed-role kernel: weights as the M operand (128-channel tiles), 256 pixels as N
    const char* et = getenv("VDM_GEMM_HALO_T");
    // measured per level: faster than the pair tiles for 128 and 384 output channels (where those are 128 / 192
    // wide), slightly slower for 256.  VDM_GEMM_HALO_T: 0 off, 1 that rule (default), 2 every N % 128 == 0
    const int tmode = et ? atoi(et) : 1;
    if (ok && a->N % 128 == 0 && (a->N % 256 != 0 || tmode == 2) && tmode > 0 && 256 % a->W == 0 && HW % 256 == 0 &&
        (epilogue_variant(p, 128) & 8) == 0) {
      CUtensorMap mh, mwt;
      const uint64_t C = a->C1;
      uint64_t dims[5] = {C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->n_img};
      uint64_t st[5] = {2, C * 2, C * 2 * a->W, C * 2 * a->W * a->H, C * 2 * a->W * a->H};
      uint32_t box[5] = {BLOCK_K, (uint32_t)a->W, (uint32_t)(256 / a->W + 2), 1, 1};
      rc = encode_map(&mh, a->a1, 5, dims, st, box);
      if (rc) return rc;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      uint32_t wbox[2] = {BLOCK_K, 64};
      rc = encode_map(&mwt, a->w, 2, wdims, wst, wbox);
      if (rc) return rc;
      return launch_halo_t<3, 4>(mh, ma2, mwt, p, stream);
    }
    VDM_REQUIRE(a->img_done == nullptr, "gemm_tc: img_done is maintained by the transposed-role conv kernels only "
                                        "(check vdm_gemm_img_done_supported first)");
    if (ok) {
      CUtensorMap mh, mw2;
      const uint64_t C = a->C1;
      uint64_t dims[5] = {C, (uint64_t)a->W, (uint64_t)a->H, 1, (uint64_t)a->n_img};
      uint64_t st[5] = {2, C * 2, C * 2 * a->W, C * 2 * a->W * a->H, C * 2 * a->W * a->H};
      uint32_t box[5] = {BLOCK_K, (uint32_t)a->W, (uint32_t)(rows / a->W + 2), 1, 1};
      rc = encode_map(&mh, a->a1, 5, dims, st, box);
      if (rc) return rc;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      uint32_t wbox[2] = {BLOCK_K, (uint32_t)(bn / 2)};
      rc = encode_map(&mw2, a->w, 2, wdims, wst, wbox);
      if (rc) return rc;
      if (bn == 256) return launch_halo<256, 1, 3, 5>(mh, ma2, mw2, p, stream);
      if (bn == 192) return launch_halo<192, 1, 3, 6>(mh, ma2, mw2, p, stream);
      return launch_halo<128, 2, 3, 4>(mh, ma2, mw2, p, stream);
    }
  }
  VDM_REQUIRE(!xf, "gemm_tc: fused normalisation (a1_coef) is not available for this shape / epilogue "
                   "(check vdm_gemm_fused_norm_supported first)");
  // Linears on the transposed-role kernel (no 3x3 part: every K block is a plain [256 pixels][64] tile of A1, loaded
  // through the kernel's second-range path): one 128 x 256 MMA per K step, weight tiles multicast across the CTA
  // pair, and the lean epilogue (lane = channel: no staging, per-lane statistics) -- the short-K attention linears
  // (qkv, proj_out) are bound by the staged epilogue of the generic kernel, not by the tensor cores.
  // MEASURED SLOWER than the generic kernels on the model's shapes (qkv 58.6 vs 54.4 us, proj_out 53.6 vs 41.6 us at
  // 16x16: one MMA block per 32 KB slot leaves the three-slot ring latency-bound), so it is opt-in.
  // VDM_GEMM_LINT: 0 off (default), 1 heuristic, 2 whenever legal (tests).
  {
    const char* e = getenv("VDM_GEMM_LINT");
    const int lmode = e ? atoi(e) : 0;
    const int v = epilogue_variant(p, 128);
    const bool stats_ok = a->stats_out == nullptr || HW % 128 == 0 || HW == 64;
    const bool legal = lmode > 0 && is_linear && a->C2 == 0 && n_prob == 1 && a->w_group_tiles == 0 && !a->out_nchw &&
                       a->N % 128 == 0 && ((v & ~7) == 0 || v == 64 || v == 65 || v == 68 || v == 69) && stats_ok && !xf;
    if (legal && (lmode == 2 || (((M + 511) / 512) * (a->N / 128) >= 40 && K <= 1024))) {
      CUtensorMap mwt;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      uint32_t wbox[2] = {BLOCK_K, 64};
      rc = encode_map(&mwt, a->w, 2, wdims, wst, wbox);
      if (rc) return rc;
      TcParams pl = p;
      pl.c2_chunks = pl.c2a_chunks = p.c1_chunks;     // all of K arrives as "second range" tiles of A1
      pl.c1_chunks = 0;
      pl.a2_f16 = p.a1_f16;                           // (an fp16 linear: those tiles are IEEE half)
      return launch_halo_t<3, 4>(ma1, ma2, mwt, pl, stream);
    }
  }
  // short-K linears with several column tiles (attention qkv / proj_out): A-stationary pair kernel
  {
    const char* e = getenv("VDM_GEMM_ASTAT");
    const bool on = e == nullptr || atoi(e) != 0;
    const bool base = on && is_linear && a->C2 == 0 && n_prob == 1 && a->w_group_tiles == 0 && !a->out_nchw &&
                      !a->rowbias && p.c1_chunks <= 8 && a->N % 128 == 0 && (M + 255) / 256 >= 37;
    if (base) {
      const int v = epilogue_variant(p, 128);
      const char* ts_env = getenv("VDM_GEMM_TS");
      const bool ts_on = ts_env == nullptr || atoi(ts_env) != 0;
      CUtensorMap mo = ma1;
      uint64_t wdims[2] = {(uint64_t)K, (uint64_t)a->N};
      uint64_t wst[2] = {2, (uint64_t)K * 2};
      if (v == 2 && ts_on && a->N / 128 >= 2) {                     // bf16 output, bias only: TMA-store epilogue
        const bool wide = a->N % 192 == 0;
        if (wide ? encode_out_map<192, 2>(&mo, p) == 0 : encode_out_map<128, 2>(&mo, p) == 0) {
          uint32_t wbox[2] = {BLOCK_K, (uint32_t)(wide ? 96 : 64)};
          rc = encode_map(&mw, a->w, 2, wdims, wst, wbox);
          if (rc) return rc;
          if (wide) return launch_astat<192, 8, 4, 34>(ma1, mw, mo, p, stream);
          return launch_astat<128, 8, 6, 34>(ma1, mw, mo, p, stream);
        }
      } else if ((v == 4 || v == 5 || v == 68 || v == 69) && a->N / 128 >= 2) {   // stream output (+ residual) + statistics
        uint32_t wbox[2] = {BLOCK_K, 64};
        rc = encode_map(&mw, a->w, 2, wdims, wst, wbox);
        if (rc) return rc;
        switch (v) {
          case 4: return launch_astat<128, 8, 6, 4>(ma1, mw, mo, p, stream);
          case 5: return launch_astat<128, 8, 6, 5>(ma1, mw, mo, p, stream);
          case 68: return launch_astat<128, 8, 6, 68>(ma1, mw, mo, p, stream);
          default: return launch_astat<128, 8, 6, 69>(ma1, mw, mo, p, stream);
        }
      }
    }
  }
  int block_n = a->out_nchw || a->N <= 16 ? 16 : (a->N % 128 == 0 ? 128 : 64);
  // 256x128 CTA tiles (two 128-row sub-tiles sharing every weight tile) halve the L2->smem operand
  // traffic per FLOP; use them unless the layer is too small to fill the SMs that way.
  int m_sub = 1;
  if (block_n == 128) {
    const int64_t t1 = ((M + 127) / 128) * (a->N / 128), t2 = ((M + 255) / 256) * (a->N / 128);
    const int sms = num_sms();
    const double cost1 = (double)((t1 + sms - 1) / sms) * 1.35, cost2 = (double)((t2 + sms - 1) / sms) * 2.0;
    if (cost2 <= cost1 && K >= 1024 && a->w_group_tiles == 0) m_sub = 2;   // short-K GEMMs are epilogue-bound: keep the 32-column epilogue
  }
  if (const char* e = getenv("VDM_GEMM_MSUB")) m_sub = atoi(e) == 2 && block_n == 128 && a->w_group_tiles == 0 ? 2 : 1;
  VDM_REQUIRE(a->w_group_tiles == 0 || a->N % block_n == 0, "gemm_tc: grouped weights need N %% %d == 0", block_n);
  // 2-CTA pairs (cta_group::2, 256 x N tiles): each SM feeds half of the weight tile from its own shared
  // memory, which is what limits the single-CTA kernel on long-K layers
  bool cta2 = false;
  {
    int mode = 1;   // VDM_GEMM_CTA2: 0 = never, 1 = heuristic (default), 2 = whenever legal (tests), 3 = heuristic + 512x128 pair tiles
    if (const char* e = getenv("VDM_GEMM_CTA2")) mode = atoi(e);
    const bool legal = !a->out_nchw && a->N % 128 == 0 && a->w_group_tiles == 0 && a->a1_mode <= 1 && n_prob == 1;
    const int bn2 = a->N % 256 == 0 ? 256 : (a->N % 192 == 0 ? 192 : 128);
    const int64_t pair_tiles = ((M + 255) / 256) * (a->N / bn2);
    // Two measured ceilings (profiles/gemm_trace.py): TMA operand delivery from L2 saturates near 10 TB/s over
    // the chip, and an N=128 MMA (both operands in shared memory) retires in ~112 cycles instead of 64, an N=256
    // one in ~165 instead of 128.  So: 256- and 192-wide pair tiles beat the single-CTA kernel by ~20 % on
    // long-K layers; 512x128 pair tiles only tie with the single-CTA 256x128 tile (mode 3 selects them); wide
    // short-K linears (qkv, N >= 1024) gain 15-18 % from pair tiles.
    cta2 = legal && mode > 0 &&
           (mode == 2 || (bn2 >= 192 && pair_tiles >= 40 && (K >= 1024 || a->N >= 1024)) ||
            (bn2 == 128 && m_sub == 2 && mode == 3));
    if (cta2) block_n = bn2;
  }
  {
    const int64_t groups = a->w_group_tiles ? ((M + BLOCK_M - 1) / BLOCK_M + a->w_group_tiles - 1) / a->w_group_tiles : 1;
    uint64_t dims[2] = {(uint64_t)K, n_prob > 1 ? (uint64_t)a->prob_w_rows * n_prob : (uint64_t)a->N * groups};
    uint64_t st[2] = {2, (uint64_t)K * 2};
    uint32_t box[2] = {BLOCK_K, (uint32_t)(cta2 ? block_n / 2 : block_n)};
    rc = encode_map(&mw, a->w, 2, dims, st, box);
    if (rc) return rc;
  }
  if (cta2) {
    if (block_n == 256) return launch<256, 1, 5, true>(ma1, ma2, mw, p, stream);
    if (block_n == 192) return launch<192, 1, 6, true>(ma1, ma2, mw, p, stream);
    if (m_sub == 2) return launch<128, 2, 4, true>(ma1, ma2, mw, p, stream);
    return launch<128, 1, 7, true>(ma1, ma2, mw, p, stream);
  }
  if (n_prob > 1)
    p.tiles_per_par = ((a->N + block_n - 1) / block_n) * (int)((M + BLOCK_M * m_sub - 1) / (BLOCK_M * m_sub));
  switch (block_n) {
    case 128:
      if (m_sub == 2) return launch<128, 2, 4>(ma1, ma2, mw, p, stream);
      return launch<128, 1, 5>(ma1, ma2, mw, p, stream);
    case 64: return launch<64, 1, 6>(ma1, ma2, mw, p, stream);
    default: return launch<16, 1, 8>(ma1, ma2, mw, p, stream);
  }
}

// nearest-x2 upsample + 3x3 conv (unet.py:63-72) as four 2x2 convolutions, one per output parity, on
// the LOW-resolution input: every output pixel (2y+a, 2x+b) only ever sees two distinct input rows and
// columns, so the 3x3 weights fold into 2x2 ones (summed on the host at pack time).  4/9 of the FLOPs,
// and the 4x larger upsampled tensor is never materialised.  w: [4 parities][N][4*C1].
int gemm_tc_upfold(const vdm_gemm_args* a, cudaStream_t stream) {
  VDM_REQUIRE(a->taps == 4, "gemm_tc: a1_mode 3 takes the 4-tap folded weights");
  VDM_REQUIRE(a->H % 2 == 0 && a->W % 2 == 0, "gemm_tc: upsample output must be even");
  VDM_REQUIRE(a->C1 > 0 && a->C1 % BLOCK_K == 0 && a->C2 == 0, "gemm_tc: C1=%d must be a multiple of 64", a->C1);
  VDM_REQUIRE(a->N % 128 == 0, "gemm_tc: folded upsample needs N %% 128 == 0");
  VDM_REQUIRE(!a->out_nchw && !a->residual && !a->rowbias && !a->out_silu_f32, "gemm_tc: unsupported epilogue for a1_mode 3");
  const int Hl = a->H / 2, Wl = a->W / 2, HWl = Hl * Wl;
  VDM_REQUIRE(a->stats_out == nullptr || HWl % 32 == 0, "gemm_tc: stats_out needs (H/2)*(W/2) %% 32 == 0");
  const int64_t M = (int64_t)a->n_img * HWl;
  TcParams p{};
  p.M = (int)M; p.N = a->N; p.taps = 4; p.c1_chunks = a->C1 / BLOCK_K; p.c2_chunks = 0;
  p.a1_mode = 3; p.is_linear = 0; p.H = Hl; p.W = Wl; p.HW = HWl;
  p.bias = a->bias; p.out_f32 = a->out_f32; p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out_bf16);
  p.ld_out = a->ld_out; p.ld_out_bf16 = a->ld_out_bf16; p.stats_out = a->stats_out;
  p.io_f16 = (a->io_dtype == VDM_F16 && a->out_f32 != nullptr) ? 1 : 0;
  p.stats_via_smem = 1;
  p.n_par = 4;
  p.par_w_rows = a->N;
  p.trace = g_trace_buf;
  if (const char* e = getenv("VDM_GEMM_DEBUG")) p.dbg = atoi(e);
  VDM_REQUIRE((Wl <= 128 && 128 % Wl == 0) || Wl % 128 == 0, "gemm_tc: unsupported width %d", Wl);
  uint32_t bw = Wl < 128 ? Wl : 128, bh = 1, bn = 1;
  if (Wl < 128) {
    if (HWl >= 128) {
      VDM_REQUIRE(HWl % 128 == 0, "gemm_tc: H*W=%d must be a multiple of 128", HWl);
      bh = 128 / Wl;
    } else {
      VDM_REQUIRE(128 % HWl == 0, "gemm_tc: H*W=%d must divide 128", HWl);
      bh = Hl;
      bn = 128 / HWl;
    }
  }
  CUtensorMap ma1, mw;
  const uint64_t C = a->C1;
  uint64_t dims[5] = {C, (uint64_t)Wl, (uint64_t)Hl, 1, (uint64_t)a->n_img};
  uint64_t st[5] = {2, C * 2, C * 2 * Wl, C * 2 * Wl * Hl, C * 2 * Wl * Hl};
  uint32_t box[5] = {BLOCK_K, bw, bh, 1, bn};
  int rc = encode_map(&ma1, a->a1, 5, dims, st, box);
  if (rc) return rc;
  const int64_t K = 4 * (int64_t)a->C1;
  const A2Maps no_a2{ma1, ma1};
  // pair tiles (256 low-res pixels x 256 / 192 channels per parity) where the layer is wide enough
  int mode = 1;
  if (const char* e = getenv("VDM_GEMM_CTA2")) mode = atoi(e);
  {
    // halo variant: both vertical parities of 256 low-res pixels x 128 channels per CTA pair share every activation box
    const char* e = getenv("VDM_GEMM_HALO");
    const int hmode = e ? atoi(e) : 1;
    const int64_t tiles_b = ((M + 255) / 256) * (a->N / 128);
    if (mode > 0 && hmode > 0 && a->N % 128 == 0 && Wl >= 16 && Wl <= 64 && 128 % Wl == 0 && HWl % 128 == 0 &&
        (hmode == 2 || 2 * tiles_b >= 40)) {
      CUtensorMap mh, mwh;
      uint32_t hbox[5] = {BLOCK_K, (uint32_t)Wl, (uint32_t)(128 / Wl + 2), 1, 1};
      rc = encode_map(&mh, a->a1, 5, dims, st, hbox);
      if (rc) return rc;
      uint64_t wd2[2] = {(uint64_t)K, (uint64_t)a->N * 4};
      uint64_t ws2[2] = {2, (uint64_t)K * 2};
      uint32_t wb2[2] = {BLOCK_K, 64};
      rc = encode_map(&mwh, a->w, 2, wd2, ws2, wb2);
      if (rc) return rc;
      p.n_par = 2;
      p.tiles_per_par = (int)tiles_b;
      return launch_upfold_halo<3, 8>(mh, mwh, p, stream);
    }
  }
  const int bn2 = a->N % 256 == 0 ? 256 : (a->N % 192 == 0 ? 192 : 0);
  const int64_t pair_tiles = bn2 ? ((M + 255) / 256) * (a->N / bn2) : 0;
  const bool cta2 = mode > 0 && bn2 && 4 * pair_tiles >= 40;
  uint64_t wd[2] = {(uint64_t)K, (uint64_t)a->N * 4};
  uint64_t ws[2] = {2, (uint64_t)K * 2};
  uint32_t wb[2] = {BLOCK_K, (uint32_t)(cta2 ? bn2 / 2 : 128)};
  rc = encode_map(&mw, a->w, 2, wd, ws, wb);
  if (rc) return rc;
  if (cta2) {
    p.tiles_per_par = (int)pair_tiles;
    if (bn2 == 256) return launch_inst<256, 1, 5, 8, true>(ma1, no_a2, mw, p, stream);
    return launch_inst<192, 1, 6, 8, true>(ma1, no_a2, mw, p, stream);
  }
  const int64_t t2 = ((M + 255) / 256) * (a->N / 128), t1 = ((M + 127) / 128) * (a->N / 128);
  const int sms = num_sms();
  const bool two = (double)((4 * t2 + sms - 1) / sms) * 2.0 <= (double)((4 * t1 + sms - 1) / sms) * 1.35;
  if (two) {
    p.tiles_per_par = (int)t2;
    return launch_inst<128, 2, 4, 8>(ma1, no_a2, mw, p, stream);
  }
  p.tiles_per_par = (int)t1;
  return launch_inst<128, 1, 5, 8>(ma1, no_a2, mw, p, stream);
}

}  // namespace vdm
