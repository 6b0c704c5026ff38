// Sampler step, q_sample and ELBO terms: HBM-bound elementwise kernels (one vectorised pass
// each, schedule coefficients gathered in-kernel from device tables by the per-row timestep).
// Arithmetic follows the reference's fp32 op order with explicit _rn intrinsics (no FMA
// contraction) so results track torch's elementwise ops to the last bit or two.
#include "common.cuh"

namespace vdm {
namespace {

struct Coef {
  float recip, recipm1, c1, c2, logvar, acp, acp_prev, post_logvar, sqrt_acp, sqrt_1m_acp, log_1m_acp, acp_next;
};

// The reference indexes numpy tables with t and raises IndexError outside [0, n_steps).  Here an out-of-range
// timestep never reads past the table: it is clamped and recorded in a host-mapped flag that the next
// vdm_sampler_error() call reports (the Python layer turns it into the reference's IndexError).
__device__ int* g_t_error = nullptr;
__device__ __forceinline__ long long check_t(long long t, int n_steps) {
  if (t < 0 || t >= n_steps) {
    if (g_t_error != nullptr) *reinterpret_cast<volatile int*>(g_t_error) = 1;
    t = t < 0 ? 0 : n_steps - 1;
  }
  return t;
}

__device__ __forceinline__ Coef load_coef(const float* tab, int n_steps, long long t) {
  Coef c;
  c.recip = tab[VDM_TAB_SQRT_RECIP_ACP * n_steps + t];
  c.recipm1 = tab[VDM_TAB_SQRT_RECIPM1_ACP * n_steps + t];
  c.c1 = tab[VDM_TAB_POST_C1 * n_steps + t];
  c.c2 = tab[VDM_TAB_POST_C2 * n_steps + t];
  c.logvar = tab[VDM_TAB_MODEL_LOGVAR * n_steps + t];
  c.acp = tab[VDM_TAB_ACP * n_steps + t];
  c.acp_prev = tab[VDM_TAB_ACP_PREV * n_steps + t];
  c.post_logvar = tab[VDM_TAB_POST_LOGVAR * n_steps + t];
  c.sqrt_acp = tab[VDM_TAB_SQRT_ACP * n_steps + t];
  c.sqrt_1m_acp = tab[VDM_TAB_SQRT_1M_ACP * n_steps + t];
  c.log_1m_acp = tab[VDM_TAB_LOG_1M_ACP * n_steps + t];
  c.acp_next = tab[VDM_TAB_ACP_NEXT * n_steps + t];
  return c;
}

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ float pred_xstart_of(const Coef& c, float x, float eps, int clip) {
  float p = sub(mul(c.recip, x), mul(c.recipm1, eps));
  if (clip) p = fminf(fmaxf(p, -1.0f), 1.0f);
  return p;
}

struct StepConsts {  // per batch row, computed once per thread
  Coef c;
  float noise_scale;   // ancestral: 1[t!=0]*exp(0.5*logvar); ddim: 1[t!=0]*sigma
  float sqrt_abp, dir;  // ddim: sqrt(acp_prev), sqrt(1-acp_prev-sigma^2)
};

__device__ __forceinline__ float step_value(int mode, const StepConsts& s, float x, float eps, float z, int clip,
                                            float* pred_out, float* mean_out) {
  const float pred = pred_xstart_of(s.c, x, eps, clip);
  float mean;
  if (mode == 0) {
    mean = add(mul(s.c.c1, pred), mul(s.c.c2, x));
  } else if (mode == 2) {   // DDIM reverse ODE (gaussian_diffusion.py:651-666): x_{t+1} from x_t
    const float e = __fdiv_rn(sub(mul(s.c.recip, x), pred), s.c.recipm1);
    mean = add(mul(pred, s.sqrt_abp), mul(s.dir, e));
  } else {
    const float e = __fdiv_rn(sub(mul(s.c.recip, x), pred), s.c.recipm1);
    mean = add(mul(pred, s.sqrt_abp), mul(s.dir, e));
  }
  *pred_out = pred;
  *mean_out = mean;
  return add(mean, mul(s.noise_scale, z));
}

template <int VEC>
__global__ void __launch_bounds__(256) sampler_step_kernel(int mode, const float* __restrict__ x,
                                                            const float* __restrict__ eps,
                                                            const float* __restrict__ noise,
                                                            const long long* __restrict__ t,
                                                            const float* __restrict__ tab, int n_steps,
                                                            long long per_batch, int clip, float eta,
                                                            float* __restrict__ sample, float* __restrict__ pred_xstart,
                                                            float* __restrict__ mean_out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int b = blockIdx.y;
  const long long tb = check_t(t[b], n_steps);
  StepConsts s;
  s.c = load_coef(tab, n_steps, tb);
  const float nz = tb != 0 ? 1.0f : 0.0f;
  if (mode == 0) {
    s.noise_scale = mul(nz, expf(mul(0.5f, s.c.logvar)));
    s.sqrt_abp = 0.f;
    s.dir = 0.f;
  } else if (mode == 2) {
    s.noise_scale = 0.f;
    s.sqrt_abp = sqrtf(s.c.acp_next);
    s.dir = sqrtf(sub(1.0f, s.c.acp_next));
  } else {
    const float ab = s.c.acp, abp = s.c.acp_prev;
    const float sigma = mul(mul(eta, sqrtf(__fdiv_rn(sub(1.0f, abp), sub(1.0f, ab)))),
                            sqrtf(sub(1.0f, __fdiv_rn(ab, abp))));
    s.noise_scale = mul(nz, sigma);
    s.sqrt_abp = sqrtf(abp);
    s.dir = sqrtf(sub(sub(1.0f, abp), mul(sigma, sigma)));
  }
  const long long base = (long long)b * per_batch;
  const long long nvec = per_batch / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long off = base + i * VEC;
    if constexpr (VEC == 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + off));
      const float4 ev = __ldg(reinterpret_cast<const float4*>(eps + off));
      const float4 zv = __ldg(reinterpret_cast<const float4*>(noise + off));
      float4 o, p, m;
      o.x = step_value(mode, s, xv.x, ev.x, zv.x, clip, &p.x, &m.x);
      o.y = step_value(mode, s, xv.y, ev.y, zv.y, clip, &p.y, &m.y);
      o.z = step_value(mode, s, xv.z, ev.z, zv.z, clip, &p.z, &m.z);
      o.w = step_value(mode, s, xv.w, ev.w, zv.w, clip, &p.w, &m.w);
      *reinterpret_cast<float4*>(sample + off) = o;
      if (pred_xstart) *reinterpret_cast<float4*>(pred_xstart + off) = p;
      if (mean_out) *reinterpret_cast<float4*>(mean_out + off) = m;
    } else {
      float p, m;
      const float o = step_value(mode, s, x[off], eps[off], noise[off], clip, &p, &m);
      sample[off] = o;
      if (pred_xstart) pred_xstart[off] = p;
      if (mean_out) mean_out[off] = m;
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                        const long long* __restrict__ t, const float* __restrict__ tab,
                                                        int n_steps, long long per_batch, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int b = blockIdx.y;
  const long long tb = check_t(t[b], n_steps);
  const float a = tab[VDM_TAB_SQRT_ACP * n_steps + tb], s = tab[VDM_TAB_SQRT_1M_ACP * n_steps + tb];
  const long long base = (long long)b * per_batch;
  const long long nvec = per_batch / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long off = base + i * VEC;
    if constexpr (VEC == 4) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x0 + off));
      const float4 zv = __ldg(reinterpret_cast<const float4*>(noise + off));
      float4 o;
      o.x = add(mul(a, xv.x), mul(s, zv.x));
      o.y = add(mul(a, xv.y), mul(s, zv.y));
      o.z = add(mul(a, xv.z), mul(s, zv.z));
      o.w = add(mul(a, xv.w), mul(s, zv.w));
      *reinterpret_cast<float4*>(out + off) = o;
    } else {
      out[off] = add(mul(a, x0[off]), mul(s, noise[off]));
    }
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) lincomb_kernel(int op, const float* __restrict__ a, const float* __restrict__ b,
                                                       const long long* __restrict__ t, const float* __restrict__ tab,
                                                       int n_steps, int row_a, int row_b, long long per_batch,
                                                       float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  const int bi = blockIdx.y;
  const long long tb = check_t(t[bi], n_steps);
  const float ca = tab[row_a * n_steps + tb], cb = op == 3 ? 0.f : tab[row_b * n_steps + tb];
  auto f = [&](float av, float bv) {
    if (op == 0) return add(mul(ca, av), mul(cb, bv));
    if (op == 1) return sub(mul(ca, av), mul(cb, bv));
    if (op == 2) return __fdiv_rn(sub(mul(ca, av), bv), cb);
    return mul(ca, av);
  };
  const long long base = (long long)bi * per_batch;
  const long long nvec = per_batch / VEC;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const long long off = base + i * VEC;
    if constexpr (VEC == 4) {
      const float4 av = __ldg(reinterpret_cast<const float4*>(a + off));
      const float4 bv = op == 3 ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldg(reinterpret_cast<const float4*>(b + off));
      *reinterpret_cast<float4*>(out + off) = make_float4(f(av.x, bv.x), f(av.y, bv.y), f(av.z, bv.z), f(av.w, bv.w));
    } else {
      out[off] = f(a[off], op == 3 ? 0.f : b[off]);
    }
  }
}

__device__ __forceinline__ float approx_cdf(float v) {
  // 0.5*(1+tanh(sqrt(2/pi)*(v+0.044715 v^3)))  (losses.py:34-38)
  const float v3 = mul(mul(v, v), v);
  return mul(0.5f, add(1.0f, tanhf(mul(0.7978845608028654f, add(v, mul(0.044715f, v3))))));
}

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (w == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  return v;  // valid in thread 0
}

// grid: (chunks, F, B); each block covers a slice of one frame so the latent mask is uniform.
__global__ void __launch_bounds__(256) vb_terms_kernel(const float* __restrict__ x0, const float* __restrict__ x_t,
                                                        const float* __restrict__ eps, const float* __restrict__ noise,
                                                        const long long* __restrict__ t, const float* __restrict__ tab,
                                                        int n_steps, const float* __restrict__ latent_mask, int F,
                                                        long long per_frame, int clip, double* __restrict__ acc) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  __shared__ float red[32];
  const int b = blockIdx.z, f = blockIdx.y;
  const float mask = latent_mask[b * F + f];
  if (mask == 0.0f) return;  // masked terms contribute exactly 0 (x*0 in the reference)
  const long long tb = check_t(t[b], n_steps);
  const Coef c = load_coef(tab, n_steps, tb);
  const float lv = c.logvar, tlv = c.post_logvar;
  const float log_scale = mul(0.5f, lv);
  const float inv_std = expf(-log_scale);
  const float e_tlv_lv = expf(sub(tlv, lv)), e_neg_lv = expf(-lv);
  const long long base = ((long long)b * F + f) * per_frame;
  float s_vb = 0.f, s_x = 0.f, s_e = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_frame;
       i += (long long)gridDim.x * blockDim.x) {
    const float x0v = x0[base + i], xt = x_t[base + i], ev = eps[base + i], zv = noise[base + i];
    const float pred = pred_xstart_of(c, xt, ev, clip);
    const float mean = add(mul(c.c1, pred), mul(c.c2, xt));
    float term;
    if (tb == 0) {
      const float cx = sub(x0v, mean);
      const float cdf_p = approx_cdf(mul(inv_std, add(cx, 1.0f / 255.0f)));
      const float cdf_m = approx_cdf(mul(inv_std, sub(cx, 1.0f / 255.0f)));
      float lp;
      if (x0v < -0.999f) lp = logf(fmaxf(cdf_p, 1e-12f));
      else if (x0v > 0.999f) lp = logf(fmaxf(sub(1.0f, cdf_m), 1e-12f));
      else lp = logf(fmaxf(sub(cdf_p, cdf_m), 1e-12f));
      term = -lp;
    } else {
      const float tm = add(mul(c.c1, x0v), mul(c.c2, xt));
      const float d = sub(tm, mean);
      term = mul(0.5f, add(add(add(sub(add(-1.0f, lv), tlv), e_tlv_lv), 0.f), mul(mul(d, d), e_neg_lv)));
    }
    s_vb += term;
    const float dx = sub(pred, x0v);
    s_x += mul(dx, dx);
    const float e2 = __fdiv_rn(sub(mul(c.recip, xt), pred), c.recipm1);
    const float de = sub(e2, zv);
    s_e += mul(de, de);
  }
  const double inv_cnt = 1.0 / ((double)per_frame * F);
  float r = block_sum(s_vb, red);
  if (threadIdx.x == 0) atomicAdd(&acc[b * 3 + 0], (double)r * (double)mask * inv_cnt / 0.6931471805599453);
  r = block_sum(s_x, red);
  if (threadIdx.x == 0) atomicAdd(&acc[b * 3 + 1], (double)r * (double)mask * inv_cnt);
  r = block_sum(s_e, red);
  if (threadIdx.x == 0) atomicAdd(&acc[b * 3 + 2], (double)r * (double)mask * inv_cnt);
}

__global__ void __launch_bounds__(256) prior_bpd_kernel(const float* __restrict__ x0, const float* __restrict__ tab,
                                                         int n_steps, const float* __restrict__ latent_mask, int F,
                                                         long long per_frame, double* __restrict__ acc) {
  pdl_launch_dependents();
  pdl_wait();   // PDL protocol (common.cuh): nothing global is touched before this line
  __shared__ float red[32];
  const int b = blockIdx.z, f = blockIdx.y;
  const float mask = latent_mask[b * F + f];
  if (mask == 0.0f) return;
  const float sa = tab[VDM_TAB_SQRT_ACP * n_steps + (n_steps - 1)];
  const float lv = tab[VDM_TAB_LOG_1M_ACP * n_steps + (n_steps - 1)];
  const float e_lv = expf(lv);
  const long long base = ((long long)b * F + f) * per_frame;
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per_frame;
       i += (long long)gridDim.x * blockDim.x) {
    const float m = mul(sa, x0[base + i]);
    // normal_kl(mean, lv, 0, 0) = 0.5*(-1 + 0 - lv + exp(lv) + mean^2)
    s += mul(0.5f, add(add(sub(add(-1.0f, 0.0f), lv), e_lv), mul(m, m)));
  }
  const float r = block_sum(s, red);
  if (threadIdx.x == 0)
    atomicAdd(&acc[b], (double)r * (double)mask / ((double)per_frame * F) / 0.6931471805599453);
}

// host-mapped error flag, created on first use and published to the kernels through g_t_error
int* g_t_error_host = nullptr;
int ensure_error_flag() {
  if (g_t_error_host) return 0;
  int* h = nullptr;
  cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&h), sizeof(int), cudaHostAllocMapped);
  if (e != cudaSuccess) {
    set_error("sampler: cudaHostAlloc failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  *h = 0;
  int* d = nullptr;
  e = cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0);
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_t_error, &d, sizeof(d));
  if (e != cudaSuccess) {
    set_error("sampler: error flag setup failed: %s", cudaGetErrorString(e));
    return (int)e;
  }
  g_t_error_host = h;
  return 0;
}

int grid_x_for(long long work_items, int other) {
  long long want = (work_items + 255) / 256;
  long long cap = (long long)num_sms() * 16 / (other > 0 ? other : 1);
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace
}  // namespace vdm

using namespace vdm;

extern "C" int vdm_sampler_step(int32_t mode, const float* x, const float* eps, const float* noise, const int64_t* t,
                                const float* tables, int32_t n_steps, int32_t B, int64_t per_batch,
                                int32_t clip_denoised, float eta, float* sample, float* pred_xstart, float* mean,
                                vdm_stream_t stream) {
  VDM_REQUIRE(mode >= 0 && mode <= 2, "sampler_step: mode must be 0 (ancestral), 1 (ddim) or 2 (ddim reverse)");
  VDM_REQUIRE(mode != 2 || eta == 0.0f, "sampler_step: the reverse ODE is deterministic (eta must be 0)");
  VDM_REQUIRE(x && eps && noise && t && tables && sample, "sampler_step: NULL pointer");
  if (int rc = ensure_error_flag()) return rc;
  VDM_REQUIRE(B > 0 && per_batch > 0 && n_steps > 0, "sampler_step: bad sizes");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = per_batch % 4 == 0 && al(x) && al(eps) && al(noise) && al(sample) && (!pred_xstart || al(pred_xstart)) &&
                   (!mean || al(mean));
  dim3 grid(grid_x_for(vec ? per_batch / 4 : per_batch, B), B);
  if (vec)
    launch_kernel(sampler_step_kernel<4>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, mode, x, eps, noise, (const long long*)t, tables,
                                                                  n_steps, per_batch, clip_denoised, eta, sample,
                                                                  pred_xstart, mean);
  else
    launch_kernel(sampler_step_kernel<1>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, mode, x, eps, noise, (const long long*)t, tables,
                                                                  n_steps, per_batch, clip_denoised, eta, sample,
                                                                  pred_xstart, mean);
  VDM_AFTER_LAUNCH("sampler_step");
  return 0;
}

extern "C" int vdm_q_sample(const float* x0, const float* noise, const int64_t* t, const float* tables, int32_t n_steps,
                            int32_t B, int64_t per_batch, float* out, vdm_stream_t stream) {
  VDM_REQUIRE(x0 && noise && t && tables && out, "q_sample: NULL pointer");
  if (int rc = ensure_error_flag()) return rc;
  VDM_REQUIRE(B > 0 && per_batch > 0, "q_sample: bad sizes");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = per_batch % 4 == 0 && al(x0) && al(noise) && al(out);
  dim3 grid(grid_x_for(vec ? per_batch / 4 : per_batch, B), B);
  if (vec)
    launch_kernel(q_sample_kernel<4>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, x0, noise, (const long long*)t, tables, n_steps, per_batch, out);
  else
    launch_kernel(q_sample_kernel<1>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, x0, noise, (const long long*)t, tables, n_steps, per_batch, out);
  VDM_AFTER_LAUNCH("q_sample");
  return 0;
}

/* 1 if any sampler-family launch that has completed since the last call saw a timestep outside [0, n_steps);
 * clears the flag.  Asynchronous like every CUDA error: synchronise first to be sure a given launch is covered. */
extern "C" int vdm_sampler_error(void) {
  if (!g_t_error_host) return 0;
  const int v = *reinterpret_cast<volatile int*>(g_t_error_host);
  if (v) *reinterpret_cast<volatile int*>(g_t_error_host) = 0;
  return v;
}

extern "C" int vdm_lincomb(int32_t op, const float* a, const float* b, const int64_t* t, const float* tables,
                           int32_t n_steps, int32_t row_a, int32_t row_b, int32_t B, int64_t per_batch, float* out,
                           vdm_stream_t stream) {
  VDM_REQUIRE(op >= 0 && op <= 3, "lincomb: unknown op %d", op);
  VDM_REQUIRE(a && t && tables && out && (b || op == 3), "lincomb: NULL pointer");
  if (int rc = ensure_error_flag()) return rc;
  VDM_REQUIRE(row_a >= 0 && row_a < VDM_TAB_COUNT && (op == 3 || (row_b >= 0 && row_b < VDM_TAB_COUNT)),
              "lincomb: table row out of range");
  VDM_REQUIRE(B > 0 && per_batch > 0 && n_steps > 0, "lincomb: bad sizes");
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = per_batch % 4 == 0 && al(a) && (!b || al(b)) && al(out);
  dim3 grid(grid_x_for(vec ? per_batch / 4 : per_batch, B), B);
  if (vec)
    launch_kernel(lincomb_kernel<4>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, op, a, b, (const long long*)t, tables, n_steps, row_a, row_b,
                                                             per_batch, out);
  else
    launch_kernel(lincomb_kernel<1>, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, op, a, b, (const long long*)t, tables, n_steps, row_a, row_b,
                                                             per_batch, out);
  VDM_AFTER_LAUNCH("lincomb");
  return 0;
}

extern "C" int vdm_vb_terms(const float* x0, const float* x_t, const float* eps, const float* noise, const int64_t* t,
                            const float* tables, int32_t n_steps, const float* latent_mask, int32_t B, int32_t F,
                            int64_t per_frame, int32_t clip_denoised, double* acc, vdm_stream_t stream) {
  VDM_REQUIRE(x0 && x_t && eps && noise && t && tables && latent_mask && acc, "vb_terms: NULL pointer");
  if (int rc = ensure_error_flag()) return rc;
  VDM_REQUIRE(B > 0 && F > 0 && per_frame > 0, "vb_terms: bad sizes");
  dim3 grid(grid_x_for(per_frame, B * F), F, B);
  launch_kernel(vb_terms_kernel, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, x0, x_t, eps, noise, (const long long*)t, tables, n_steps,
                                                         latent_mask, F, per_frame, clip_denoised, acc);
  VDM_AFTER_LAUNCH("vb_terms");
  return 0;
}

extern "C" int vdm_prior_bpd(const float* x0, const float* tables, int32_t n_steps, const float* latent_mask, int32_t B,
                             int32_t F, int64_t per_frame, double* acc, vdm_stream_t stream) {
  VDM_REQUIRE(x0 && tables && latent_mask && acc, "prior_bpd: NULL pointer");
  dim3 grid(grid_x_for(per_frame, B * F), F, B);
  launch_kernel(prior_bpd_kernel, grid, 256, 0, (cudaStream_t)(cudaStream_t)stream, 1, x0, tables, n_steps, latent_mask, F, per_frame, acc);
  VDM_AFTER_LAUNCH("prior_bpd");
  return 0;
}
