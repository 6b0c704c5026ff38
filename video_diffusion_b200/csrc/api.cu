// extern "C" surface shared by all kernels: version, error string, launch counter, GEMM dispatch.
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace vdm {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gemm_tc(const vdm_gemm_args* a, cudaStream_t stream, int* probe = nullptr);
void gemm_tc_set_trace(void* buf);
int gemm_simt(const vdm_gemm_args* a, cudaStream_t stream);

}  // namespace vdm

extern "C" int vdm_version(void) { return 100; }
extern "C" const char* vdm_last_error_string(void) { return vdm::g_err; }
extern "C" int64_t vdm_launch_count(void) { return vdm::g_launches.load(); }
extern "C" void vdm_gemm_set_trace(void* buf) { vdm::gemm_tc_set_trace(buf); }

extern "C" int vdm_gemm(const vdm_gemm_args* a, vdm_stream_t stream) {
  VDM_REQUIRE(a != nullptr, "gemm: NULL args");
  VDM_REQUIRE(a->a1 && a->w, "gemm: NULL operand");
  VDM_REQUIRE(a->out_f32 || a->out_bf16, "gemm: no output");
  VDM_REQUIRE(a->n_img > 0 && a->H > 0 && a->W > 0 && a->N > 0, "gemm: bad geometry");
  VDM_REQUIRE(a->a1_coef == nullptr || a->dtype == VDM_BF16, "gemm: a1_coef (fused normalisation) is bf16-kernel only");
  VDM_REQUIRE(a->img_done == nullptr || a->dtype == VDM_BF16, "gemm: img_done is bf16-kernel only");
  VDM_REQUIRE(a->dtype == VDM_BF16 || a->dtype == VDM_F16 ||
                  (a->a2b == nullptr && a->C2b == 0 && a->a2_dtype != VDM_F16 && a->io_dtype != VDM_F16),
              "gemm: a2b / a2_dtype / io_dtype belong to the bf16 kernel");
  // VDM_F16: a plain linear whose activations AND weights are IEEE half (the normalised fp16 stream itself feeding the
  // attention qkv projection): the tcgen05 kernels with the f16 operand format
  if (a->dtype == VDM_BF16 || a->dtype == VDM_F16) return vdm::gemm_tc(a, (cudaStream_t)stream);
  if (a->dtype == VDM_F32) return vdm::gemm_simt(a, (cudaStream_t)stream);
  vdm::set_error("gemm: unknown dtype %d", a->dtype);
  return -1;
}

extern "C" int vdm_gemm_fused_norm_supported(const vdm_gemm_args* a) {
  if (a == nullptr || a->dtype != VDM_BF16 || a->taps != 9 || a->a1_mode != 0 || a->out_nchw) return 0;
  int ok = 0;
  if (vdm::gemm_tc(a, nullptr, &ok) != 0) return 0;
  return ok & 1;
}

extern "C" int vdm_gemm_img_done_supported(const vdm_gemm_args* a) {
  if (a == nullptr || a->dtype != VDM_BF16 || a->taps != 9 || a->a1_mode != 0 || a->out_nchw || a->a1_coef) return 0;
  int ok = 0;
  if (vdm::gemm_tc(a, nullptr, &ok) != 0) return 0;
  return (ok >> 1) & 1;
}
