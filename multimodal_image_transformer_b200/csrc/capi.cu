// extern "C" surface declared in include/b200_decoder.h: argument validation + dispatch only.
#include "../../include/b200_decoder.h"
#include "common.cuh"
#include "gemm.cuh"
#include "kernels.cuh"
#include <stdarg.h>
#include <string.h>

namespace b200 {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace b200

using namespace b200;

extern "C" {

int b200_version(void) { return B200_ABI_VERSION; }
const char* b200_last_error(void) { return g_err; }

int b200_check_device(int dev) {
  int major = 0, minor = 0;
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  B200_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  B200_REQUIRE(major == 10 && minor == 0, "device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
  return 0;
}

static int to_problem(const b200_gemm_args* a, GemmProblem* p) {
  B200_REQUIRE(a != nullptr, "gemm: null args");
  p->M = a->M; p->N = a->N; p->K = a->K;
  p->A = static_cast<const bf16*>(a->A); p->lda = a->lda; p->a_mn = a->a_mn_major != 0;
  p->B = static_cast<const bf16*>(a->B); p->ldb = a->ldb; p->b_mn = a->b_mn_major != 0;
  p->D = a->D; p->ldd = a->ldd; p->d_fp32 = a->d_fp32 != 0; p->accumulate = a->accumulate != 0;
  p->bias = a->bias;
  p->residual = static_cast<const bf16*>(a->residual); p->ldr = a->ldr;
  p->relu_mask = static_cast<const bf16*>(a->relu_mask); p->ldm = a->ldm;
  p->act = a->act; p->split_k = a->split_k; p->block_n = a->block_n;
  return 0;
}

int b200_gemm(const b200_gemm_args* a, void* stream) {
  GemmProblem p;
  if (int rc = to_problem(a, &p)) return rc;
  return gemm_launch(p, static_cast<cudaStream_t>(stream));
}

int b200_gemm_check(const b200_gemm_args* a, void* stream) {
  GemmProblem p;
  if (int rc = to_problem(a, &p)) return rc;
  return gemm_check_launch(p, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
