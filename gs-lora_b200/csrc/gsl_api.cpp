// gslora-b200: extern "C" surface declared in include/gslora.h.
#include "../../include/gslora.h"
#include "gsl_kernels.h"

#include <cstdarg>
#include <cstdio>

namespace gsl {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gsl

using namespace gsl;

extern "C" {

const char* gsl_last_error(void) { return g_err; }
int gsl_version(void) { return 100; }
void gsl_set_gemm_cta_group(int cta_group) { gemm_set_default_cta_group(cta_group); }

int gsl_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K, int epi,
                 const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1, const void* aux, int64_t ldaux,
                 int64_t aux_period, int cta_group, int block_n, void* stream) {
    GemmArgs a;
    a.A = (const __half*)A; a.lda = lda; a.B = (const __half*)B; a.ldb = ldb;
    a.M = M; a.N = N; a.K = K; a.epi = epi; a.bias = bias;
    a.out0 = out0; a.ld0 = ld0; a.out1 = out1; a.ld1 = ld1;
    a.aux = aux; a.ldaux = ldaux; a.aux_period = aux_period;
    a.cta_group = cta_group; a.block_n = block_n;
    return gemm_f16(a, (cudaStream_t)stream);
}

}  // extern "C"
