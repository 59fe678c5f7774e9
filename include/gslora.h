/* gslora-b200 -- C ABI of the Blackwell-native GS-LoRA unlearning hot path.
 *
 * The reference (bjzhb666/GS-LoRA) is pure Python/PyTorch and has NO FFI: its seam is the Python module
 * surface.  These entry points are what a ctypes binding on the reference side calls
 * (see INTEGRATION.md); each cites the reference code it replaces.  Conventions:
 *   - every pointer is a raw DEVICE pointer owned by the caller (PyTorch's allocator); the library borrows
 *     it for the call, allocates nothing persistent and never synchronises; `stream` is a cudaStream_t.
 *   - int64 sizes / leading dimensions in ELEMENTS; return 0 = ok, otherwise a cudaError_t or -1 with a
 *     message available from gsl_last_error().
 *   - no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef GSLORA_H
#define GSLORA_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* gsl_last_error(void);
int gsl_version(void);
/* default tcgen05 cta_group (1 or 2) for the GEMM family */
void gsl_set_gemm_cta_group(int cta_group);

/* epilogues of gsl_gemm_f16 */
enum {
  GSL_EPI_F16 = 0,          /* out0 fp16 = acc + bias                                              */
  GSL_EPI_F32 = 1,          /* out0 fp32 = acc + bias                     [+ out1 fp16 copy]        */
  GSL_EPI_GELU = 2,         /* out0 fp16 = h = acc + bias, out1 fp16 = gelu_erf(h)                  */
  GSL_EPI_GELU_BWD = 3,     /* out0 fp16 = acc * gelu'(aux fp16)                                    */
  GSL_EPI_RES_F32 = 4,      /* out0 fp32 = acc + bias + aux fp32          [+ out1 fp16 copy]        */
  GSL_EPI_PERIODIC_F32 = 5  /* out0 fp32 = acc + aux_table fp32[row % aux_period]                   */
};

/* C[M,N] = epi(A[M,K] * B[N,K]^T), fp16 operands, fp32 accumulation on tcgen05 tensor cores.
 * Replaces torch.nn.functional.linear / loralib.Linear.forward on the hot path
 * (vit_pytorch_face/vit_face.py:330-334,360,377,531; loralib Linear.forward: the LoRA term is an extra
 * K = 16 step when A carries T = x*A^T and B carries s*lora_B in 16 trailing columns) and the dX GEMMs
 * autograd builds for engine_cl.py:124. */
int gsl_gemm_f16(const void* A, int64_t lda, const void* B, int64_t ldb, int64_t M, int64_t N, int64_t K,
                 int epi, const float* bias, void* out0, int64_t ld0, void* out1, int64_t ld1,
                 const void* aux, int64_t ldaux, int64_t aux_period, int cta_group, int block_n, void* stream);

#ifdef __cplusplus
}
#endif
#endif
