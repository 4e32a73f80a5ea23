// cfgpp_b200 — C ABI, operator-level entry points (one call = one kernel launch on the caller's stream).
// Declared in include/cfgpp_b200.h. No C++ exception crosses the boundary: every entry point returns an int
// status (0 = OK) and records a message retrievable with cfgpp_last_error().
#include "capi_util.h"
#include "gemm.cuh"

using namespace cfgpp;

extern "C" {

CFGPP_API int cfgpp_op_linear(const void* a, int lda, const void* a2, int lda2, int k_split, const void* w, int M,
                              int N, int K, const void* bias, const void* addend, int ld_add,
                              int add_rows_per_group, void* out, int ldc, int geglu, int force_bn, void* stream) {
  return guarded([&] {
    GemmOp op = make_linear_op((const __half*)a, lda, (const __half*)a2, lda2, k_split, (const __half*)w, M, N, K,
                               (const __half*)bias, (const __half*)addend, ld_add, add_rows_per_group, (__half*)out,
                               ldc, geglu != 0, force_bn);
    run_gemm_op(op, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_op_conv3x3(const void* x, int B, int H, int W, int Cin, const void* w, int Cout, const void* bias,
                               const void* addend, int ld_add, int add_rows_per_group, void* out, int force_bn,
                               void* stream) {
  return guarded([&] {
    GemmOp op = make_conv3x3_op((const __half*)x, B, H, W, Cin, (const __half*)w, Cout, (const __half*)bias,
                                (const __half*)addend, ld_add, add_rows_per_group, (__half*)out, force_bn);
    run_gemm_op(op, (cudaStream_t)stream);
  });
}

}  // extern "C"
