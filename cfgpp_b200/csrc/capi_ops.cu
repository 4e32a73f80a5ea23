// cfgpp_b200 — C ABI, operator-level entry points (one call = one kernel launch on the caller's stream).
// Declared in include/cfgpp_b200.h. No C++ exception crosses the boundary: every entry point returns an int
// status (0 = OK) and records a message retrievable with cfgpp_last_error().
#include "capi_util.h"
#include "attention.cuh"
#include "gemm.cuh"
#include "ops.cuh"
#include "../../include/cfgpp_b200.h"

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

// Debug aid (not in the public header): run the linear op `iters` times and return per-CTA timestamps of the last run.
CFGPP_API int cfgpp_dbg_linear_timeline(const void* a, int lda, const void* w, int M, int N, int K, const void* bias,
                                        const void* addend, void* out, int force_bn, int iters,
                                        unsigned long long* host_out /*[grid][16]*/, int* grid_out, void* stream) {
  return guarded([&] {
    GemmOp op = make_linear_op((const __half*)a, lda, nullptr, 0, 0, (const __half*)w, M, N, K, (const __half*)bias,
                               (const __half*)addend, N, 1, (__half*)out, N, false, force_bn);
    unsigned long long* d = nullptr;
    CFGPP_CHECK_CUDA(cudaMalloc(&d, sizeof(unsigned long long) * 16 * op.grid));
    CFGPP_CHECK_CUDA(cudaMemset(d, 0, sizeof(unsigned long long) * 16 * op.grid));
    op.p.timeline = d;
    for (int i = 0; i < iters; ++i) run_gemm_op(op, (cudaStream_t)stream);
    CFGPP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    CFGPP_CHECK_CUDA(cudaMemcpy(host_out, d, sizeof(unsigned long long) * 16 * op.grid, cudaMemcpyDeviceToHost));
    *grid_out = op.grid;
    cudaFree(d);
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

CFGPP_API int cfgpp_op_conv3x3_s2(const void* x, int B, int H, int W, int Cin, const void* w, int Cout, const void* bias,
                                  int pad, void* out, void* stream) {
  return guarded([&] {
    GemmOp op = make_conv3x3_op((const __half*)x, B, H, W, Cin, (const __half*)w, Cout, (const __half*)bias, nullptr, 0, 1,
                                (__half*)out, 0, 2, pad);
    run_gemm_op(op, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_op_attention(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, void* out,
                                 int ldo, int B, int H, int Nq, int Nkv, int head_dim, void* stream) {
  return guarded([&] {
    AttnOp op = make_attn_op((const __half*)q, ldq, (const __half*)k, ldk, (const __half*)v, ldv, (__half*)out, ldo,
                             B, H, Nq, Nkv, head_dim);
    run_attn_op(op, (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_op_groupnorm(const void* x1, int C1, const void* x2, int C2, int B, int HW, const void* gamma,
                                 const void* beta, float eps, int silu, void* out, void* stream) {
  return guarded([&] {
    float* partial = nullptr;
    CFGPP_CHECK_CUDA(cudaMalloc(&partial, gn_partial_floats(B, HW) * sizeof(float)));
    try {
      run_groupnorm((const __half*)x1, C1, (const __half*)x2, C2, B, HW, (const __half*)gamma, (const __half*)beta,
                    eps, silu != 0, partial, (__half*)out, (cudaStream_t)stream);
    } catch (...) {
      cudaFree(partial);
      throw;
    }
    CFGPP_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));  // test-only entry point: scratch freed below
    cudaFree(partial);
  });
}

CFGPP_API int cfgpp_op_layernorm(const void* x, int M, int C, const void* gamma, const void* beta, float eps,
                                 void* out, void* stream) {
  return guarded([&] {
    run_layernorm((const __half*)x, M, C, (const __half*)gamma, (const __half*)beta, eps, (__half*)out,
                  (cudaStream_t)stream);
  });
}

CFGPP_API int cfgpp_op_cfgpp_step(const void* eps_uc, const void* eps_c, int n, int method, int state_dtype,
                                  const cfgpp_step_coef* coef_host, void* z, void* aux, void* z0t_out,
                                  const void* noise_dev, void* stream) {
  return guarded([&] {
    static_assert(sizeof(cfgpp_step_coef) == sizeof(StepCoef), "ABI struct mismatch");
    static_assert(sizeof(StepCoef) % sizeof(void*) == 0, "the noise word is stored right behind the coefficients");
    StepCoef* coef_dev = nullptr;
    CFGPP_CHECK_CUDA(cudaMalloc(&coef_dev, sizeof(StepCoef) + sizeof(void*)));
    const __half** slot = reinterpret_cast<const __half**>(coef_dev + 1);
    cudaError_t e = cudaMemcpy(coef_dev, coef_host, sizeof(StepCoef), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(slot, &noise_dev, sizeof(void*), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
      try {
        run_step_only((const __half*)eps_uc, (const __half*)eps_c, n, method | (state_dtype == CFGPP_F16 ? 0x100 : 0),
                      coef_dev, z, aux, z0t_out, (cudaStream_t)stream, slot);
      } catch (...) {
        cudaFree(coef_dev);
        throw;
      }
      e = cudaStreamSynchronize((cudaStream_t)stream);  // test-only entry point
    }
    cudaFree(coef_dev);
    CFGPP_CHECK_CUDA(e);
  });
}

}  // extern "C"
