// Bring-up test for round 2 (lead 1): tcgen05.mma with the A operand in TENSOR MEMORY.
//   D[128 x 64] (fp32, TMEM) = A[128 x 128] (fp16, TMEM, written with tcgen05.st) * B[64 x 128]^T (fp16, smem, K-major,
//   128B swizzle)
// This is the shape of the attention PV product when P goes softmax -> TMEM -> MMA instead of through shared memory.
// NOT part of the library (nothing here is built into libcfgpp_b200.so). Result on B200 (round 1, last GPU call):
// "max |err| = 0 -> OK", i.e. the layout below is what the hardware expects: (1) a K-major fp16 A lives with lane = row
// and one 32-bit TMEM column = two consecutive K elements (exactly what tcgen05.st.32x32b writes when thread i holds
// row i as packed half2 words), (2) a K = 16 instruction reads 8 columns, so K step k starts at column base + 8 k.
//   nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I cfgpp_b200/csrc tools/bringup/tmem_a_mma.cu -o /tmp/tmem_a_mma && /tmp/tmem_a_mma
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

using namespace cfgpp;

namespace {

constexpr int M = 128, N = 64, K = 128;
constexpr uint32_t TMEM_COLS = 256;   // D: 64 columns at 0, A: 64 columns (128 fp16) at 128
constexpr uint32_t A_COL = 128;

// D[tmem] (+)= A[tmem] * B[smem desc]
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) ts_mma_kernel(const __half* __restrict__ a, const __half* __restrict__ b,
                                                        float* __restrict__ d) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* sB = smem;                                           // two atoms of [64 rows x 128 B], 8 KB each
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * 8192);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, row = threadIdx.x;

  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  // B (N x K, K-major) into the 128B-swizzled layout the descriptors describe
  for (int i = threadIdx.x; i < N * K / 8; i += blockDim.x) {    // 16-byte pieces
    const int n = i / (K / 8), k8 = i % (K / 8);
    const int atom = k8 / 8, chunk = k8 % 8;
    const uint4 v = *reinterpret_cast<const uint4*>(b + n * K + k8 * 8);
    *reinterpret_cast<uint4*>(sB + atom * 8192 + n * 128 + ((chunk ^ (n & 7)) << 4)) = v;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // A: thread = row, 64 packed half2 words = columns A_COL .. A_COL + 63 of lane `row`
  const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
  for (int h = 0; h < 2; ++h) {
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) w[j] = *reinterpret_cast<const uint32_t*>(a + row * K + (h * 32 + j) * 2);
    tmem_st_x32(tmem_base + A_COL + h * 32 + lane_off, w);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = make_idesc_f16(M, N, 0, 0);
#pragma unroll
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t b_desc = make_sdesc_sw128(smem_u32(sB + (k >> 2) * 8192), 1024, 0) + 2 * (k & 3);
      umma_f16_ts(tmem_base, tmem_base + A_COL + 8 * k, b_desc, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld_x32(tmem_base + h * 32 + lane_off, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) d[row * N + h * 32 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int main() {
  std::vector<__half> ha(M * K), hb(N * K);
  std::vector<float> fa(M * K), fb(N * K), ref(M * N, 0.f), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { fa[i] = (rand() % 17 - 8) / 8.0f; ha[i] = __float2half(fa[i]); }
  for (int i = 0; i < N * K; ++i) { fb[i] = (rand() % 13 - 6) / 8.0f; hb[i] = __float2half(fb[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += fa[m * K + k] * fb[n * K + k];
      ref[m * N + n] = s;
    }
  __half *da, *db;
  float* dd;
  cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, out.size() * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  const int smem = 2 * 8192 + 1024 + 64;
  cudaFuncSetAttribute(ts_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  ts_mma_kernel<<<1, 128, smem>>>(da, db, dd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0;
  for (int i = 0; i < M * N; ++i) worst = fmax(worst, fabs(out[i] - ref[i]));
  printf("A-from-TMEM MMA 128x64x128: max |err| = %g  -> %s\n", worst, worst < 1e-3 ? "OK" : "MISMATCH (layout assumption wrong)");
  return worst < 1e-3 ? 0 : 2;
}
