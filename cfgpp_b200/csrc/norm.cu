// cfgpp_b200 — GroupNorm(32)(+SiLU) and LayerNorm on NHWC / token-major fp16 activations (HBM-bound kernels).
//
// Numerics follow the reference's autocast graph: statistics and the affine transform (and SiLU) are computed
// in fp32 from the fp16 input, and the result is rounded to fp16 exactly once — the point where the reference's
// fp32 norm output is cast for the following conv / linear.
//
// GroupNorm is two launches: (1) per-(sample, pixel-chunk) partial (sum, sum^2) per group — deterministic, no
// atomics in global memory; (2) apply, which reduces the partials on the fly. Both take an optional second source
// so that torch.cat([h, skip], dim=1) of the up-blocks is never materialised un-normalised.
#include "common.cuh"
#include "ops.cuh"

namespace cfgpp {

namespace {

constexpr int GROUPS = 32;

struct GnSrc {
  const __half* x1;
  const __half* x2;
  int C1, C2;
};

CFGPP_DEVICE uint4 load_vec(const GnSrc& s, size_t pix, int c) {  // c multiple of 8, never straddles C1
  if (c < s.C1) return *reinterpret_cast<const uint4*>(s.x1 + pix * s.C1 + c);
  return *reinterpret_cast<const uint4*>(s.x2 + pix * s.C2 + (c - s.C1));
}

// grid (nchunk, B); block = vpp * k threads (vpp = C / 8 vectors per pixel) so a thread keeps one channel vector.
__global__ void gn_stats_kernel(GnSrc src, int HW, int C, int px_per_block, float* __restrict__ partial) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sm[];  // [pstep][C] sums, then [pstep][C] sums of squares (one slot per thread: no atomics)
  const int vpp = C >> 3;
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int vec = threadIdx.x % vpp;
  const int prow = threadIdx.x / vpp;
  const int pstep = blockDim.x / vpp;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
  const int p0 = chunk * px_per_block;
  const int pend = min(px_per_block, HW - p0);
  const size_t base = static_cast<size_t>(b) * HW + p0;
  constexpr int U = 4;  // independent 16 B loads in flight per thread
  int pp = prow;
  for (; pp + (U - 1) * pstep < pend; pp += U * pstep) {
    uint4 u[U];
#pragma unroll
    for (int j = 0; j < U; ++j) u[j] = load_vec(src, base + pp + j * pstep, vec * 8);
#pragma unroll
    for (int j = 0; j < U; ++j) {
      const __half2* h = reinterpret_cast<const __half2*>(&u[j]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h[i]);
        s[2 * i] += f.x;
        q[2 * i] += f.x * f.x;
        s[2 * i + 1] += f.y;
        q[2 * i + 1] += f.y * f.y;
      }
    }
  }
  for (; pp < pend; pp += pstep) {
    const uint4 u = load_vec(src, base + pp, vec * 8);
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(h[i]);
      s[2 * i] += f.x;
      q[2 * i] += f.x * f.x;
      s[2 * i + 1] += f.y;
      q[2 * i + 1] += f.y * f.y;
    }
  }
  float* sq = sm + pstep * C;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sm[prow * C + vec * 8 + i] = s[i];
    sq[prow * C + vec * 8 + i] = q[i];
  }
  __syncthreads();
  const int cpg = C / GROUPS;
  if (threadIdx.x < GROUPS) {  // fixed summation order -> bit-reproducible statistics
    float a = 0.f, bsum = 0.f;
    for (int r = 0; r < pstep; ++r)
      for (int i = 0; i < cpg; ++i) {
        a += sm[r * C + threadIdx.x * cpg + i];
        bsum += sq[r * C + threadIdx.x * cpg + i];
      }
    float* dst = partial + ((static_cast<size_t>(b) * gridDim.x + chunk) * GROUPS + threadIdx.x) * 2;
    dst[0] = a;
    dst[1] = bsum;
  }
}

// grid (nchunk, B); block = vpp * k threads: a thread owns one 8-channel vector, so the per-channel affine
// ((x - mean) * rstd) * gamma + beta (the reference's evaluation order) uses 32 registers set up once per thread.
__global__ void gn_apply_kernel(GnSrc src, int HW, int C, int px_per_block, const float* __restrict__ partial,
                                int nchunk, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                                float eps, int silu, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_mean[GROUPS], s_rstd[GROUPS];
  __shared__ float2 s_part[8][GROUPS];
  const int b = blockIdx.y;
  // Cross-chunk reduction of the partial sums, spread over 8 x 32 threads (each sums every 8th chunk, loads
  // independent), then combined in a fixed order: a single thread per group walking all (up to 128) chunks exposed
  // ~10 us of serialised L2 latency at the head of EVERY block.
  for (int idx = threadIdx.x; idx < 8 * GROUPS; idx += blockDim.x) {
    const int g = idx & (GROUPS - 1), part = idx / GROUPS;
    float a = 0.f, q = 0.f;
    const float* src_p = partial + (static_cast<size_t>(b) * nchunk * GROUPS + g) * 2;
    for (int i = part; i < nchunk; i += 8) {
      const float2 v = *reinterpret_cast<const float2*>(src_p + static_cast<size_t>(i) * GROUPS * 2);
      a += v.x;
      q += v.y;
    }
    s_part[part][g] = make_float2(a, q);
  }
  __syncthreads();
  if (threadIdx.x < GROUPS) {
    float a = 0.f, q = 0.f;
#pragma unroll
    for (int part = 0; part < 8; ++part) {
      a += s_part[part][threadIdx.x].x;
      q += s_part[part][threadIdx.x].y;
    }
    const float n = static_cast<float>(HW) * (C / GROUPS);
    const float mean = a / n;
    const float var = fmaxf(q / n - mean * mean, 0.f);
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + eps);
  }
  __syncthreads();
  const int vpp = C >> 3;
  const int cpg = C / GROUPS;
  const int vec = threadIdx.x % vpp;
  const int prow = threadIdx.x / vpp;
  const int pstep = blockDim.x / vpp;
  const int c0 = vec * 8;
  float scale[8], shift[8];  // rstd / mean of the group each of this thread's 8 channels belongs to
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int g = (c0 + k) / cpg;
    scale[k] = s_rstd[g];
    shift[k] = s_mean[g];
  }
  const uint4 ug = *reinterpret_cast<const uint4*>(gamma + c0);
  const uint4 ub = *reinterpret_cast<const uint4*>(beta + c0);
  const __half* hg = reinterpret_cast<const __half*>(&ug);
  const __half* hb = reinterpret_cast<const __half*>(&ub);
  float gm[8], bt[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    gm[k] = __half2float(hg[k]);
    bt[k] = __half2float(hb[k]);
  }
  const int p0 = blockIdx.x * px_per_block;
  const int pend = min(px_per_block, HW - p0);
  const size_t base = static_cast<size_t>(b) * HW + p0;
  constexpr int U = 4;
  auto emit = [&](const uint4& u, size_t gp) {
    const __half* hx = reinterpret_cast<const __half*>(&u);
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = (__half2float(hx[k]) - shift[k]) * scale[k] * gm[k] + bt[k];
      if (silu) v = silu_f(v);
      y[k] = v;
    }
    uint4 o;
    o.x = pack_half2(y[0], y[1]);
    o.y = pack_half2(y[2], y[3]);
    o.z = pack_half2(y[4], y[5]);
    o.w = pack_half2(y[6], y[7]);
    *reinterpret_cast<uint4*>(out + gp * C + c0) = o;
  };
  int pp = prow;
  for (; pp + (U - 1) * pstep < pend; pp += U * pstep) {
    uint4 u[U];
#pragma unroll
    for (int j = 0; j < U; ++j) u[j] = load_vec(src, base + pp + j * pstep, c0);
#pragma unroll
    for (int j = 0; j < U; ++j) emit(u[j], base + pp + j * pstep);
  }
  for (; pp < pend; pp += pstep) emit(load_vec(src, base + pp, c0), base + pp);
}

// one warp per row; C % 8 == 0, C <= 2048
__global__ void layernorm_kernel(const __half* __restrict__ x, int M, int C, const __half* __restrict__ gamma,
                                 const __half* __restrict__ beta, float eps, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C >> 3;
  constexpr int MAXV = 8;
  uint4 v[MAXV];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      v[i] = *reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * C + vi * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        sum += f.x + f.y;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const __half2* h = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        sq += (f.x - mean) * (f.x - mean) + (f.y - mean) * (f.y - mean);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq / C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 ug = *reinterpret_cast<const uint4*>(gamma + vi * 8);
      const uint4 ub = *reinterpret_cast<const uint4*>(beta + vi * 8);
      const __half* hx = reinterpret_cast<const __half*>(&v[i]);
      const __half* hg = reinterpret_cast<const __half*>(&ug);
      const __half* hb = reinterpret_cast<const __half*>(&ub);
      float y[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        y[k] = (__half2float(hx[k]) - mean) * rstd * __half2float(hg[k]) + __half2float(hb[k]);
      uint4 o;
      o.x = pack_half2(y[0], y[1]);
      o.y = pack_half2(y[2], y[3]);
      o.z = pack_half2(y[4], y[5]);
      o.w = pack_half2(y[6], y[7]);
      *reinterpret_cast<uint4*>(out + static_cast<size_t>(row) * C + vi * 8) = o;
    }
  }
}

}  // namespace

// up to 128 pixel chunks per sample (the partial-sum buffer holds 128) so that even the 32x32 level launches
// >= 512 blocks; at least 4 pixels per block
int gn_px_per_block(int HW) {
  int ppb = (HW + 127) / 128;
  return ppb < 4 ? 4 : ppb;
}
int gn_num_chunks(int HW) { return (HW + gn_px_per_block(HW) - 1) / gn_px_per_block(HW); }
size_t gn_partial_floats(int B, int HW) { return static_cast<size_t>(B) * gn_num_chunks(HW) * GROUPS * 2; }

void run_groupnorm(const __half* x1, int C1, const __half* x2, int C2, int B, int HW, const __half* gamma,
                   const __half* beta, float eps, bool silu, float* partial, __half* out, cudaStream_t stream) {
  const int C = C1 + C2;
  CFGPP_REQUIRE(C % GROUPS == 0 && C % 8 == 0 && C1 % 8 == 0, "GroupNorm needs C % 32 == 0 and 8-aligned sources");
  GnSrc src{x1, x2 ? x2 : x1, C1, C2};
  const int vpp = C / 8;
  const int ppb = gn_px_per_block(HW);
  const int nchunk = gn_num_chunks(HW);
  int k = 256 / vpp;
  if (k < 1) k = 1;
  const int threads = vpp * k;
  CFGPP_REQUIRE(threads <= 1024, "GroupNorm channel count too large");
  launch_pdl(gn_stats_kernel, dim3(nchunk, B), dim3(threads), 2 * static_cast<size_t>(k) * C * sizeof(float), stream, src, HW, C,
             ppb, partial);
  launch_pdl(gn_apply_kernel, dim3(nchunk, B), dim3(threads), 0, stream, src, HW, C, ppb, partial, nchunk, gamma, beta, eps,
                                                           silu ? 1 : 0, out);
}

void run_layernorm(const __half* x, int M, int C, const __half* gamma, const __half* beta, float eps, __half* out,
                   cudaStream_t stream) {
  CFGPP_REQUIRE(C % 8 == 0 && C <= 2048, "LayerNorm needs C % 8 == 0 and C <= 2048");
  const int warps = 8;
  launch_pdl(layernorm_kernel, dim3((M + warps - 1) / warps), dim3(warps * 32), 0, stream, x, M, C, gamma, beta, eps, out);
}

}  // namespace cfgpp
