// cfgpp_b200 — cross-attention (Nkv <= 128: the 77 text tokens) on tcgen05/TMEM. See attention.cuh for the contract.
//
// With a single K/V tile there is no running softmax, and the flash kernel of attention.cu spends its 23 us per launch
// on per-CTA latency (Q/K/V fetch -> QK -> softmax -> PV -> write-out, three rounds of 320 short CTAs). Here one CTA
// keeps K and V of its (batch, head) resident and streams `tiles_per_cta` consecutive 128-row query tiles through a
// software pipeline, two CTAs co-resident per SM (head dim 64, NC = 80: 98 KB shared memory, 256 TMEM columns, <= 128 registers each):
//   warp 0 : TMA producer   K, V once; Q tiles through a QS-deep ring
//   warp 1 : MMA issuer     S = Q_i K^T (M128 x NC x HD)  as soon as the softmax warps have pulled S(i-1) into
//                           registers; O[i & 1] = P_i V (M128 x HD x NC) as soon as P_i is in shared memory
//   warp 2 : TMEM allocator S [0,128)  O0 [128,128+HD)  O1 [128+HD,128+2HD)
//   warps 4..7 : softmax, one query row per thread: NC scores from TMEM, masked row max, p = exp2((s - m) scale log2e)
//                in fp32, row sum in fp32, P rounded to fp16 into swizzled smem; then — while the tensor pipe runs PV_i
//                and QK_(i+1) — the write-out of O_(i-1) (fp32 TMEM -> x 1/l -> fp16 -> global).
// NC = key columns processed: 80 when Nkv <= 80 (the 77-token prompt: 5 K-steps of 16 instead of 8, 80 exps per row
// instead of 128), else 128. Columns >= Nkv (zero rows of K / V from the TMA out-of-bounds fill) are masked to -inf.
#include <cmath>

#include "attention.cuh"
#include "common.cuh"

namespace cfgpp {

namespace {

constexpr int BQ = 128;
constexpr int ATOM_BYTES = 128 * 64 * 2;  // [128 rows x 64 fp16], 128B swizzle
constexpr int kThreads = 256;

template <int HD, int QS>
struct XCfg {
  static constexpr int NA = HD / 64;
  static constexpr int TILE_BYTES = NA * ATOM_BYTES;
  static constexpr int P_BYTES = 2 * ATOM_BYTES;
  static constexpr int SMEM_BYTES = TILE_BYTES * (QS + 2) + P_BYTES + 1024 + 256;
  static constexpr uint32_t TMEM_COLS = (128 + 2 * HD <= 256) ? 256 : 512;
  static constexpr uint32_t O_COL = 128;
  static_assert(128 + 2 * HD <= 512, "TMEM overflow");
  static_assert(SMEM_BYTES <= 232448, "shared memory overflow");
};

template <int HD, int QS, int NC>
__global__ void __launch_bounds__(kThreads, ((HD == 64 && NC == 80) ? 2 : 1))
xattn_kernel(const AttnParams p, const int tiles_per_cta, const __grid_constant__ CUtensorMap map_q,
             const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v) {
  using X = XCfg<HD, QS>;
  constexpr int TILE_BYTES = X::TILE_BYTES;
  constexpr int NA = X::NA;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                    // QS tiles
  uint8_t* sK = sQ + QS * TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;         // 2 atoms of 64 columns
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + X::P_BYTES);
  uint64_t* kv_full = bars;              // [1]
  uint64_t* q_full = kv_full + 1;        // [QS]
  uint64_t* q_empty = q_full + QS;       // [QS]
  uint64_t* s_full = q_empty + QS;       // [1]
  uint64_t* s_free = s_full + 1;         // [1]
  uint64_t* p_full = s_free + 1;         // [1]
  uint64_t* pv_done = p_full + 1;        // [1]
  uint64_t* o_free = pv_done + 1;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_free + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int total_tiles = (p.Nq + BQ - 1) / BQ;
  const int tile0 = blockIdx.x * tiles_per_cta;
  const int n_tiles = min(tiles_per_cta, total_tiles - tile0);

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < QS; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    mbar_init(&o_free[0], 128);
    mbar_init(&o_free[1], 128);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, X::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_launch_dependents();
  pdl_wait();

  if (warp_idx == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(kv_full, 2 * TILE_BYTES);
      for (int a = 0; a < NA; ++a) {
        tma_load_3d(sK + a * ATOM_BYTES, &map_k, kv_full, head * HD + a * 64, 0, batch);
        tma_load_3d(sV + a * ATOM_BYTES, &map_v, kv_full, head * HD + a * 64, 0, batch);
      }
      for (int i = 0; i < n_tiles; ++i) {
        const int s = i % QS;
        mbar_wait(&q_empty[s], ((i / QS) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], TILE_BYTES);
        for (int a = 0; a < NA; ++a)
          tma_load_3d(sQ + s * TILE_BYTES + a * ATOM_BYTES, &map_q, &q_full[s], head * HD + a * 64, (tile0 + i) * BQ,
                      batch);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, NC, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, HD, 0, 1);  // B (= V) is MN-major
      auto issue_qk = [&](int i) {
        const int s = i % QS;
        mbar_wait(&q_full[s], (i / QS) & 1);
        if (i > 0) mbar_wait(s_free, (i - 1) & 1);  // the softmax warps hold S(i-1) in registers
        tc_fence_after();
#pragma unroll
        for (int a = 0; a < NA; ++a) {
          const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ + s * TILE_BYTES + a * ATOM_BYTES), 1024, 0);
          const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + a * ATOM_BYTES), 1024, 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, (a | k) != 0 ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(&q_empty[s]);
      };
      mbar_wait(kv_full, 0);
      issue_qk(0);
      for (int i = 0; i < n_tiles; ++i) {
        if (i + 1 < n_tiles) issue_qk(i + 1);  // runs under the softmax of tile i
        mbar_wait(p_full, i & 1);
        if (i >= 2) mbar_wait(&o_free[i & 1], ((i >> 1) - 1) & 1);  // O(i-2) has been read out of this buffer
        tc_fence_after();
        const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV), 1024, ATOM_BYTES);
#pragma unroll
        for (int k = 0; k < NC / 16; ++k) {
          const uint64_t p_desc = make_sdesc_sw128(smem_u32(sP + (k >> 2) * ATOM_BYTES), 1024, 0) + 2 * (k & 3);
          umma_f16(tmem_base + X::O_COL + (i & 1) * HD, p_desc, v_desc + 128 * k, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(pv_done);
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax / output =====================
    const int qw = warp_idx & 3;
    const int row = qw * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_off;
    uint8_t* prow = sP + row * 128;
    const float c = p.scale_log2e;
    float inv_l_prev = 0.f;

    auto write_out = [&](int i, float inv_l) {  // O(i): TMEM fp32 -> x 1/l -> fp16 -> global
      const uint32_t o_addr = tmem_base + X::O_COL + (i & 1) * HD + lane_off;
      const int qrow = (tile0 + i) * BQ + row;
      __half* dst = p.out + (static_cast<size_t>(batch) * p.Nq + qrow) * p.ldo + head * HD;
#pragma unroll 1
      for (int h = 0; h < HD / 32; ++h) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + h * 32, o);
        tmem_ld_wait();
        if (qrow < p.Nq) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 w;
            w.x = pack_half2(__uint_as_float(o[8 * j + 0]) * inv_l, __uint_as_float(o[8 * j + 1]) * inv_l);
            w.y = pack_half2(__uint_as_float(o[8 * j + 2]) * inv_l, __uint_as_float(o[8 * j + 3]) * inv_l);
            w.z = pack_half2(__uint_as_float(o[8 * j + 4]) * inv_l, __uint_as_float(o[8 * j + 5]) * inv_l);
            w.w = pack_half2(__uint_as_float(o[8 * j + 6]) * inv_l, __uint_as_float(o[8 * j + 7]) * inv_l);
            reinterpret_cast<uint4*>(dst + h * 32)[j] = w;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&o_free[i & 1]);
    };

    for (int i = 0; i < n_tiles; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      uint32_t s[NC];
#pragma unroll
      for (int g = 0; g < NC / 16; ++g) tmem_ld_x16(s_addr + g * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[g * 16]));
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);  // scores are in registers: QK(i+1) may overwrite S
#pragma unroll
      for (int j = 0; j < NC; ++j)
        if (j >= p.Nkv) s[j] = 0xff800000u;  // -inf: padding columns (warp-uniform predicate)
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
#pragma unroll
      for (int j = 2; j < NC; j += 2) {
        mx0 = fmaxf(mx0, __uint_as_float(s[j]));
        mx1 = fmaxf(mx1, __uint_as_float(s[j + 1]));
      }
      const float mc = fmaxf(mx0, mx1) * c;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pk[NC / 2];
#pragma unroll
      for (int j = 0; j < NC / 2; ++j) {
        const float p0 = fast_exp2(__uint_as_float(s[2 * j]) * c - mc);
        const float p1 = fast_exp2(__uint_as_float(s[2 * j + 1]) * c - mc);
        rs0 += p0;
        rs1 += p1;
        pk[j] = pack_half2(p0, p1);
      }
      if (i > 0) mbar_wait(pv_done, (i - 1) & 1);  // PV(i-1) has consumed the P buffer (and O(i-1) is complete)
#pragma unroll
      for (int g = 0; g < NC / 8; ++g) {  // 16-byte chunks of the P row (two 128-byte halves)
        const int half_idx = g >> 3;
        const int ch = (g & 7) ^ (row & 7);  // 128B swizzle
        *reinterpret_cast<uint4*>(prow + half_idx * ATOM_BYTES + ch * 16) =
            make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      if (i > 0) {
        tc_fence_after();
        write_out(i - 1, inv_l_prev);  // overlaps PV(i) and QK(i+1) on the tensor pipe
      }
      inv_l_prev = 1.0f / (rs0 + rs1);
    }
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    write_out(n_tiles - 1, inv_l_prev);
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, X::TMEM_COLS);
  }
}

template <int HD, int QS, int NC>
void launch_x(const AttnOp& op, cudaStream_t stream) {
  using X = XCfg<HD, QS>;
  static bool configured = false;
  if (!configured) {
    CFGPP_CHECK_CUDA(cudaFuncSetAttribute(xattn_kernel<HD, QS, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          X::SMEM_BYTES));
    configured = true;
  }
  // tiles per CTA: as few as keeps every CTA co-resident in one wave (two per SM for head dim 64), at least 2 so the
  // K / V fetch and the pipeline fill are amortised
  const int total_tiles = (op.p.Nq + BQ - 1) / BQ;
  const int slots = num_sms() * ((HD == 64 && NC == 80) ? 2 : 1);
  const int bh = op.p.B * op.p.H;
  int tpc = 2;
  while (tpc < total_tiles && ((total_tiles + tpc - 1) / tpc) * bh > slots) ++tpc;
  if (tpc > total_tiles) tpc = total_tiles;
  dim3 grid((total_tiles + tpc - 1) / tpc, op.p.H, op.p.B);
  launch_pdl(xattn_kernel<HD, QS, NC>, grid, dim3(kThreads), X::SMEM_BYTES, stream, op.p, tpc, op.map_q, op.map_k,
             op.map_v);
}

}  // namespace

// opt every instantiation into its dynamic shared memory size (called once per process, outside graph capture)
void xattn_configure() {
  auto cfg = [](auto kernel, int bytes) {
    CFGPP_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  };
  cfg(xattn_kernel<64, 2, 80>, XCfg<64, 2>::SMEM_BYTES);
  cfg(xattn_kernel<64, 2, 128>, XCfg<64, 2>::SMEM_BYTES);
  cfg(xattn_kernel<128, 2, 80>, XCfg<128, 2>::SMEM_BYTES);
  cfg(xattn_kernel<128, 2, 128>, XCfg<128, 2>::SMEM_BYTES);
  cfg(xattn_kernel<192, 1, 80>, XCfg<192, 1>::SMEM_BYTES);
  cfg(xattn_kernel<192, 1, 128>, XCfg<192, 1>::SMEM_BYTES);
}

// (head dim 192 at two query tiles measured slower than the flash kernel: one CTA per SM and a single-stage Q ring)
bool xattn_applicable(const AttnOp& op) {
  return op.p.Nkv <= 128 && op.p.Nq >= (op.hd_pad == 192 ? 4 : 2) * BQ;
}

void run_xattn_op(const AttnOp& op, cudaStream_t stream) {
  const bool narrow = op.p.Nkv <= 80;
  switch (op.hd_pad) {
    case 64: return narrow ? launch_x<64, 2, 80>(op, stream) : launch_x<64, 2, 128>(op, stream);
    case 128: return narrow ? launch_x<128, 2, 80>(op, stream) : launch_x<128, 2, 128>(op, stream);
    case 192: return narrow ? launch_x<192, 1, 80>(op, stream) : launch_x<192, 1, 128>(op, stream);
    default: throw Error(-1, "unsupported padded head dim");
  }
}

}  // namespace cfgpp
