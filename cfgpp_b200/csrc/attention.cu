// cfgpp_b200 — flash-style attention forward for head_dim 64 on tcgen05/TMEM (sm_100a). See attention.cuh.
//
// One CTA = 128 query rows of one (batch, head); loop over 128-wide KV tiles.
//   warp 0 lane 0 : TMA producer (Q once; K/V rings, 128B swizzle)
//   warp 1 lane 0 : MMA issuer   S_b = Q K_j^T  (M128 N128 K64 -> TMEM, double-buffered b = j&1)
//                                O_t = P_j V_j  (M128 N64 K128, A = P from smem, B = V MN-major)
//   warp 2        : TMEM allocator (512 columns: S0 [0,128) S1 [128,256) O_t [256,320))
//   warps 4..7    : softmax, one query row per thread: online max / exp2 / running sum in fp32, P rounded to
//                   fp16 into swizzled smem (the SS-operand of the PV MMA), O accumulated in registers with the
//                   usual exp2((m_old - m_new) c) rescale, final 1/l normalisation and fp16 store.
// QK_{j+1} is issued before PV_j so the tensor pipe works on the next scores while the softmax warps are busy.
#include "attention.cuh"
#include "common.cuh"

namespace cfgpp {

namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 64;
constexpr int KS = 3;  // K / V ring depth
constexpr int TILE_BYTES = 128 * HD * 2;  // 16 KB: Q, K, V tiles and each of the two P sub-tiles
constexpr int SMEM_BYTES = TILE_BYTES * (1 + 2 * KS + 2) + 1024 + 256;
constexpr int kThreads = 256;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t O_COL = 256;

__global__ void __launch_bounds__(kThreads, 1)
attn_kernel(const AttnParams p, const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
            const __grid_constant__ CUtensorMap map_v) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;
  uint8_t* sV = sK + KS * TILE_BYTES;
  uint8_t* sP = sV + KS * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + KS;
  uint64_t* v_full = k_empty + KS;
  uint64_t* v_empty = v_full + KS;
  uint64_t* s_full = v_empty + KS;
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* pv_done = p_full + 1;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pv_done + 1);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ;
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_tiles = (p.Nkv + BKV - 1) / BKV;

  if (warp_idx == 0 && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp_idx == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
    }
    mbar_init(p_full, 128);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp_idx == 2) {
    tmem_alloc(tmem_ptr_smem, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp_idx == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_3d(sQ, &map_q, q_full, head * HD, q0, batch);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j % KS;
        const uint32_t ph = (j / KS) & 1;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
        tma_load_3d(sK + s * TILE_BYTES, &map_k, &k_full[s], head * HD, j * BKV, batch);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
        tma_load_3d(sV + s * TILE_BYTES, &map_v, &v_full[s], head * HD, j * BKV, batch);
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, HD, 0, 1);  // B (= V) is MN-major
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ), 1024, 0);
      auto issue_qk = [&](int j) {
        const int s = j % KS;
        const int b = j & 1;
        mbar_wait(&k_full[s], (j / KS) & 1);
        mbar_wait(&s_empty[b], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK + s * TILE_BYTES), 1024, 0);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_f16(tmem_base + b * BKV, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[b]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) issue_qk(j + 1);
        const int s = j % KS;
        mbar_wait(&v_full[s], (j / KS) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        // V tile: 128 kv rows x 64 d (128 B per row, swizzled) = MN-major B operand with a single 64-wide MN atom:
        // 8-row K groups are 1024 B apart (SBO); a K step of 16 rows advances 2048 B.
        const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV + s * TILE_BYTES), 1024, TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k) {
          const uint64_t p_desc = make_sdesc_sw128(smem_u32(sP + (k >> 2) * TILE_BYTES), 1024, 0) + 2 * (k & 3);
          umma_f16(tmem_base + O_COL, p_desc, v_desc + 128 * k, idesc_pv, k != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(pv_done);
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax / output =====================
    const int qw = warp_idx - 4;
    const int row = qw * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const float c = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f, alpha_pending = 0.f;
    float o_acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o_acc[d] = 0.f;

    auto fold_o = [&]() {  // o_acc = o_acc * alpha_pending + O_t
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + O_COL + h * 32 + lane_off, v);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) o_acc[h * 32 + d] = o_acc[h * 32 + d] * alpha_pending + __uint_as_float(v[d]);
      }
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int b = j & 1;
      const int valid = p.Nkv - j * BKV;  // columns >= valid are padding
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t s_addr = tmem_base + b * BKV + lane_off;
      // pass A: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(s_addr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = (c0 + i < valid) ? __uint_as_float(v[i]) : -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * c);
      const float mc = m_new * c;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // PV_{j-1} retired: O_t readable, P buffer reusable
        tc_fence_after();
        fold_o();
      }
      // pass B: p = exp2(s c - m c); row sum; fp16 P into the swizzled K-major A tile(s)
      float rs = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < BKV; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(s_addr + c0, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = (c0 + 2 * i < valid) ? fast_exp2(__uint_as_float(v[2 * i]) * c - mc) : 0.f;
          float p1 = (c0 + 2 * i + 1 < valid) ? fast_exp2(__uint_as_float(v[2 * i + 1]) * c - mc) : 0.f;
          rs += p0 + p1;
          pk[i] = pack_half2(p0, p1);
        }
        uint8_t* prow = sP + (c0 >> 6) * TILE_BYTES + row * 128;
        const int chunk0 = (c0 & 63) >> 3;  // first 16-byte chunk of this 32-column group within the 128 B row
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ch = (chunk0 + i) ^ (row & 7);
          *reinterpret_cast<uint4*>(prow + ch * 16) = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&s_empty[b]);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
      l_run = l_run * alpha + rs;
      alpha_pending = alpha;
      m_run = m_new;
    }
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    fold_o();
    const float inv_l = 1.0f / l_run;
    if (q0 + row < p.Nq) {
      __half* dst = p.out + (static_cast<size_t>(batch) * p.Nq + q0 + row) * p.ldo + head * HD;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        uint4 o;
        o.x = pack_half2(o_acc[8 * i + 0] * inv_l, o_acc[8 * i + 1] * inv_l);
        o.y = pack_half2(o_acc[8 * i + 2] * inv_l, o_acc[8 * i + 3] * inv_l);
        o.z = pack_half2(o_acc[8 * i + 4] * inv_l, o_acc[8 * i + 5] * inv_l);
        o.w = pack_half2(o_acc[8 * i + 6] * inv_l, o_acc[8 * i + 7] * inv_l);
        reinterpret_cast<uint4*>(dst)[i] = o;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

CUtensorMap make_head_map(const __half* base, int ld, int B, int N, int cols) {
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)N, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)N * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_f16(base, 3, dims, strides, box);
}

}  // namespace

AttnOp make_attn_op(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, __half* out,
                    int ldo, int B, int H, int Nq, int Nkv) {
  AttnOp op{};
  CFGPP_REQUIRE(Nkv >= 1 && Nq >= 1, "empty attention");
  CFGPP_REQUIRE(ldo % 8 == 0, "ldo must be a multiple of 8");
  op.p.B = B; op.p.H = H; op.p.Nq = Nq; op.p.Nkv = Nkv; op.p.ldo = ldo; op.p.out = out;
  op.p.scale_log2e = 0.125f * 1.4426950408889634f;
  op.map_q = make_head_map(q, ldq, B, Nq, H * HD);
  op.map_k = make_head_map(k, ldk, B, Nkv, H * HD);
  op.map_v = make_head_map(v, ldv, B, Nkv, H * HD);
  return op;
}

void attn_configure() {
  static bool done = false;
  if (done) return;
  CFGPP_CHECK_CUDA(cudaFuncSetAttribute(attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  done = true;
}

void run_attn_op(const AttnOp& op, cudaStream_t stream) {
  attn_configure();
  dim3 grid((op.p.Nq + BQ - 1) / BQ, op.p.H, op.p.B);
  attn_kernel<<<grid, kThreads, SMEM_BYTES, stream>>>(op.p, op.map_q, op.map_k, op.map_v);
  CFGPP_CHECK_CUDA(cudaGetLastError());
}

}  // namespace cfgpp
