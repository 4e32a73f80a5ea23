// cfgpp_b200 — persistent flash-style self-attention for head dim 64 on tcgen05 / TMEM (sm_100a). See attention.cuh.
//
// attention.cu launches one CTA per pair of query tiles: at the SDXL shapes that is 320 CTAs (N = 1024, 20 heads, UNet
// batch 4) on 148 SMs = 2.16 waves, i.e. three rounds of which the last is 16 % full, and every CTA pays its own
// prologue / epilogue (~3 us of a ~17 us life). Here the grid is ONE CTA per SM and each CTA walks a contiguous, evenly
// sized range of the B x H x (Nq / 128) query tiles with two independent pipelines ("slots"):
//   warp 0 / warp 3 : TMA producers of slot 0 / 1   (Q of the slot's next tile; its own K / V ring, 128B swizzle)
//   warp 1 lane 0   : MMA issuer for both slots, event driven:  S_s = Q K_j^T  (M128 N128 K64)  as soon as the
//                     softmax warps of slot s hold S_s(j-1) in registers;  O_s += P_s V_j  (M128 N64 K128) as soon as
//                     P_s(j) is in shared memory — also ACROSS tiles: the next tile's Q / K arrive and its first QK^T is
//                     issued under the current tile's last softmax, its first PV waits only for the O write-out
//   warp 2          : TMEM allocator  (S0 [0,128) S1 [128,256) O0 [256,320) O1 [320,384))
//   warps 4..7 / 8..11 : softmax + output of slot 0 / 1, one query row per thread (same arithmetic as attention.cu:
//                     exp2 with folded scale, fp32 row sums, lazy running-max rescale through tcgen05.ld / st)
// Slot s takes tiles t0 + s, t0 + s + 2, ... of the CTA's range. The two slots no longer share K / V tiles (adjacent
// tiles usually belong to the same head, but a range may straddle heads): K / V come out of L2 twice, ~4.6 TB/s
// aggregate at the N = 1024 shape — inside the L2 -> SM fabric budget — in exchange for a makespan of ceil(T / SMs)
// tiles instead of 2 x ceil(T / 2 / SMs).
#include <cmath>

#include "attention.cuh"
#include "common.cuh"

namespace cfgpp {

namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int HD = 64;
constexpr int KS = 2;                      // K / V ring depth per slot
constexpr int TILE_BYTES = 128 * 64 * 2;   // one [128 x 64] fp16 tile = one 128B-swizzle atom column
constexpr int P_BYTES = 2 * TILE_BYTES;    // P tile: 128 x 128 fp16 as two 64-column halves
constexpr int kThreads = 384;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t O_COL = 256;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
constexpr int SMEM_BYTES = 2 * (TILE_BYTES * (1 + 2 * KS) + P_BYTES) + 1024 + 512;
static_assert(SMEM_BYTES <= 232448, "shared memory overflow");

struct TileCoord {
  int batch, head, q0;
};

__global__ void __launch_bounds__(kThreads, 1)
attn_persist_kernel(const AttnParams p, const __grid_constant__ CUtensorMap map_q,
                    const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_v) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // per slot: Q | K ring | V ring | P
  constexpr int SLOT_BYTES = TILE_BYTES * (1 + 2 * KS) + P_BYTES;
  auto sQ = [&](int s) { return smem + s * SLOT_BYTES; };
  auto sK = [&](int s, int st) { return smem + s * SLOT_BYTES + TILE_BYTES * (1 + st); };
  auto sV = [&](int s, int st) { return smem + s * SLOT_BYTES + TILE_BYTES * (1 + KS + st); };
  auto sP = [&](int s) { return smem + s * SLOT_BYTES + TILE_BYTES * (1 + 2 * KS); };
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * SLOT_BYTES);
  // per slot: q_full, q_empty, s_full, s_free, p_full, pv_done, o_free, k_full[KS], k_empty[KS], v_full[KS], v_empty[KS]
  constexpr int BARS_PER_SLOT = 7 + 4 * KS;
  auto bar = [&](int s, int i) { return bars + s * BARS_PER_SLOT + i; };
  enum { Q_FULL = 0, Q_EMPTY, S_FULL, S_FREE, P_FULL, PV_DONE, O_FREE, K_FULL, K_EMPTY = K_FULL + KS,
         V_FULL = K_EMPTY + KS, V_EMPTY = V_FULL + KS };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * BARS_PER_SLOT);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles = (p.Nkv + BKV - 1) / BKV;
  const int qt_per_head = (p.Nq + BQ - 1) / BQ;
  const int total = p.B * p.H * qt_per_head;
  const int t_begin = static_cast<int>(static_cast<long>(total) * blockIdx.x / gridDim.x);
  const int t_end = static_cast<int>(static_cast<long>(total) * (blockIdx.x + 1) / gridDim.x);
  auto n_items = [&](int s) { return (t_end - t_begin - s + 1) / 2; };  // tiles t_begin + s, + 2, ...
  auto coord = [&](int s, int it) {
    const int t = t_begin + s + 2 * it;
    TileCoord c;
    c.batch = t / (p.H * qt_per_head);
    const int r = t - c.batch * (p.H * qt_per_head);
    c.head = r / qt_per_head;
    c.q0 = (r - c.head * qt_per_head) * BQ;
    return c;
  };

  if ((warp_idx == 0 || warp_idx == 3) && lane == 0) {
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(bar(s, Q_FULL), 1);
      mbar_init(bar(s, Q_EMPTY), 1);
      mbar_init(bar(s, S_FULL), 1);
      mbar_init(bar(s, S_FREE), 128);
      mbar_init(bar(s, P_FULL), 128);
      mbar_init(bar(s, PV_DONE), 1);
      mbar_init(bar(s, O_FREE), 128);
      for (int i = 0; i < KS; ++i) {
        mbar_init(bar(s, K_FULL + i), 1);
        mbar_init(bar(s, K_EMPTY + i), 1);
        mbar_init(bar(s, V_FULL + i), 1);
        mbar_init(bar(s, V_EMPTY + i), 1);
      }
    }
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
  pdl_launch_dependents();
  pdl_wait();

  if (warp_idx == 0 || warp_idx == 3) {
    if (lane == 0) {
      // ===================== TMA producer of one slot =====================
      const int s = warp_idx == 0 ? 0 : 1;
      const int items = n_items(s);
      // K runs one tile ahead of V (K0, K1 V0, K2 V1, ...): a V stage is only recycled when PV(g-2) has retired,
      // late in step g-1; waiting for it BEFORE issuing K(g+1) would hand the tensor pipe its next K tile late.
      const int total_g = items * n_tiles;
      auto tile_of = [&](int g, TileCoord& c, int& j) {
        const int it = g / n_tiles;
        j = g - it * n_tiles;
        c = coord(s, it);
        return it;
      };
      for (int i = 0; i <= total_g; ++i) {
        if (i < total_g) {
          TileCoord c;
          int j;
          const int it = tile_of(i, c, j);
          if (j == 0) {
            mbar_wait(bar(s, Q_EMPTY), (it & 1) ^ 1);  // every QK^T of the previous tile has retired
            mbar_arrive_expect_tx(bar(s, Q_FULL), TILE_BYTES);
            tma_load_3d(sQ(s), &map_q, bar(s, Q_FULL), c.head * HD, c.q0, c.batch);
          }
          const int st = i % KS;
          mbar_wait(bar(s, K_EMPTY + st), ((i / KS) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(s, K_FULL + st), TILE_BYTES);
          tma_load_3d(sK(s, st), &map_k, bar(s, K_FULL + st), c.head * HD, j * BKV, c.batch);
        }
        if (i >= 1) {
          const int g = i - 1;
          TileCoord c;
          int j;
          tile_of(g, c, j);
          const int st = g % KS;
          mbar_wait(bar(s, V_EMPTY + st), ((g / KS) & 1) ^ 1);
          mbar_arrive_expect_tx(bar(s, V_FULL + st), TILE_BYTES);
          tma_load_3d(sV(s, st), &map_v, bar(s, V_FULL + st), c.head * HD, j * BKV, c.batch);
        }
      }
    }
  } else if (warp_idx == 1) {
    if (lane == 0) {
      // ===================== MMA issuer (both slots, event driven) =====================
      constexpr uint32_t idesc_qk = make_idesc_f16(128, BKV, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, HD, 0, 1);  // B (= V) is MN-major
      int total_g[2], g_qk[2] = {0, 0}, g_pv[2] = {0, 0};
      for (int s = 0; s < 2; ++s) total_g[s] = n_items(s) * n_tiles;
      int remaining = 2 * (total_g[0] + total_g[1]);
      long long t_start = clock64();
      while (remaining > 0) {
        bool progressed = false;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          // ---- O_s += P_s(g) V(g) ----
          int g = g_pv[s];
          if (g < total_g[s]) {
            const int j = g % n_tiles, it = g / n_tiles, st = g % KS;
            if (mbar_try_wait(bar(s, P_FULL), g & 1) && mbar_try_wait(bar(s, V_FULL + st), (g / KS) & 1) &&
                (j != 0 || it == 0 || mbar_try_wait(bar(s, O_FREE), (it - 1) & 1))) {
              tc_fence_after();
              const uint64_t v_desc = make_sdesc_sw128(smem_u32(sV(s, st)), 1024, TILE_BYTES);
#pragma unroll
              for (int k = 0; k < BKV / 16; ++k) {
                const uint64_t p_desc =
                    make_sdesc_sw128(smem_u32(sP(s) + (k >> 2) * TILE_BYTES), 1024, 0) + 2 * (k & 3);
                umma_f16(tmem_base + O_COL + s * HD, p_desc, v_desc + 128 * k, idesc_pv, (j | k) != 0 ? 1u : 0u);
              }
              umma_commit(bar(s, PV_DONE));
              umma_commit(bar(s, V_EMPTY + st));
              ++g_pv[s];
              --remaining;
              progressed = true;
            }
          }
          // ---- S_s = Q K(g)^T ----
          g = g_qk[s];
          if (g < total_g[s]) {
            const int j = g % n_tiles, it = g / n_tiles, st = g % KS;
            if ((g == 0 || mbar_try_wait(bar(s, S_FREE), (g - 1) & 1)) &&
                mbar_try_wait(bar(s, K_FULL + st), (g / KS) & 1) &&
                (j != 0 || mbar_try_wait(bar(s, Q_FULL), it & 1))) {
              tc_fence_after();
              const uint64_t q_desc = make_sdesc_sw128(smem_u32(sQ(s)), 1024, 0);
              const uint64_t k_desc = make_sdesc_sw128(smem_u32(sK(s, st)), 1024, 0);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base + s * BKV, q_desc + 2 * k, k_desc + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
              umma_commit(bar(s, S_FULL));
              umma_commit(bar(s, K_EMPTY + st));
              if (j == n_tiles - 1) umma_commit(bar(s, Q_EMPTY));  // the slot's Q buffer may take the next tile
              ++g_qk[s];
              --remaining;
              progressed = true;
            }
          }
        }
        if (progressed) {
          t_start = clock64();
        } else if (clock64() - t_start > 4000000000LL) {
          printf("cfgpp: persistent attention MMA issuer stalled (block %d)\n", blockIdx.x);
          __trap();
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== softmax / output of one slot =====================
    const int s = (warp_idx - 4) >> 2;
    const int qw = warp_idx & 3;  // TMEM lane quarter of this warp
    const int row = qw * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(qw * 32) << 16;
    const uint32_t s_addr = tmem_base + s * BKV + lane_off;
    const uint32_t o_addr = tmem_base + O_COL + s * HD + lane_off;
    uint8_t* prow = sP(s) + row * 128;
    const float c = p.scale_log2e;
    const int items = n_items(s);
    int g = 0;
    for (int it = 0; it < items; ++it) {
      const TileCoord tc = coord(s, it);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < n_tiles; ++j, ++g) {
        const int valid = p.Nkv - j * BKV;  // columns >= valid are padding (last tile only)
        mbar_wait(bar(s, S_FULL), g & 1);
        tc_fence_after();
        uint32_t sc[128];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) tmem_ld_x32(s_addr + q4 * 32, *reinterpret_cast<uint32_t(*)[32]>(&sc[q4 * 32]));
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar(s, S_FREE));  // scores are in registers: the tensor pipe may already produce the next S
        if (valid < BKV) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= valid) sc[i] = 0xff800000u;  // -inf
        }
        float mx0 = __uint_as_float(sc[0]), mx1 = __uint_as_float(sc[1]);
#pragma unroll
        for (int i = 2; i < 128; i += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sc[i]));
          mx1 = fmaxf(mx1, __uint_as_float(sc[i + 1]));
        }
        const float mx = fmaxf(mx0, mx1);
        // lazy running max: move the reference only when it would otherwise leave the comfortable range
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          m_ref = mx;
        } else if ((mx - m_ref) * c > kRescaleThreshold) {
          alpha = fast_exp2((m_ref - mx) * c);
          m_ref = mx;
          need = true;
        }
        if (g > 0) {
          // PV(g-1) retired: the P buffer is reusable (also across tiles) and, within a tile, O is rescalable
          mbar_wait(bar(s, PV_DONE), (g - 1) & 1);
          tc_fence_after();
          if (j > 0 && __any_sync(0xffffffffu, need)) {
#pragma unroll 1
            for (int h = 0; h < HD / 16; ++h) {
              uint32_t o[16];
              tmem_ld_x16(o_addr + h * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int d = 0; d < 16; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) * alpha);
              tmem_st_x16(o_addr + h * 16, o);
            }
            tmem_st_wait();
            l_run *= alpha;
          }
        }
        const float mc = m_ref * c;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int gq = 0; gq < 16; ++gq) {  // 16-byte chunks of the 256-byte P row (two 128-byte halves)
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float p0 = fast_exp2(__uint_as_float(sc[gq * 8 + 2 * e]) * c - mc);
            const float p1 = fast_exp2(__uint_as_float(sc[gq * 8 + 2 * e + 1]) * c - mc);
            rs0 += p0;
            rs1 += p1;
            pk[e] = pack_half2(p0, p1);
          }
          const int half_idx = gq >> 3;        // which 64-column half
          const int ch = (gq & 7) ^ (row & 7);  // 128B swizzle
          *reinterpret_cast<uint4*>(prow + half_idx * TILE_BYTES + ch * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
        l_run += rs0 + rs1;
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar(s, P_FULL));
      }
      // ---- tile done: O / l -> fp16 -> global; then the MMA issuer may start the next tile's first PV ----
      mbar_wait(bar(s, PV_DONE), (g - 1) & 1);
      tc_fence_after();
      const float inv_l = 1.0f / l_run;
      const int qrow = tc.q0 + row;
      __half* dst = p.out + (static_cast<size_t>(tc.batch) * p.Nq + qrow) * p.ldo + tc.head * HD;
#pragma unroll 1
      for (int h = 0; h < HD / 32; ++h) {
        uint32_t o[32];
        tmem_ld_x32(o_addr + h * 32, o);
        tmem_ld_wait();
        if (qrow < p.Nq) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            uint4 w;
            w.x = pack_half2(__uint_as_float(o[8 * i + 0]) * inv_l, __uint_as_float(o[8 * i + 1]) * inv_l);
            w.y = pack_half2(__uint_as_float(o[8 * i + 2]) * inv_l, __uint_as_float(o[8 * i + 3]) * inv_l);
            w.z = pack_half2(__uint_as_float(o[8 * i + 4]) * inv_l, __uint_as_float(o[8 * i + 5]) * inv_l);
            w.w = pack_half2(__uint_as_float(o[8 * i + 6]) * inv_l, __uint_as_float(o[8 * i + 7]) * inv_l);
            reinterpret_cast<uint4*>(dst + h * 32)[i] = w;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar(s, O_FREE));
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

void attn_persist_configure() {
  CFGPP_CHECK_CUDA(cudaFuncSetAttribute(attn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
}

// Worth it when there are more query tiles than the pair-per-CTA grid can balance: at least ~2 tiles per SM.
bool attn_persist_applicable(const AttnOp& op) {
  if (op.hd_pad != 64 || op.p.Nkv <= 128) return false;
  const int total = op.p.B * op.p.H * ((op.p.Nq + BQ - 1) / BQ);
  return total >= 2 * num_sms();
}

void run_attn_persist_op(const AttnOp& op, cudaStream_t stream) {
  const int total = op.p.B * op.p.H * ((op.p.Nq + BQ - 1) / BQ);
  const int grid = total < num_sms() ? total : num_sms();
  launch_pdl(attn_persist_kernel, dim3(grid), dim3(kThreads), SMEM_BYTES, stream, op.p, op.map_q, op.map_k, op.map_v);
}

}  // namespace cfgpp
