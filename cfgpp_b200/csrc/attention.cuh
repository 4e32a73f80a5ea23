// cfgpp_b200 — fused scaled-dot-product attention (no mask, no dropout) for head_dim 64 on tcgen05.
//   out[b, i, h*64 + :] = softmax(q_i k^T / sqrt(64)) v       (AttnProcessor2_0 / F.scaled_dot_product_attention)
// q / k / v are strided views into token-major activation buffers ([B*N, ld] fp16, head h at column h*64),
// so the fused QKV GEMM output (self-attention) and the fused KV GEMM output (cross-attention) are consumed
// in place; the output is written token-major [B*Nq, ldo] ready for the to_out GEMM.
#pragma once
#include "host.h"

namespace cfgpp {

struct AttnParams {
  int B, H, Nq, Nkv;
  int ldo;
  __half* out;
  float scale_log2e;  // (1/sqrt(d)) * log2(e)
};

struct AttnOp {
  CUtensorMap map_q, map_k, map_v;
  AttnParams p;
  double flops() const { return 4.0 * p.B * p.H * (double)p.Nq * p.Nkv * 64; }
};

AttnOp make_attn_op(const __half* q, int ldq, const __half* k, int ldk, const __half* v, int ldv, __half* out,
                    int ldo, int B, int H, int Nq, int Nkv);
void run_attn_op(const AttnOp& op, cudaStream_t stream);

}  // namespace cfgpp
