#!/bin/bash
# Same-box A/B of the round-2 switches on the headline workload (one gpurun call): prints value / e2e / per-class ms.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "CFGPP_AB=default" "CFGPP_NO_XATTN=1" "CFGPP_NO_STREAMK=1" "CFGPP_SPLIT=1" "CFGPP_ATTN_ROWSUM_MMA=1"; do
  env $v timeout 400 python bench.py --steps 3 --warmup 3 --no-baselines > gpurun_out/bench_ab.json 2> gpurun_out/bench_ab.err
  python - "$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/bench_ab.json"))
r = d["roofline"]
print(f"{sys.argv[1]:28s} value {d['value']:.4f} e2e {d['e2e']['value']:.4f} ms/traj {d['ms_per_step']:.1f} "
      f"by_kind_ms {{{', '.join(f'{k}: {v:.2f}' for k, v in r['by_kind_ms'].items())}}} clocks {d['clocks']['sm_mhz']}")
PY
done
