#!/bin/bash
for cfg in "SGB200_PDL=2" "SGB200_PDL=0" "SGB200_PDL=2 SGB200_PDL_MAX_CTAS=600" "SGB200_PDL=2" "SGB200_PDL=0" "SGB200_PDL=2 SGB200_PDL_MAX_CTAS=600"; do
  env $cfg timeout 300 python bench.py --no-cpu-baseline --no-profile --steps 30 --warmup 5 > gpurun_out/b_x.json 2>gpurun_out/b_x.err
  python - "$cfg" gpurun_out/b_x.json <<'P'
import json,sys; d=json.loads(open(sys.argv[2]).read()); print(sys.argv[1].ljust(44), round(d['ms_per_step'],3), round(d['value'],3), d['clocks']['sm_mhz'])
P
done
for cfg in "SGB200_PDL=2" "SGB200_PDL=0"; do echo "== latency $cfg"; env $cfg timeout 300 python scripts/latency_sweep.py 2>/dev/null | grep -E "^ +(1|8|64|256) +bf16"; done
