for m in fp32; do timeout 600 python bench.py --mode $m --batch 64 --steps 5 --no-cpu-baseline > gpurun_out/bench_r2_$m.json 2> gpurun_out/bench_r2_$m.err; python - <<PY
import json
d=json.load(open("gpurun_out/bench_r2_$m.json")); print("$m", d["ms_per_step"], d["value"], {k:(v["ms"] if isinstance(v,dict) else v) for k,v in d["kernels"].items()})
PY
done
