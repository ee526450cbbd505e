mkdir -p gpurun_out
while read -r cap lanes wl; do
  RT_WAVE_CAPACITY=$cap RT_WAVE_LANES=$lanes timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python - $cap $lanes $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("capacity/lanes/workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], flush=True)
PY
done <<'CFG'
6291456 2 c2
8388608 2 c2
10485760 2 c2
12582912 2 c2
16777216 2 c2
8388608 1 c2
8388608 3 c2
4194304 2 c4
8388608 2 c4
16777216 2 c4
4194304 2 c3-slice
8388608 2 c3-slice
16777216 2 c3-slice
CFG
