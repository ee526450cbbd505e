mkdir -p gpurun_out
while read -r cap tail; do
  RT_WAVE_CAPACITY=$cap RT_TAIL_ENTRIES=$tail timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err
  python - $cap $tail <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json")); k=d["roofline"]["dominant_kernel"]
print("capacity/tail_entries",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "cull %.2f (%.3f)"%(k["ms_per_step"],k["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), "launches", d["gpu_launches"], flush=True)
PY
done <<'CFG'
4194304 65536
2097152 65536
3145728 65536
6291456 65536
4194304 16384
4194304 262144
4194304 1048576
3145728 262144
CFG
