mkdir -p gpurun_out
while read -r kb ctas wl steps; do
  RT_LIGHT_SMEM_KB=$kb RT_CULL_CTAS_PER_SM=$ctas timeout 300 python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/bench_s.json 2> gpurun_out/bench_s.err || tail -3 gpurun_out/bench_s.err
  python - $kb $ctas $wl <<'PY'
import json,sys
d=json.load(open("gpurun_out/bench_s.json"))
print("light smem KB / cull ctas / workload",*sys.argv[1:], "ms %.3f frac %.4f"%(d["ms_per_step"], d["roofline"]["frac"]), "e2e %.1fM"%(d["e2e"]["value"]/1e6), flush=True)
PY
done <<'CFG'
0 4 c2 20
20 4 c2 20
30 4 c2 20
40 4 c2 20
48 4 c2 20
0 4 c3-slice 4
30 4 c3-slice 4
40 4 c3-slice 4
48 4 c3-slice 4
CFG
