mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_r1b.json 2> gpurun_out/bench_r1b.err; echo "bench exit $?"
RT_WAVE_LANES=1 RT_CULL_CTAS_PER_SM=5 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1b_1lane.json 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:wf_cull -s 6 -c 1 -o gpurun_out/prof_cull_g16 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_cull_g16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_shade -s 6 -c 1 -o gpurun_out/prof_shade_g16 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_shade_g16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_refine -s 6 -c 1 -o gpurun_out/prof_refine_g16 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_refine_g16.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for f in ("bench_r1b","bench_r1b_1lane"):
    d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"], d["e2e"]["value"])
PY
