# One-GPU evidence run: tests, bench, ncu launch list of the same command, ncu --set full of each wavefront kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 50 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in wf_cull wf_refine wf_tiebreak wf_shade; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2.json")); print(d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"))
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err; echo "reference arm exit $?"
python bench.py --variant 0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2_mega.json 2> gpurun_out/bench_c2_mega.err; echo "megakernel exit $?"
