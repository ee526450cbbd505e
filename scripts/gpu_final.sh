# Final record of a round: GPU tests, smoke, the default bench (both arms), the ncu launch list of the same command.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_c2_reference.json 2> gpurun_out/bench_c2_reference.err; echo "reference arm exit $?"
if [ "$1" = "launches" ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
fi
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2.json")); print(d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["dominant_kernel"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"), d["gpu_launches"], d["clocks"])
PY
