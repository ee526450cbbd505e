# Final round-2 state: whole GPU suite, bench records (C2 default run, C1, C4), launch list, ncu capture of wf_tail, timeline
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02_final_pytest_gpu.log
run() { name=$1; shift; python bench.py "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err || { echo "$name FAILED"; tail -3 gpurun_out/$name.err; }; python - <<P
import json
try:
    d=json.loads([l for l in open("gpurun_out/$name.json") if l.startswith("{")][-1])
    print("$name", round(d["ms_per_step"],3), "ms  frac", round(d["roofline"]["frac"],4), " e2e ms", round(d["e2e"]["ms_per_step"],3), " dom", d["roofline"]["dominant_kernel"] and round(d["roofline"]["dominant_kernel"]["frac"],3), " launches", d.get("gpu_launches"))
except Exception as e: print("$name", "parse error", e)
P
}
run r02_final_bench_c2 --steps 30 --warmup 5
run r02_final_bench_c1 --workload c1 --steps 30 --warmup 5 --no-cpu-baseline
run r02_final_bench_c4 --workload c4 --steps 10 --warmup 3 --no-cpu-baseline
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c3 --e2e-steps 1"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_final_ncu_launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_tail -s 0 -c 1 -o gpurun_out/prof_wf_tail -f $CMD > gpurun_out/ncu_wf_tail.log 2>&1
python profiles/summarize_ncu.py gpurun_out/prof_wf_tail.ncu-rep > gpurun_out/r02_final_ncu_wf_tail_summary.txt 2>&1
ncu -i gpurun_out/prof_wf_tail.ncu-rep --page details > gpurun_out/r02_final_ncu_wf_tail_details.txt 2>&1
rm -f gpurun_out/prof_wf_tail.ncu-rep
timeout 100 python scripts/tc_trace.py gpurun_out/r02_final_timeline_c2.txt 2 1
python profiles/launch_summary.py gpurun_out/r02_final_ncu_launches_c2.csv 2>/dev/null | tail -12
head -12 gpurun_out/r02_final_ncu_wf_tail_summary.txt
