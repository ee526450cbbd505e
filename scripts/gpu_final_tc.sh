# final round-2 evidence with the tensor-core cull: launch list, ncu captures of wf_cull_tc, timeline, make-final record
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c3 --e2e-steps 1"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_ncu_launches_c2_tc.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cuobjdump -xelf all raytrace_clj_b200/libraytrace_b200.so > /dev/null 2>&1; CUBIN=$(ls *.cubin | head -1)
for skip in 0 3; do
  ncu --set full --clock-control none --import-source on -k regex:wf_cull_tc -s $skip -c 1 -o gpurun_out/prof_tc_$skip -f $CMD > gpurun_out/ncu_tc_$skip.log 2>&1
  python profiles/summarize_ncu.py gpurun_out/prof_tc_$skip.ncu-rep > gpurun_out/r02_ncu_wf_cull_tc_s${skip}_summary.txt 2>&1
  ncu -i gpurun_out/prof_tc_$skip.ncu-rep --page details > gpurun_out/r02_ncu_wf_cull_tc_s${skip}_details.txt 2>&1
  ncu -i gpurun_out/prof_tc_$skip.ncu-rep --page source --csv > gpurun_out/tc_source_$skip.csv 2>&1
  python profiles/sass_by_line.py gpurun_out/tc_source_$skip.csv $CUBIN _ZN2rt10wf_cull_tcENS_10WaveParamsE 30 > gpurun_out/r02_ncu_wf_cull_tc_s${skip}_by_line.txt 2>&1
  rm -f gpurun_out/prof_tc_$skip.ncu-rep gpurun_out/tc_source_$skip.csv
done
rm -f *.cubin
timeout 100 python scripts/tc_trace.py gpurun_out/r02_timeline_c2_tc_two_lanes.txt 2 1
python bench.py --workload final --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_final_tc.json 2> gpurun_out/final.err; tail -2 gpurun_out/final.err
python - <<P
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_final_tc.json") if l.startswith("{")][-1])
print("final", d["ms_per_step"], d["value"], d["roofline"]["frac"])
P
grep -E "duration|issue slots|ALU pipe|executed warp" gpurun_out/r02_ncu_wf_cull_tc_s3_summary.txt
head -12 gpurun_out/r02_ncu_wf_cull_tc_s3_by_line.txt
