# final round-2 evidence with the tensor-core cull: tests, smoke, records, launch list, ncu captures, timeline
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
bash scripts/gpu_records_tc.sh 2>&1 | tail -8
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c3 --e2e-steps 1"
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
head -12 gpurun_out/r02_ncu_wf_cull_tc_s3_summary.txt
