# ncu capture of wf_cull_tc (C2, tensor-core cull), summarised on the box: summary, details, per-source-line attribution
mkdir -p gpurun_out
export RT_CULL_TC=1
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c3 --e2e-steps 1"
$CMD > gpurun_out/tc_plain.log 2> gpurun_out/tc_plain.err || { echo "plain run failed"; tail -5 gpurun_out/tc_plain.err; exit 1; }
tail -1 gpurun_out/tc_plain.log | cut -c1-300
cuobjdump -xelf all raytrace_clj_b200/libraytrace_b200.so > /dev/null 2>&1; CUBIN=$(ls *.cubin | head -1)
for skip in ${SKIPS:-1}; do
  ncu --set full --clock-control none --import-source on -k regex:wf_cull_tc -s $skip -c 1 -o gpurun_out/prof_tc_$skip -f $CMD > gpurun_out/ncu_tc_$skip.log 2>&1
  python profiles/summarize_ncu.py gpurun_out/prof_tc_$skip.ncu-rep > gpurun_out/r02_ncu_wf_cull_tc_s${skip}_summary.txt 2>&1
  ncu -i gpurun_out/prof_tc_$skip.ncu-rep --page details > gpurun_out/r02_ncu_wf_cull_tc_s${skip}_details.txt 2>&1
  ncu -i gpurun_out/prof_tc_$skip.ncu-rep --page source --csv > gpurun_out/tc_source_$skip.csv 2>&1
  python profiles/sass_by_line.py gpurun_out/tc_source_$skip.csv $CUBIN _ZN2rt10wf_cull_tcENS_10WaveParamsE 45 > gpurun_out/tc_by_line_$skip.txt 2>&1
  rm -f gpurun_out/prof_tc_$skip.ncu-rep
done
rm -f *.cubin
head -30 gpurun_out/r02_ncu_wf_cull_tc_s${SKIPS:-1}_summary.txt
cat gpurun_out/tc_by_line_${SKIPS:-1}.txt
