# One-GPU evidence run: ncu --set full captures of every wavefront kernel (bench.py workload, C2).
mkdir -p gpurun_out
# the first cull launches of a render cull camera rays (common-origin form); the later ones cull paths in flight (general form)
ncu --set full --clock-control none --import-source on -k regex:wf_cull -s 0 -c 1 -o gpurun_out/prof_wf_cull_common -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wf_cull_common.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wf_cull -s 2 -c 1 -o gpurun_out/prof_wf_cull -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wf_cull.log 2>&1
for k in wf_refine wf_tiebreak wf_shade; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:wf_tail -s 0 -c 1 -o gpurun_out/prof_wf_tail -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_wf_tail.log 2>&1
ls -la gpurun_out/*.ncu-rep
