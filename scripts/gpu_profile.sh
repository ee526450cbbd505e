# One-GPU evidence run (round 2): launch list + ncu --set full captures of every wavefront kernel (bench.py workload, C2).
# Each ncu pass only after the same command has exited 0 without ncu.  Reports are summarised ON THE BOX (gpurun copies at
# most 64 MiB back): profiles/summarize_ncu.py -> *_summary.txt, `ncu --page details` -> *_details.txt.
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strong-c3 --e2e-steps 1"
$CMD > gpurun_out/plain.log 2> gpurun_out/plain.err || { echo "plain run failed"; tail -5 gpurun_out/plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_ncu_launches_c2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
cap() {   # name, kernel regex, skip, extra bench args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_$1 -f $CMD $4 > gpurun_out/ncu_$1.log 2>&1
  python profiles/summarize_ncu.py gpurun_out/prof_$1.ncu-rep > gpurun_out/r02_ncu_$1_summary.txt 2>&1
  ncu -i gpurun_out/prof_$1.ncu-rep --page details > gpurun_out/r02_ncu_$1_details.txt 2>&1
  rm -f gpurun_out/prof_$1.ncu-rep gpurun_out/ncu_$1.log
}
# the first cull launches of a render cull camera rays (common-origin form, rays generated in the kernel); the later ones paths in flight
cap wf_cull_common wf_cull 0
cap wf_cull wf_cull 2
cap wf_refine wf_refine 2
cap wf_tiebreak wf_tiebreak 2
cap wf_shade wf_shade 2
cap wf_tail wf_tail 0
cap wf_bvh wf_bvh 2 "--accel bvh"
ls -la gpurun_out/
