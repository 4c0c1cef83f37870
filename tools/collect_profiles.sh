#!/bin/bash
# Round-end measurement pass (run on the GPU box from the repo root, one GPU): the bench lines, launch lists, one full ncu
# capture of the dominant kernel per workload and the size sweeps, all into gpurun_out/ under the round's prefix;
# tools/ncu_select.py then cuts the captures down and refreshes profiles/ncu_traffic.json.  Copy what should be judged
# into profiles/.     tools/collect_profiles.sh [prefix, default r02]
set -u
R=${1:-r02}
O=gpurun_out
mkdir -p $O
timeout 900 python bench.py 2>$O/${R}_bench_default.err | tail -1 > $O/${R}_bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>$O/${R}_ref_arm.err | tail -1 > $O/${R}_ref_arm.json
timeout 200 python bench.py --only ala2 --no-subs 2>/dev/null | tail -1 > $O/${R}_bench_ala2.json
for w in ovm superpose allpairs_20k; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${R}_${w}_launches.csv \
      python bench.py --only $w --steps 2 --warmup 3 --no-e2e --no-subs > $O/${R}_${w}_launches.log 2>&1
done
for spec in ovm:ovm_tma_kernel 'superpose:frame_resident_kernel|superpose_pipe_kernel' allpairs_20k:allpairs_tc144_kernel; do
  w=${spec%%:*}; k=${spec#*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o $O/${R}_prof_$w \
      python bench.py --only $w --steps 1 --warmup 3 --no-e2e --no-subs > $O/${R}_prof_$w.log 2>&1
  ncu -i $O/${R}_prof_$w.ncu-rep --page raw --csv > $O/${R}_prof_${w}_raw.csv 2>/dev/null
done
timeout 300 python tools/ovm_sweep.py 22 50 100 200 256 300 400 516 600 700 800 1000 2000 5000 2>/dev/null > $O/${R}_ovm_sweep_final.jsonl
SWEEP_QUICK=1 timeout 300 python tools/fused_sweep.py 22 50 100 200 300 500 700 1000 1400 2000 3000 5000 2>/dev/null | grep -v BEST > $O/${R}_fused_final.jsonl
timeout 200 python tools/cluster_time.py 20000 300 2>/dev/null | tail -1 > $O/${R}_cluster_time.json
timeout 200 python tools/ap_time.py 2>/dev/null | tail -5 > $O/${R}_ap_time.log
timeout 200 python tools/ap_operand_check.py 2>/dev/null > $O/${R}_ap_operand_check.jsonl
timeout 200 python tools/lprmsd_time.py 2>/dev/null > $O/${R}_lprmsd_time.jsonl
ls -la $O | tail -40
