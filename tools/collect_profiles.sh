#!/bin/bash
# Round-end measurement pass (run on the GPU box from the repo root): bench lines, launch lists and one full ncu capture
# of the dominant kernel per workload, all into gpurun_out/ (copy what should be judged into profiles/).
set -u
O=gpurun_out
mkdir -p $O
for w in ovm superpose allpairs ovm25k ala2; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 2>$O/bench_$w.err | tail -1 > $O/bench_$w.json
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>$O/bench_ref_arm.err | tail -1 > $O/bench_ref_arm.json
for w in ovm superpose allpairs; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${w}_launches.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu > $O/${w}_launches.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ovm_tma_kernel -c 1 -f -o $O/prof_ovm \
    python bench.py --workload ovm --steps 1 --warmup 3 --no-e2e --no-cpu > $O/prof_ovm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:frame_resident_kernel -c 1 -f -o $O/prof_superpose \
    python bench.py --workload superpose --steps 1 --warmup 3 --no-e2e --no-cpu > $O/prof_superpose.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:allpairs_tc144_kernel -c 1 -f -o $O/prof_allpairs \
    python bench.py --workload allpairs --steps 1 --warmup 3 --no-e2e --no-cpu > $O/prof_allpairs.log 2>&1
for w in ovm superpose allpairs; do
  ncu -i $O/prof_$w.ncu-rep --page raw --csv > $O/prof_${w}_raw.csv 2>/dev/null
done
timeout 200 python tools/quick_time.py 200000 1000 10 2>/dev/null | tail -1 > $O/quick_1000.json
timeout 200 python tools/quick_time.py 40000 5000 10 2>/dev/null | tail -1 > $O/quick_5000.json
timeout 300 python tools/ovm_sweep.py 22 50 100 200 256 300 400 516 600 700 800 1000 2000 5000 2>/dev/null > $O/ovm_sweep_final.jsonl
SWEEP_QUICK=1 timeout 300 python tools/fused_sweep.py 22 50 100 200 300 500 700 1000 1400 2000 3000 5000 2>/dev/null | grep -v BEST > $O/fused_final.jsonl
timeout 200 python tools/cluster_time.py 20000 300 2>/dev/null | tail -1 > $O/cluster_time.json
timeout 200 python tools/ap_time.py 2>/dev/null | tail -5 > $O/ap_time.log
timeout 200 python tools/ap_variants.py 2>/dev/null > $O/ap_variants.jsonl
timeout 200 python tools/ap_operand_check.py 2>/dev/null > $O/ap_operand_check.jsonl
# BASELINE configs[3] at full size on ONE GPU: 100k x 100k x 300 atoms, the 40 GB matrix stays in HBM
timeout 300 python bench.py --workload allpairs --frames 100000 --steps 3 --warmup 3 --no-e2e --no-cpu 2>$O/bench_allpairs_100k.err | tail -1 > $O/bench_allpairs_100k.json
ls -la $O | head -50
