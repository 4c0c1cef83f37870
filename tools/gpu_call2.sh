#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/ap_variants.py > $O/ap_variants2.jsonl 2> $O/ap_variants2.err
echo "ap_variants exit $?" > $O/status2.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "allpairs or degenerate or clustering or matrix" > $O/pytest_allpairs2.log 2>&1
echo "pytest allpairs exit $?" >> $O/status2.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:allpairs_tc144_kernel -c 1 -f -o $O/prof_allpairs144 \
    python bench.py --workload allpairs --steps 1 --warmup 3 --no-e2e --no-cpu > $O/prof_allpairs144.log 2>&1
echo "ncu exit $?" >> $O/status2.txt
cat $O/status2.txt; tail -3 $O/pytest_allpairs2.log; cut -c1-600 $O/ap_variants2.jsonl; ls -la $O
