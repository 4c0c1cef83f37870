#!/bin/bash
# one-shot validation pass for the dense-layout all-pairs kernel and the solver changes (run under gpurun)
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > $O/gpu.txt 2>&1
timeout 300 python tools/ap_variants.py > $O/ap_variants.jsonl 2> $O/ap_variants.err
echo "ap_variants exit $?" >> $O/status.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "allpairs or degenerate or clustering or matrix" > $O/pytest_allpairs.log 2>&1
echo "pytest allpairs exit $?" >> $O/status.txt
timeout 500 python -m pytest tests -m gpu -q --timeout 200 -k "not (allpairs or degenerate or clustering or matrix)" > $O/pytest_rest.log 2>&1
echo "pytest rest exit $?" >> $O/status.txt
timeout 200 python bench.py --workload allpairs --steps 10 --warmup 3 > $O/bench_allpairs.json 2> $O/bench_allpairs.err
echo "bench allpairs exit $?" >> $O/status.txt
cat $O/status.txt; tail -3 $O/pytest_allpairs.log; tail -3 $O/pytest_rest.log; cat $O/ap_variants.jsonl | cut -c1-400
