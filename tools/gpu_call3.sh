#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/ap_variants.py > $O/ap_variants3.jsonl 2> $O/ap_variants3.err
echo "ap_variants exit $?" > $O/status3.txt
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "allpairs or degenerate or clustering or matrix" > $O/pytest_allpairs3.log 2>&1
echo "pytest allpairs exit $?" >> $O/status3.txt
cat $O/status3.txt; tail -5 $O/pytest_allpairs3.log; cut -c1-1200 $O/ap_variants3.jsonl; tail -5 $O/ap_variants3.err
