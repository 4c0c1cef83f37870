#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/ap_variants.py > $O/ap_variants5.jsonl 2> $O/ap_variants5.err
echo "ap_variants exit $?" > $O/status5.txt
timeout 600 python -m pytest tests -m gpu -q --timeout 300 > $O/pytest_gpu5.log 2>&1
echo "pytest gpu exit $?" >> $O/status5.txt
cat $O/status5.txt; grep -E "^E  .*(AssertionError|max abs)|passed|failed" $O/pytest_gpu5.log | head; cut -c1-700 $O/ap_variants5.jsonl | tail -2
