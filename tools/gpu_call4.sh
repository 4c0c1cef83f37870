#!/bin/bash
set -u
O=gpurun_out
mkdir -p $O
timeout 300 python tools/ap_variants.py > $O/ap_variants4.jsonl 2> $O/ap_variants4.err
echo "ap_variants exit $?" > $O/status4.txt
timeout 200 python tools/ap_operand_check.py > $O/ap_operand_check4.jsonl 2>&1
timeout 300 python -m pytest tests -m gpu -q --timeout 200 -k "allpairs or degenerate or clustering or matrix" > $O/pytest_allpairs4.log 2>&1
echo "pytest allpairs exit $?" >> $O/status4.txt
cat $O/status4.txt; grep -E "^E  .*(AssertionError|max abs)|passed|failed" $O/pytest_allpairs4.log | head; cut -c1-1400 $O/ap_variants4.jsonl | tail -2; cat $O/ap_operand_check4.jsonl | tail -2
