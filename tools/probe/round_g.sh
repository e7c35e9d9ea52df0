#!/bin/bash
# GPU round G of r2: direct export of the pipelined step -- parity, then timing by number of groups
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelin or compact_rows or cfg2_full or end_to_end" > $O/rg_tests.log 2>&1
echo "tests rc=$?" >> $O/rg_tests.log; tail -5 $O/rg_tests.log
timeout 600 python tools/e2e_variants.py 4 6 8 12 16 > $O/rg_e2e.txt 2>&1; cat $O/rg_e2e.txt
LGMI_PIPE_DEBUG=1 timeout 300 python tools/e2e_variants.py 8 2>&1 | grep -v "^chunks" | tail -3 > $O/rg_e2e_debug.txt; cat $O/rg_e2e_debug.txt
LGMI_PIPE_DIRECT=0 timeout 600 python tools/e2e_variants.py 4 8 > $O/rg_e2e_copies.txt 2>&1; cat $O/rg_e2e_copies.txt
