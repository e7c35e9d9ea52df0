#!/bin/bash
O=gpurun_out
mkdir -p $O
for i in 1 2 3; do timeout 120 python tools/sanitize_paths.py > $O/rs_run$i.log 2>&1; echo "run $i rc=$?"; tail -2 $O/rs_run$i.log | cut -c1-200; done
LGMI_GRAPHS=0 timeout 120 python tools/sanitize_paths.py > $O/rs_nograph.log 2>&1; echo "nograph rc=$?"; tail -2 $O/rs_nograph.log | cut -c1-200
LGMI_DENSE_PATH=9 timeout 120 python tools/sanitize_paths.py > $O/rs_nine.log 2>&1; echo "nine rc=$?"; tail -2 $O/rs_nine.log | cut -c1-200
