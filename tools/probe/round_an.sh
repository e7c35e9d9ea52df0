#!/bin/bash
# GPU round AN of r2: ncu capture of k_pairs_generic<1> (ordering / emission of the mid-depth units) on the cfg4 sample
O=gpurun_out
mkdir -p $O
timeout 300 python tools/time_cfg4.py 6000 > $O/ran_cfg4.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_generic -s 2 -c 1 -o $O/ran_prof_k_pairs_generic1 python tools/time_cfg4.py 6000 > $O/ran_ncu.log 2>&1; tail -1 $O/ran_ncu.log
echo done
