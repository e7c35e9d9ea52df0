#!/bin/bash
# GPU round AO of r2: k_pairs_generic<1> over its own list of item records; record and MI values of the next item requested a whole item ahead
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py -x -q -m gpu -k "heavy_tail or many_sites or mid_units or cfg4 or all_paths or pipelined or counts" > $O/rao_tests.log 2>&1
echo "tests rc=$?" >> $O/rao_tests.log; tail -3 $O/rao_tests.log
timeout 300 python tools/time_cfg4.py 6000 > $O/rao_cfg4.log 2>&1; tail -1 $O/rao_cfg4.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rao_launches_cfg4.csv python tools/time_cfg4.py 6000 > /dev/null 2>&1
echo done
