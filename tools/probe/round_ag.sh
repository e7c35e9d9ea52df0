#!/bin/bash
# GPU round AG of r2: k_tile_finish pass 2 with the next listed pair's counts prefetched
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py -x -q -m gpu -k "heavy_tail or many_sites or mid_units or cfg4 or all_paths" > $O/rag_tests.log 2>&1
echo "tests rc=$?" >> $O/rag_tests.log; tail -3 $O/rag_tests.log
timeout 300 python tools/time_cfg4.py 6000 > $O/rag_cfg4.log 2>&1; tail -3 $O/rag_cfg4.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rag_launches_cfg4.csv python tools/time_cfg4.py 6000 > /dev/null 2>&1
echo done
