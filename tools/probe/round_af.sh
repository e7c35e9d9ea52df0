#!/bin/bash
# GPU round AF of r2: k_other_fix at four CTAs per SM, k_tile_finish with the next pair's counts prefetched
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py -x -q -m gpu -k "heavy_tail or many_sites or mid_units or cfg4 or all_paths or dense or cfg3" > $O/raf_tests.log 2>&1
echo "tests rc=$?" >> $O/raf_tests.log; tail -3 $O/raf_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/raf_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/raf_launches_cfg4.csv python tools/time_cfg4.py 6000 > /dev/null 2>&1
echo done
