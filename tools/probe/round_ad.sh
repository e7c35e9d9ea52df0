#!/bin/bash
# GPU round AD of r2: per-site-sum CTAs of the largest units first -- mid-depth / heavy-tail parity, cfg4 timing
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py tests/test_multi_gpu.py -x -q -m gpu -k "heavy_tail or many_sites or mid_units or cfg4 or all_paths or pipelin or dense" > $O/rad_tests.log 2>&1
echo "tests rc=$?" >> $O/rad_tests.log; tail -3 $O/rad_tests.log
for i in 1 2; do timeout 600 python tools/time_cfg4.py 6000 2>/dev/null | cut -c1-60,210-300; done
timeout 600 python tools/time_cfg4.py 20000 2>/dev/null | cut -c1-60,210-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rad_launches_cfg4.csv python tools/time_cfg4.py 6000 > /dev/null 2>&1
echo done
