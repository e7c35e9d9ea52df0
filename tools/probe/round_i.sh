#!/bin/bash
# GPU round I of r2: deep-unit path after the latency work (prep + lists fused, fix-up loads in flight, deep-unit pair
# kernel, tight per-site sums) -- the whole -m gpu suite, cfg3 and cfg4 timing, launch lists
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/ri_tests.log 2>&1
echo "tests rc=$?" >> $O/ri_tests.log; tail -6 $O/ri_tests.log
timeout 600 python bench.py --dense-only --steps 5 --warmup 2 > $O/ri_dense4.json 2> $O/ri_dense4.err; cut -c1-700 $O/ri_dense4.json; tail -3 $O/ri_dense4.err
LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 6000 > $O/ri_cfg4_path2.json 2> $O/ri_cfg4_path2.err; cat $O/ri_cfg4_path2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ri_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > $O/ri_ncu.log 2>&1
LGMI_TILE_PATH=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ri_launches_cfg4.csv python tools/time_cfg4.py 6000 > $O/ri_ncu4.log 2>&1
echo done
