#!/bin/bash
# GPU round U of r2: k_tile_finish writes the rows itself (item counts from the read-out, scan first) -- whole suite, cfg4 timing
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/ru_tests.log 2>&1
echo "tests rc=$?" >> $O/ru_tests.log; tail -4 $O/ru_tests.log
for p in 2 1; do LGMI_TILE_PATH=$p timeout 600 python tools/time_cfg4.py 6000 > $O/ru_cfg4_path$p.json 2> $O/ru_cfg4_path$p.err; cat $O/ru_cfg4_path$p.json; done
LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 20000 > $O/ru_cfg4_full.json 2> $O/ru_cfg4_full.err; cat $O/ru_cfg4_full.json
LGMI_TILE_PATH=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ru_launches_cfg4.csv python tools/time_cfg4.py 6000 > $O/ru_ncu4.log 2>&1
echo done
