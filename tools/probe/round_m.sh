#!/bin/bash
# GPU round M of r2: two-sum compensated sums, fix-up beside the GEMM -- whole suite, cfg2/cfg3/cfg4 timing
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rm_tests.log 2>&1
echo "tests rc=$?" >> $O/rm_tests.log; tail -5 $O/rm_tests.log
timeout 600 python bench.py --dense-only --steps 5 --warmup 2 > $O/rm_dense4.json 2> $O/rm_dense4.err; cut -c1-420 $O/rm_dense4.json; tail -3 $O/rm_dense4.err
LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 6000 > $O/rm_cfg4_path2.json 2> $O/rm_cfg4_path2.err; cat $O/rm_cfg4_path2.json
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dense --no-strong --no-cfg5 --no-cfg1 > $O/rm_bench_cfg2.json 2> $O/rm_bench_cfg2.err; cut -c1-300 $O/rm_bench_cfg2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rm_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > $O/rm_ncu.log 2>&1
LGMI_TILE_PATH=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rm_launches_cfg4.csv python tools/time_cfg4.py 6000 > $O/rm_ncu4.log 2>&1
echo done
