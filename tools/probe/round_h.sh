#!/bin/bash
# GPU round H of r2: the four-block form of the deep-unit path -- parity, cfg3 timing both forms, launch list
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or cfg3 or deep_unit or all_paths" > $O/rh_tests.log 2>&1
echo "tests rc=$?" >> $O/rh_tests.log; tail -15 $O/rh_tests.log
timeout 600 python bench.py --dense-only --steps 5 --warmup 2 > $O/rh_dense4.json 2> $O/rh_dense4.err; cut -c1-900 $O/rh_dense4.json; tail -3 $O/rh_dense4.err
LGMI_DENSE_PATH=9 timeout 600 python bench.py --dense-only --steps 5 --warmup 2 > $O/rh_dense9.json 2> $O/rh_dense9.err; cut -c1-600 $O/rh_dense9.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rh_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > $O/rh_ncu.log 2>&1
echo done
