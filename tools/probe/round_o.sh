#!/bin/bash
# GPU round O of r2: deep-unit helpers after the second pass (fix-up by label-owning warps, coalesced X stores, one
# barrier per tile in the per-site sums) -- parity, cfg3 / cfg4 timing, launch lists, ncu of the deep-unit pair kernel
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py -x -q -m gpu -k "dense or cfg3 or deep_unit or all_paths or heavy_tail or many_sites or mid_units or cfg4" > $O/ro_tests.log 2>&1
echo "tests rc=$?" >> $O/ro_tests.log; tail -6 $O/ro_tests.log
timeout 600 python bench.py --dense-only --steps 5 --warmup 2 > $O/ro_dense4.json 2> $O/ro_dense4.err; cut -c1-420 $O/ro_dense4.json; tail -3 $O/ro_dense4.err
LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 6000 > $O/ro_cfg4_path2.json 2> $O/ro_cfg4_path2.err; cat $O/ro_cfg4_path2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ro_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > $O/ro_ncu.log 2>&1
LGMI_TILE_PATH=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/ro_launches_cfg4.csv python tools/time_cfg4.py 6000 > $O/ro_ncu4.log 2>&1
echo done
