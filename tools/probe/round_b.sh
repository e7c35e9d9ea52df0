#!/bin/bash
# GPU round B of r2: parity of everything new, cfg4 timing, full ncu of the two new kernels, e2e forms, one bench line
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -x -q -m gpu -k "mid_units or heavy_tail or many_sites or deep_unit_large or ragged or pipelined_step or compact_rows or device_pool or default_device or site_splice or ecdf or mip or kat or golden_units_batched or all_paths or cfg4_full" > $O/rb_tests.log 2>&1
echo "tests rc=$?" >> $O/rb_tests.log
tail -15 $O/rb_tests.log
LGMI_TILE_PATH=1 timeout 600 python tools/time_cfg4.py 6000 > $O/rb_cfg4_gram.json 2> $O/rb_cfg4_gram.err
LGMI_TILE_PATH=0 timeout 600 python tools/time_cfg4.py 6000 > $O/rb_cfg4_popc.json 2> $O/rb_cfg4_popc.err
cat $O/rb_cfg4_gram.json $O/rb_cfg4_popc.json
timeout 900 python -m pytest tests/test_cli_cfg1.py tests/test_batched_region.py -x -q -m gpu > $O/rb_cli.log 2>&1
echo "cli rc=$?" >> $O/rb_cli.log
tail -5 $O/rb_cli.log
timeout 600 python tools/e2e_variants.py 4 8 12 > $O/rb_e2e.txt 2>&1; cat $O/rb_e2e.txt
timeout 1500 python bench.py --steps 5 --warmup 3 > $O/rb_bench.json 2> $O/rb_bench.err
echo "bench rc=$?"; tail -c 1200 $O/rb_bench.err; cut -c1-400 $O/rb_bench.json
LGMI_TILE_PATH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rb_launches_cfg4.csv python tools/time_cfg4.py 6000 > $O/rb_ncu.log 2>&1
LGMI_TILE_PATH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_gram -s 2 -c 1 -o $O/rb_prof_tile_gram python tools/time_cfg4.py 6000 > $O/rb_ncu2.log 2>&1
LGMI_TILE_PATH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_finish -s 2 -c 1 -o $O/rb_prof_tile_finish python tools/time_cfg4.py 6000 > $O/rb_ncu3.log 2>&1
tail -2 $O/rb_ncu2.log $O/rb_ncu3.log
echo done
