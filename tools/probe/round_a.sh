#!/bin/bash
# GPU round A of r2: topology, parity of everything new, cfg4 timing on both tile paths, one full bench line
mkdir -p gpurun_out
bash tools/probe/topo.sh
timeout 1500 python -m pytest tests -x -q -m gpu -k "mid_units or heavy_tail or many_sites or deep_unit_large or ragged or pipelined_step or compact_rows or device_pool or default_device or site_splice or ecdf or mip or kat or golden_units_batched or all_paths" > gpurun_out/ra_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ra_tests.log
tail -15 gpurun_out/ra_tests.log
LGMI_TILE_PATH=1 timeout 600 python tools/time_cfg4.py 6000 > gpurun_out/ra_cfg4_gram.json 2> gpurun_out/ra_cfg4_gram.err
LGMI_TILE_PATH=0 timeout 600 python tools/time_cfg4.py 6000 > gpurun_out/ra_cfg4_popc.json 2> gpurun_out/ra_cfg4_popc.err
cat gpurun_out/ra_cfg4_gram.json gpurun_out/ra_cfg4_popc.json
timeout 900 python -m pytest tests/test_cli_cfg1.py -x -q -m gpu > gpurun_out/ra_cli.log 2>&1
echo "cli rc=$?" >> gpurun_out/ra_cli.log
tail -5 gpurun_out/ra_cli.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/ra_bench.json 2> gpurun_out/ra_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/ra_bench.err
LGMI_TILE_PATH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ra_launches_cfg4.csv python tools/time_cfg4.py 6000 > gpurun_out/ra_ncu.log 2>&1
echo done
