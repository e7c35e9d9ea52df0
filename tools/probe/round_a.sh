#!/bin/bash
# first GPU round of r2: k_tile_gram parity + cfg4 timing + topology
mkdir -p gpurun_out
bash tools/probe/topo.sh
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "mid_units or heavy_tail or many_sites or deep_unit_large or ragged" > gpurun_out/ra_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/ra_tests.log
tail -5 gpurun_out/ra_tests.log
LGMI_TILE_PATH=1 timeout 600 python tools/time_cfg4.py 6000 > gpurun_out/ra_cfg4_gram.json 2> gpurun_out/ra_cfg4_gram.err
LGMI_TILE_PATH=0 timeout 600 python tools/time_cfg4.py 6000 > gpurun_out/ra_cfg4_popc.json 2> gpurun_out/ra_cfg4_popc.err
cat gpurun_out/ra_cfg4_gram.json gpurun_out/ra_cfg4_popc.json
LGMI_TILE_PATH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/ra_launches_cfg4.csv python tools/time_cfg4.py 6000 > gpurun_out/ra_ncu.log 2>&1
echo done
