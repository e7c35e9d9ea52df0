#!/bin/bash
# GPU round C of r2: the three mid-depth paths (parity at full cfg4 size, timing), finish-kernel occupancy variants
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mid_units or heavy_tail or many_sites or deep_unit_large or cfg4_full or pipelined_step or compact_rows" > $O/rc_tests.log 2>&1
echo "tests rc=$?" >> $O/rc_tests.log
tail -12 $O/rc_tests.log
for p in 2 1 0; do
  LGMI_TILE_PATH=$p timeout 600 python tools/time_cfg4.py 6000 > $O/rc_cfg4_path$p.json 2> $O/rc_cfg4_path$p.err; cat $O/rc_cfg4_path$p.json
done
LGMI_LIB=build/liblgmi_fin3.so LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 6000 > $O/rc_cfg4_path2_fin3.json 2>&1; cat $O/rc_cfg4_path2_fin3.json
for p in 2 1; do
  LGMI_TILE_PATH=$p timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rc_launches_path$p.csv python tools/time_cfg4.py 6000 > $O/rc_ncu_path$p.log 2>&1
done
LGMI_LIB=build/liblgmi_fin3.so LGMI_TILE_PATH=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rc_launches_path2_fin3.csv python tools/time_cfg4.py 6000 > $O/rc_ncu_fin3.log 2>&1
LGMI_TILE_PATH=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tile_gram_ws -s 2 -c 1 -o $O/rc_prof_tile_gram_ws python tools/time_cfg4.py 6000 > $O/rc_ncu2.log 2>&1
tail -2 $O/rc_ncu2.log
echo done
