#!/bin/bash
# GPU round W of r2 (final evidence, 1 GPU): launch lists of the three workloads, full ncu captures of the kernels that matter
O=gpurun_out
mkdir -p $O
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-dense --no-strong --no-cfg5 --no-cfg1"
CMD3="python bench.py --dense-only --steps 2 --warmup 1"
CMD4="python tools/time_cfg4.py 6000"
$CMD2 > $O/rw_plain2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/rw_launches.csv $CMD2 > $O/rw_ncu_l2.log 2>&1
$CMD3 > $O/rw_plain3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rw_launches_cfg3.csv $CMD3 > $O/rw_ncu_l3.log 2>&1
$CMD4 > $O/rw_plain4.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rw_launches_cfg4.csv $CMD4 > $O/rw_ncu_l4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_pairs_fast -s 5 -c 1 -o $O/rw_prof_k_pairs_fast $CMD2 > $O/rw_ncu_a.log 2>&1; tail -1 $O/rw_ncu_a.log
for k in k_gram_i8 k_other_fix k_dense_prep; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/rw_prof_$k $CMD3 > $O/rw_ncu_$k.log 2>&1; tail -1 $O/rw_ncu_$k.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_generic -s 2 -c 1 -o $O/rw_prof_k_pairs_generic2 $CMD3 > $O/rw_ncu_g2.log 2>&1; tail -1 $O/rw_ncu_g2.log
for k in k_tile_gram_ws k_tile_finish; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o $O/rw_prof_$k $CMD4 > $O/rw_ncu_$k.log 2>&1; tail -1 $O/rw_ncu_$k.log
done
echo done
