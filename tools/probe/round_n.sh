#!/bin/bash
# GPU round N of r2: full ncu captures of k_site_mean_dense and k_other_fix as they are now, and of cfg4's finish / ordering kernels
O=gpurun_out
mkdir -p $O
CMD="python bench.py --dense-only --steps 2 --warmup 1"
$CMD > $O/rn_plain.log 2>&1 || exit 1
for k in k_other_fix k_site_mean_dense; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/rn_prof_$k $CMD > $O/rn_ncu_$k.log 2>&1
  tail -1 $O/rn_ncu_$k.log
done
export LGMI_TILE_PATH=2
CMD4="python tools/time_cfg4.py 6000"
$CMD4 > $O/rn_plain4.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pairs_generic -s 2 -c 1 -o $O/rn_prof_generic1 $CMD4 > $O/rn_ncu_g1.log 2>&1; tail -1 $O/rn_ncu_g1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_site_mean_dense -s 1 -c 1 -o $O/rn_prof_mean4 $CMD4 > $O/rn_ncu_m4.log 2>&1; tail -1 $O/rn_ncu_m4.log
echo done
