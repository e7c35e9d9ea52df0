#!/bin/bash
# GPU round J of r2: full ncu captures of the deep-unit kernels around the GEMM
O=gpurun_out
mkdir -p $O
CMD="python bench.py --dense-only --steps 2 --warmup 1"
$CMD > $O/rj_plain.log 2>&1 || exit 1
for k in k_other_fix k_dense_prep k_pairs_generic k_site_mean_dense; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $O/rj_prof_$k $CMD > $O/rj_ncu_$k.log 2>&1
  tail -1 $O/rj_ncu_$k.log
done
echo done
