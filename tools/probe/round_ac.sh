#!/bin/bash
# GPU round AC of r2: deep units -- lists / transposed planes first, then k_other_fix beside k_dense_x, k_gram_i8 alone
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_all_paths_agree.py -x -q -m gpu -k "dense or cfg3 or deep_unit or all_paths" > $O/rac_tests.log 2>&1
echo "tests rc=$?" >> $O/rac_tests.log; tail -3 $O/rac_tests.log
for i in 1 2; do timeout 300 python bench.py --dense-only --steps 5 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin)['dense']; print('cfg3 ms %.4f gram %.4f issued_frac %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['issued_frac']))"; done
LGMI_DENSE_PATH=9 timeout 300 python bench.py --dense-only --steps 5 --warmup 2 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin)['dense']; print('nine: cfg3 ms %.4f gram %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/rac_launches_cfg3.csv python bench.py --dense-only --steps 2 --warmup 1 > /dev/null 2>&1
echo done
