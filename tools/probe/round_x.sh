#!/bin/bash
# GPU round X of r2: lists of k_pairs_fast built from class masks (no atomics in the counts loop) -- parity, cfg2 timing
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rx_tests.log 2>&1
echo "tests rc=$?" >> $O/rx_tests.log; tail -4 $O/rx_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dense --no-strong --no-cfg5 --no-cfg1 > $O/rx_bench_cfg2.json 2> $O/rx_bench_cfg2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/rx_bench_cfg2.json'))
print('cfg2 ms %.4f k_pairs_fast %.4f frac %.4f e2e %.3f het %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['het_only']['ms_per_step']))
PY
LGMI_TILE_PATH=2 timeout 600 python tools/time_cfg4.py 6000 > $O/rx_cfg4.json 2>/dev/null; cat $O/rx_cfg4.json
echo done
