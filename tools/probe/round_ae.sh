#!/bin/bash
# GPU round AE of r2: coverage-based skip in k_count_fast -- whole suite, cfg2 + cfg5 timing
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rae_tests.log 2>&1
echo "tests rc=$?" >> $O/rae_tests.log; tail -3 $O/rae_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dense --no-strong --no-cfg1 > $O/rae_bench.json 2> $O/rae_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/rae_bench.json'))
print('cfg2 ms %.4f k_pairs_fast %.4f frac %.4f e2e %.3f het %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['het_only']['ms_per_step']))
print('cfg5', [(p['cov'],p['min_common'],round(p['ms_per_step'],3)) for p in d['cfg5']['points']])
PY
echo done
