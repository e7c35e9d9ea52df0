#!/bin/bash
# GPU round AB of r2: "other" cells of k_pairs_fast scattered per (site, word) -- whole suite, cfg2 timing
O=gpurun_out
mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/rab_tests.log 2>&1
echo "tests rc=$?" >> $O/rab_tests.log; tail -3 $O/rab_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-dense --no-strong --no-cfg1 > $O/rab_bench.json 2> $O/rab_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/rab_bench.json'))
print('cfg2 ms %.4f k_pairs_fast %.4f frac %.4f e2e %.3f het %.3f' % (d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['het_only']['ms_per_step']))
print('cfg5', [(p['cov'],p['min_common'],round(p['ms_per_step'],3)) for p in d['cfg5']['points']][::3])
PY
echo done
