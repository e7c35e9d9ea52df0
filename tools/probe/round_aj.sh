#!/bin/bash
# GPU round AJ of r2: lgmi_pipeline_collect (downloads of step k + 1 queued before step k is waited for)
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or pipelined or compact" > $O/raj_tests.log 2>&1
echo "tests rc=$?" >> $O/raj_tests.log; tail -3 $O/raj_tests.log
for v in "3 1" "4 1" "3 0"; do set -- $v
  timeout 300 python bench.py --no-strong --no-cfg5 --no-dense --no-cpu-baseline --no-cfg1 --e2e-depth $1 --e2e-collect $2 --steps 20 > $O/raj_d$1c$2.json 2> $O/raj_d$1c$2.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/raj_d*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
        print(f, round(e['ms_per_step'],3), round(e['one_step_at_a_time']['ms_per_step'],3), round(e['link_probe']['alone_ms'],3))
    except Exception as ex: print(f, 'ERR', ex)
PY
echo done
