#!/bin/bash
# GPU round AI of r2: streamed end-to-end step at 2, 3 and 4 steps in flight (and 4 / 6 groups)
O=gpurun_out
mkdir -p $O
for d in 2 3 4; do
  timeout 300 python bench.py --no-strong --no-cfg5 --no-dense --no-cpu-baseline --no-cfg1 --e2e-depth $d --steps 20 > $O/rai_d$d.json 2> $O/rai_d$d.err
done
timeout 300 python bench.py --no-strong --no-cfg5 --no-dense --no-cpu-baseline --no-cfg1 --e2e-depth 3 --e2e-chunks 6 --steps 20 > $O/rai_d3c6.json 2> $O/rai_d3c6.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/rai_d*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); e=d['e2e']
        print(f, round(e['ms_per_step'],3), round(e['one_step_at_a_time']['ms_per_step'],3), round(e['link_probe']['alone_ms'],3))
    except Exception as ex: print(f, 'ERR', ex)
PY
echo done
