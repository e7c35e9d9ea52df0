#!/bin/bash
# GPU round AP of r2 (last seconds of the budget): lg.stream_steps in the parity test and in the bench's e2e leg
O=gpurun_out
mkdir -p $O
timeout 40 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_pipelines" > $O/rap_tests.log 2>&1
echo "tests rc=$?" >> $O/rap_tests.log; tail -2 $O/rap_tests.log
timeout 60 python bench.py --no-strong --no-cfg5 --no-dense --no-cpu-baseline --no-cfg1 > $O/rap_bench.json 2> $O/rap_bench.err
echo "bench rc=$?"; tail -c 300 $O/rap_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/rap_bench.json').read().strip().splitlines()[-1]); e=d['e2e']
print(round(e['ms_per_step'],3), round(e['one_step_at_a_time']['ms_per_step'],3), e['steps_in_flight'])
PY
