#!/bin/bash
# GPU round AH of r2: lgmi_pipeline_begin* / lgmi_pipeline_finish (two steps in flight) -- parity, then the bench line
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipeline or pipelined or compact" > $O/rah_tests.log 2>&1
echo "tests rc=$?" >> $O/rah_tests.log; tail -3 $O/rah_tests.log
timeout 600 python bench.py --no-strong --no-cfg5 --no-dense --no-cpu-baseline --no-cfg1 > $O/rah_bench.json 2> $O/rah_bench.err
echo "bench rc=$?"; tail -c 600 $O/rah_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/rah_bench.json').read().strip().splitlines()[-1])
e=d['e2e']; print({k:e[k] for k in ('value','ms_per_step','one_step_at_a_time','serial_upload_run_download_ms','gpu_launches')}); print(e['link_probe'])
PY
echo done
