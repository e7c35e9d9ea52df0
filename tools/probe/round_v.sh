#!/bin/bash
# GPU round V of r2: tapered groups in the pipelined step
O=gpurun_out
mkdir -p $O
timeout 600 python tools/e2e_variants.py 4 5 6 8 > $O/rv_e2e_taper.txt 2>&1; grep tight $O/rv_e2e_taper.txt
LGMI_PIPE_TAPER=0 timeout 600 python tools/e2e_variants.py 4 5 6 8 > $O/rv_e2e_flat.txt 2>&1; grep tight $O/rv_e2e_flat.txt
LGMI_PIPE_DEBUG=1 timeout 300 python tools/e2e_variants.py 5 2>&1 | grep -v "^chunks" | tail -2 | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelin or compact" 2>&1 | tail -2
