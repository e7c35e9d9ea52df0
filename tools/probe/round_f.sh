#!/bin/bash
# GPU round F of r2: end-to-end step, number of groups x forms; the link alone at this step's byte counts
O=gpurun_out
mkdir -p $O
timeout 600 python tools/e2e_variants.py 4 6 8 12 16 > $O/rf_e2e.txt 2>&1; cat $O/rf_e2e.txt
LGMI_PIPE_DEBUG=1 timeout 300 python tools/e2e_variants.py 8 2>&1 | grep -v "^chunks" | tail -4 > $O/rf_e2e_debug.txt; cat $O/rf_e2e_debug.txt
timeout 300 python tools/pcie_probe.py > $O/rf_pcie.txt 2>&1; cat $O/rf_pcie.txt
