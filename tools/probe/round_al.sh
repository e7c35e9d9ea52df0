#!/bin/bash
# GPU round AL of r2: streamed step by groups per step x steps in flight
O=gpurun_out
mkdir -p $O
timeout 600 python tools/stream_host_times.py 40 1x3 1x4 1x5 2x3 2x4 2x5 4x4 > $O/ral_host.log 2>&1; cat $O/ral_host.log
echo done
