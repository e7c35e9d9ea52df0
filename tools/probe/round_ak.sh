#!/bin/bash
# GPU round AK of r2: host time inside begin / collect / finish of the streamed step, by group count
O=gpurun_out
mkdir -p $O
timeout 300 python tools/stream_host_times.py 30 > $O/rak_host.log 2>&1; cat $O/rak_host.log
LGMI_PIPE_DEBUG=1 timeout 300 python tools/stream_host_times.py 6 2>&1 | grep -A0 "lgmi pipeline" | tail -12 > $O/rak_dbg.log; tail -6 $O/rak_dbg.log
echo done
