#!/bin/bash
# topology probe of the GPU box: what bind_near_gpu can rely on
out=gpurun_out/topo.txt
{
echo "== nproc"; nproc; 
echo "== lscpu"; lscpu | head -40
echo "== nvidia-smi topo -m"; nvidia-smi topo -m
echo "== nvidia-smi -q pci"; nvidia-smi --query-gpu=index,pci.bus_id,pci.domain,pci.bus,pci.device --format=csv
echo "== sysfs numa_node"; for d in /sys/bus/pci/devices/*; do c=$(cat $d/class 2>/dev/null); if [[ "$c" == 0x0302* || "$c" == 0x0300* ]]; then echo "$d class=$c numa=$(cat $d/numa_node) local_cpulist=$(cat $d/local_cpulist)"; fi; done
echo "== /sys/devices/system/node"; ls /sys/devices/system/node/ ; for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist)"; done
echo "== affinity"; python - <<'PY'
import os
print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8], '...')
PY
echo "== numactl"; which numactl; numactl -H 2>&1 | head -20
echo "== meminfo"; head -5 /proc/meminfo
echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
echo "== set_mempolicy probe"; python - <<'PY'
import ctypes, os
libc = ctypes.CDLL(None, use_errno=True)
# get_mempolicy syscall 239 on x86_64
mode = ctypes.c_int(); mask = (ctypes.c_ulong*16)()
r = libc.syscall(239, ctypes.byref(mode), mask, 1024, 0, 0)
print("get_mempolicy rc", r, "errno", ctypes.get_errno(), "mode", mode.value, list(mask)[:2])
PY
} > $out 2>&1
echo done
