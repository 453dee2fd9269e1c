#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
O=gpurun_out
{
nproc; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpu.max 2>/dev/null
ls /sys/devices/system/node/ | grep node
for n in /sys/devices/system/node/node*; do echo $n $(cat $n/cpulist) $(grep MemTotal $n/meminfo); done
nvidia-smi topo -m
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo $d $(cat $d/numa_node); fi; done
python - <<'P'
import os
print('affinity', sorted(os.sched_getaffinity(0)))
P
cat /proc/self/status | grep -i "Mems_allowed_list\|Cpus_allowed_list"
} > $O/r2c_topology.log 2>&1
cat $O/r2c_topology.log
