#!/bin/bash
# Host / GPU topology of the box a gpurun call landed on (NUMA nodes, CPU affinity, PCIe / NVLink matrix):
# what the e2e staging path and the multi-GPU numbers have to be read against.  Writes gpurun_out/topology.txt.
mkdir -p gpurun_out
{
  echo "== nproc / affinity =="; nproc; grep -i cpus_allowed_list /proc/self/status
  echo "== lscpu =="; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"
  echo "== numa nodes =="; for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== cpuset =="; cat /sys/fs/cgroup/cpuset.cpus.effective /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null
  echo "== memory =="; free -g | head -2
  echo "== gpus =="; nvidia-smi --query-gpu=index,name,pci.bus_id,memory.total,clocks.max.sm,clocks.max.mem --format=csv
  for d in $(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader); do b=$(echo ${d#0000} | tr 'A-Z' 'a-z'); echo "$d numa_node=$(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)"; done
  echo "== topo =="; nvidia-smi topo -m
} > gpurun_out/topology.txt 2>&1
