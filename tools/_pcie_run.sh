#!/bin/bash
# box topology + the PCIe / host ceiling at 1..8 GPUs (tools/pcie_bw.cu)
mkdir -p gpurun_out
{
  echo "== nvidia-smi topo -m"; nvidia-smi topo -m
  echo "== lscpu"; lscpu | egrep 'Model name|Socket|NUMA|^CPU\(s\)|Thread|Core|Hypervisor|Virtualization'
  echo "== numa nodes"; ls /sys/devices/system/node/ | tr '\n' ' '; echo
  for n in /sys/devices/system/node/node*; do echo "$n: cpus $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== affinity of this shell: $(taskset -p $$)"
  echo "== pcie link"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv
} > gpurun_out/r02_box_topology.txt 2>&1
./tools/pcie_bw --gpus 1,2,4,8 --quick --seconds 0.5 > gpurun_out/r02_pcie_bw_scaling.jsonl 2> gpurun_out/r02_pcie_bw.err
./tools/pcie_bw --gpus 8 --seconds 0.5 > gpurun_out/r02_pcie_bw_variants_n8.jsonl 2>> gpurun_out/r02_pcie_bw.err
./tools/pcie_bw --gpus 1 --seconds 0.5 > gpurun_out/r02_pcie_bw_variants_n1.jsonl 2>> gpurun_out/r02_pcie_bw.err
tail -n 40 gpurun_out/r02_pcie_bw_scaling.jsonl gpurun_out/r02_pcie_bw_variants_n8.jsonl
