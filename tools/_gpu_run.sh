lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"
nvidia-smi topo -m 2>&1 | head -20
nvidia-smi --query-gpu=index,pci.bus_id --format=csv
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302\|^0x0300" $d/class 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/local_cpulist)"; fi; done | head -12
python - <<'PY'
import os
print("affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:4], "...")
PY
cat /proc/meminfo | head -3
