echo "== nproc"; nproc; echo "== lscpu"; lscpu | grep -i -E "model name|socket|numa|^cpu\(s\)|thread"; echo "== mem"; free -g | head -3
echo "== topo"; nvidia-smi topo -m 2>&1 | head -30
echo "== numa_node of gpus"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/class 2>/dev/null)" = "0x030200" ]; then echo "$d $(cat $d/numa_node) $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
echo "== nodes"; ls /sys/devices/system/node/ 2>/dev/null; cat /sys/devices/system/node/node*/cpulist 2>/dev/null
echo "== affinity"; python -c "import os; print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:64])"
echo "== cgroup cpu"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null
echo "== ulimit -l"; ulimit -l
