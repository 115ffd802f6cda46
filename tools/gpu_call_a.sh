lscpu | grep -i -E "socket|numa|model name|^CPU\(s\)" ; nvidia-smi topo -m 2>&1 | head -20; cat /sys/bus/pci/devices/*/numa_node 2>/dev/null | sort | uniq -c | head; nproc
python - <<'PY'
import torch
p = torch.cuda.get_device_properties(0)
print(p.pci_bus_id, getattr(p, "pci_domain_id", None), getattr(p, "pci_device_id", None))
import sys; sys.path.insert(0, '.')
import bench
print(bench.bind_to_gpu_numa_node(0))
PY
timeout 1200 python -m pytest tests/test_gpu_select_variants.py tests/test_gpu_nullable.py tests/test_host_gpu.py -m gpu -q > gpurun_out/pytest_a.log 2>&1; tail -8 gpurun_out/pytest_a.log
