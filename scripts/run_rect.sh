#!/bin/bash
# r01: compile-time rank-2 flat kernels: full GPU suite, smoke, elementwise probes
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python scripts/probe_ew.py > gpurun_out/probe_ew.txt 2>&1
python scripts/probe_cliffs.py > gpurun_out/probe_cliffs.txt 2>&1
cat gpurun_out/probe_ew.txt; grep -E "outer|fill|\(n,n\)|CLIFF|a\[|flip|, 3\)|, 5\)|, 17\)" gpurun_out/probe_cliffs.txt
