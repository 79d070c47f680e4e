#!/bin/bash
# r01 closing run: full GPU suite + smoke on the final library; one ncu capture of the outer-sum launch (known limit)
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 2 -c 1 -f -o gpurun_out/r01_outer python scripts/run_outer.py > gpurun_out/ncu_outer.log 2>&1
python scripts/ncu_summary.py r01_ncu_outer_summary.json > gpurun_out/ncu_outer_summary.txt 2>&1
ncu -i gpurun_out/r01_outer.ncu-rep --page source --csv > gpurun_out/outer_source.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/ncu_outer_summary.txt
