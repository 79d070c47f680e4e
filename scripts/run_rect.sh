#!/bin/bash
# r01 closing check (66 GPU-seconds left): elementwise / copy parity files + the elementwise probe on the final library
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests/test_gpu_elementwise.py tests/test_gpu_copy.py tests/test_gpu_core_func_ops.py -m gpu -x -q > gpurun_out/pytest_gpu_ew_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_ew_final.log
python scripts/probe_ew.py > gpurun_out/probe_ew.txt 2>&1; cat gpurun_out/probe_ew.txt
