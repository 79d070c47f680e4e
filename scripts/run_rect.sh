#!/bin/bash
# The GPU-box command behind the round-1 numbers of ew_tile_rect_kernel and the rank-2 scalar kernel
# (gpurun -- 'bash scripts/run_rect.sh'): full GPU suite, smoke, the short-axis / cliff / elementwise probes, one
# ncu --set full capture per regime of the rectangular tile kernel and of the outer-sum launch, summarised on the box.
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python scripts/probe_smalldim.py > gpurun_out/smalldim_rect_final.txt 2>&1
RC_TILE_RECT=0 python scripts/probe_smalldim.py > gpurun_out/smalldim_rect0.txt 2>&1
python scripts/probe_ew.py > gpurun_out/probe_ew.txt 2>&1
python scripts/probe_cliffs.py > gpurun_out/probe_cliffs.txt 2>&1
for c in shorty_f64 shorty_f32 shortx_f64; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:ew_tile_rect -s 2 -c 1 -f -o gpurun_out/r01_rect_$c python scripts/run_rect_shapes.py $c > gpurun_out/ncu_rect_$c.log 2>&1
done
python scripts/ncu_summary.py r01_ncu_rect_summary.json > gpurun_out/ncu_rect_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
timeout 100 ncu --set full --clock-control none --import-source on -k regex:ew_kernel -s 2 -c 1 -f -o gpurun_out/r01_outer python scripts/run_outer.py > gpurun_out/ncu_outer.log 2>&1
python scripts/ncu_summary.py r01_ncu_outer_summary.json > gpurun_out/ncu_outer_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
cat gpurun_out/smalldim_rect_final.txt gpurun_out/probe_ew.txt gpurun_out/ncu_rect_summary.txt
