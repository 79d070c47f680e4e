#!/bin/bash
# r01: validate the rectangular tile kernel: full GPU suite, short-axis probes, ncu captures of its three regimes
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_rect.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_rect.log
python scripts/probe_smalldim.py > gpurun_out/smalldim_rect_final.txt 2>&1
python scripts/probe_cliffs.py 2>&1 | grep -E "17|, 5\)|\(5,|, 3\)|\(3,|outer|fill" > gpurun_out/cliffs_rect.txt
for c in shorty_f64 shorty_f32 shortx_f64; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:ew_tile_rect -s 2 -c 1 -f -o gpurun_out/r01_rect_$c python scripts/run_rect_shapes.py $c > gpurun_out/ncu_rect_$c.log 2>&1
done
python scripts/ncu_summary.py r01_ncu_rect_summary.json > gpurun_out/ncu_rect_summary.txt 2>&1
for c in shorty_f64 shorty_f32; do ncu -i gpurun_out/r01_rect_$c.ncu-rep --page source --csv > gpurun_out/rect_source_$c.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep; rm -f gpurun_out/*.ncu-rep
cat gpurun_out/smalldim_rect_final.txt gpurun_out/cliffs_rect.txt gpurun_out/ncu_rect_summary.txt
