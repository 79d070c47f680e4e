#!/bin/bash
# r01 final validation of the rectangular tile kernel: full GPU suite, smoke, short-axis probes, ncu captures of its
# three regimes, and the headline bench (no regression check)
cd /root/repo; mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_rect.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_rect.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scripts/probe_smalldim.py > gpurun_out/smalldim_rect_final.txt 2>&1
python scripts/probe_cliffs.py > gpurun_out/probe_cliffs.txt 2>&1
for c in shorty_f64 shorty_f32 shortx_f64; do
  timeout 120 ncu --set full --clock-control none --import-source on -k regex:ew_tile_rect -s 2 -c 1 -f -o gpurun_out/r01_rect_$c python scripts/run_rect_shapes.py $c > gpurun_out/ncu_rect_$c.log 2>&1
done
python scripts/ncu_summary.py r01_ncu_rect_summary.json > gpurun_out/ncu_rect_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
python bench.py > gpurun_out/bench_n1_rect.json 2> gpurun_out/bench_n1_rect.err; echo "bench rc=$?"
cat gpurun_out/smalldim_rect_final.txt; grep -E "17|, 5\)|\(5,|, 3\)|\(3,|CLIFF" gpurun_out/probe_cliffs.txt; grep -E "kernel =|time_duration|dram_throughput|issue_active|registers_per" gpurun_out/ncu_rect_summary.txt; cat gpurun_out/bench_n1_rect.json
