#!/bin/bash
# GPU-box command behind profiles/r02b_*: `ncu --set full` of the two tile kernels added late in round 2
# (ew_tile_wide_kernel, ew_tile_short_kernel) and of the square tile they replace, plus the launch list of the bench step.
#   gpurun --timeout 1500 -- 'bash scripts/run_r02b_profiles.sh'
cd /root/repo; mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r02b_launches_bench.log 2>&1
cap() {  # cap <case> <kernel regex> [env]
  env $3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/r02b_$1 \
      python scripts/run_r02_shapes.py $1 > gpurun_out/ncu_r02b_$1.log 2>&1
}
cap wide_cfg2 ew_tile_wide_kernel X=1
cap tile_cfg2 ew_tile_kernel RC_TILE_WIDE=0
cap short_f32_deint ew_tile_short_kernel X=1
cap short_f32_inter ew_tile_short_kernel X=1
cap short_u8_deint ew_tile_short_kernel X=1
cap short_u8_inter ew_tile_short_kernel X=1
python scripts/ncu_summary.py r02b_ncu_full_summary.json > gpurun_out/r02b_ncu_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
tail -80 gpurun_out/r02b_ncu_summary.txt
