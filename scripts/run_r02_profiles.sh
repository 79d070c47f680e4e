#!/bin/bash
# The GPU-box command behind profiles/r02_*: launch list of the bench step, one `ncu --set full` capture per round-2
# kernel (and of the dominant tile kernel), summarised on the box (the .ncu-rep files are 5-40 MB each and not kept).
#   gpurun --timeout 1500 -- 'bash scripts/run_r02_profiles.sh'
cd /root/repo; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-per-config > gpurun_out/r02_launches_bench.log 2>&1
cap() {  # cap <case> <kernel regex> [env]
  env $3 timeout 200 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/r02_$1 \
      python scripts/run_r02_shapes.py $1 > gpurun_out/ncu_r02_$1.log 2>&1
}
cap tile_cfg2 ew_tile_kernel RC_TILE_BULK=0
cap tma_cfg2 ew_tile_tma_kernel RC_TILE_BULK=1
cap narrow_u8 ew_tile_narrow_kernel X=1
cap narrow_i16 ew_tile_narrow_kernel X=1
cap outer_f64 ew_outer_kernel X=1
cap cast_u8_f32 ew_kernel X=1
python scripts/ncu_summary.py r02_ncu_full_summary.json > gpurun_out/r02_ncu_summary.txt 2>&1
rm -f gpurun_out/*.ncu-rep
tail -60 gpurun_out/r02_ncu_summary.txt
