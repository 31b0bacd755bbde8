#!/bin/bash
# Launch list of OUR kernels over a short bench run (B200_PROFILING.md: gpu__time_duration pass). Usage:
#   benchmarks/ncu_launches.sh gpurun_out/name.csv [extra bench.py flags]
out=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none \
  -k "regex:^(bbox|build_score|camera_prep|cell_|gemm_kernel|init_bbox|init_minmax|kept_|mask_offsets|refine_|row_normalize|scan_|seg_histogram|segmented_|unpack_|view_|visibility_|compact_|scatter_|pixel_|patch_|minmax_|zero_|vox_|project_visibility|sort_|pair_|morton_|vis_)" \
  --csv --log-file "$out" python bench.py --steps 2 --warmup 1 --no-e2e --no-extras --no-cpu-baseline --job-scenes 0 "$@" > /dev/null 2>&1
python benchmarks/launch_table.py "$out"
