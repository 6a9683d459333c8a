#!/bin/bash
# Round-2 ncu captures (run on the GPU box through gpurun, ONE GPU): launch list of the default bench command,
# full captures of the dominant kernel and of the exact-stage kernels.  Outputs land in gpurun_out/.
set -x
cd "$(dirname "$0")/.."
FILT='--kernel-name-base demangled -k regex:msb::|cub::DeviceRadix|cub::DeviceScan'
ncu --metrics gpu__time_duration.sum --clock-control none $FILT -c 2200 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/r2_launches_bench.json 2> gpurun_out/r2_launches_bench.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:prefilter_tc_kernel -s 12 -c 1 \
    -o gpurun_out/r2_prefilter -f python bench.py --steps 1 --warmup 3 --no-cpu --scale 0.12 > /dev/null 2> gpurun_out/r2_prefilter.err
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:exact_records_kernel|exact_dirty_kernel|select_ranks_kernel' -s 6 -c 3 \
    -o gpurun_out/r2_exact -f python bench.py --steps 1 --warmup 3 --no-cpu --scale 0.12 > /dev/null 2> gpurun_out/r2_exact.err
ls -la gpurun_out/r2_*
