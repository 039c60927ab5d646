#!/bin/bash
# Round-2 ncu captures (1 GPU).  Numbers printed under ncu are not bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
# (1) launch list of the bench command (the driver's flags, shortened): kernel shares of the step
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu > gpurun_out/r2_launches_bench.log 2>&1
# (2) full captures at C3: the force kernel, the gather, the scatter, the scan
timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_kernel_staged -s 3 -c 1 -f -o gpurun_out/r2_force_c3 python tools/quick_time.py C3 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"gather_f32|scatter_perm|scan_onepass" -s 9 -c 3 -f -o gpurun_out/r2_sort_c3 python tools/quick_time.py C3 > /dev/null 2>&1
# (3) the HBM-bound regime (C3-lo): every kernel of one step
timeout 600 ncu --set full --clock-control none -k regex:"force_kernel|gather_f32|scatter_perm|scan_onepass" -s 12 -c 4 -f -o gpurun_out/r2_step_c3lo python tools/quick_time.py C3lo > /dev/null 2>&1
ls -la gpurun_out/r2_*.ncu-rep gpurun_out/r2_launches.csv
