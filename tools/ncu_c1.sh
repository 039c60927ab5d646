timeout 300 ncu --set full --clock-control none --import-source on -k regex:force_kernel_staged -s 30 -c 1 -f -o gpurun_out/force_r2_c1 python tools/quick_time.py C1 > gpurun_out/ncu_c1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 24 --csv --log-file gpurun_out/launches_c1.csv python tools/quick_time.py C1 > /dev/null 2>&1
tail -30 gpurun_out/launches_c1.csv | cut -d, -f5,12- 
