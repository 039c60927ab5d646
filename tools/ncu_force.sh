set -x
for k in 8 1; do
PLIFE_BINS=$k timeout 600 ncu --set full --clock-control none --import-source on -k regex:force_kernel_staged -s 3 -c 1 -f -o gpurun_out/force_r2a_k$k python tools/quick_time.py C3 > gpurun_out/ncu_r2a_k$k.log 2>&1
done
PLIFE_BINS=8 timeout 600 ncu --set full --clock-control none -k regex:gather_f32 -s 3 -c 1 -f -o gpurun_out/gather_r2a_k8 python tools/quick_time.py C3 > gpurun_out/ncu_r2a_g.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
