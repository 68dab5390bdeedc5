ncu --set full --clock-control none --import-source on -k regex:'^bwd_kernel' -s 2 -c 2 -o gpurun_out/prof_r2k_res2_bwd -f python tools/resident_probe.py --batches 32 --clusters 0 > gpurun_out/r2k_ncu_bwd.log 2>&1
ls -la gpurun_out/prof_r2k_res2_bwd.ncu-rep
