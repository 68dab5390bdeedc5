timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r3m_pytest.log 2>&1; tail -3 gpurun_out/r3m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-config-legs --kernels-json gpurun_out/r3m_kernels.json > gpurun_out/r3m_bench.json 2> gpurun_out/r3m_bench.err; tail -c 300 gpurun_out/r3m_bench.err
ncu --set full --clock-control none --import-source on -k regex:'bwd_kernel' -c 2 -o gpurun_out/prof_r3l_bwd -f python tools/resident_probe.py --batches 32 > gpurun_out/r3l_ncu_bwd.log 2>&1
