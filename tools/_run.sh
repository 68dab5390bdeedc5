timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q > gpurun_out/r3r_pytest.log 2>&1; tail -3 gpurun_out/r3r_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --kernels-json gpurun_out/r3r_kernels.json > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; tail -c 300 gpurun_out/r3r_bench.err
