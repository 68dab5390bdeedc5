timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "many_snapshots" > gpurun_out/r3t_pytest.log 2>&1; tail -3 gpurun_out/r3t_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-config-legs --kernels-json gpurun_out/r3t_kernels.json > gpurun_out/r3t_bench.json 2> gpurun_out/r3t_bench.err; tail -c 300 gpurun_out/r3t_bench.err
