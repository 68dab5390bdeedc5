timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r4e_pytest.log 2>&1; tail -3 gpurun_out/r4e_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline > gpurun_out/r4e_bench.json 2> gpurun_out/r4e_bench.err; tail -c 200 gpurun_out/r4e_bench.err
for b in 64 128; do timeout 300 python bench.py --steps 20 --warmup 5 --batch $b --skip-cpu-baseline --skip-kernel-leg --skip-config-legs > gpurun_out/r4e_bench_b$b.json 2>> gpurun_out/r4e_bench.err; done
