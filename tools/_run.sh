timeout 900 python -m pytest tests/test_gpu_resident2.py -x -q > gpurun_out/r4b_pytest.log 2>&1; tail -3 gpurun_out/r4b_pytest.log
timeout 300 python tools/resident_probe.py --batches 32 --phases --out gpurun_out/r4b_probe.json > gpurun_out/r4b_probe.log 2>&1
tail -c 1300 gpurun_out/r4b_probe.log | head -c 700
