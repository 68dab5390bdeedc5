timeout 300 python tools/resident_probe.py --batches 32 --phases --out gpurun_out/r3o_probe.json > gpurun_out/r3o_probe.log 2>&1
tail -c 1300 gpurun_out/r3o_probe.log | head -c 1000
