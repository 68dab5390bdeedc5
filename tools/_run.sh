python -m pytest tests/test_gpu_resident2.py -x -q > gpurun_out/r2f_pytest.log 2>&1; tail -15 gpurun_out/r2f_pytest.log
python tools/resident_probe.py --batches 8,32 --clusters 0 --phases --out gpurun_out/r2f_probe.json > gpurun_out/r2f_probe.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2f_probe.json')):
    print(r['B'], 'fwd_train', r['fwd_train_us'], 'infer', r['fwd_infer_us'], 'bwd', r['bwd_us'], 'step', r['step_us'])
    print(' fwd  ', r['phases']['fwd'])
    print(' bwd  ', r['phases']['bwd'])
PY
tail -3 gpurun_out/r2f_probe.log | cut -c1-300
