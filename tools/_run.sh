python -m pytest tests/test_gpu_resident2.py -x -q 2>&1 | tail -2
python tools/resident_probe.py --batches 8,32 --clusters 0 --phases --out gpurun_out/r2j_probe.json > gpurun_out/r2j_probe.log 2>&1; python - <<PY
import json
for r in json.load(open('gpurun_out/r2j_probe.json')):
    print(r['B'], 'fwd_train', round(r['fwd_train_us'],1), 'infer', round(r['fwd_infer_us'],1), 'bwd', round(r['bwd_us'],1), 'step', round(r['step_us'],1))
    print(' fwd  ', r['phases']['fwd'])
    print(' bwd  ', r['phases']['bwd'])
PY
