python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1; tail -4 gpurun_out/r2i_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -3 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
print(json.dumps(d['roofline'])[:1500])
for l in d['configs']: print(l['config'], l.get('value'), l.get('ms_per_step'), l.get('error'))
PY
