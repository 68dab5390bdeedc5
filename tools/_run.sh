python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -k "gat_conv or model_matches or large or tile or train_step" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-config-legs --kernels-json gpurun_out/r2l_kernels.json > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; tail -2 gpurun_out/r2l_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2l_kernels.json'))
for r in d['hbm']: print(r['kernel'], round(r['us'],1), r['bytes_per_node'], round(r['GBps']), round(r['GBps']/d['peak_gbs'],3))
PY
python bench.py --batch 1024 --steps 10 --warmup 3 --skip-cpu-baseline --skip-config-legs --skip-kernel-leg 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('B=1024', d['value'], d['ms_per_step'])"
