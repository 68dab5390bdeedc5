# scratch script for gpurun calls: `gpurun --timeout 900 -- 'bash tools/_run.sh'`
python -m pytest tests -m gpu -x -q
python bench.py --steps 20 --warmup 5
