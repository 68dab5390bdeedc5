timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r3u_pytest.log 2>&1; tail -3 gpurun_out/r3u_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/r3u_smoke.log 2>&1; tail -2 gpurun_out/r3u_smoke.log
timeout 900 python bench.py > gpurun_out/r3u_bench.json 2> gpurun_out/r3u_bench.err; tail -c 200 gpurun_out/r3u_bench.err; head -c 300 gpurun_out/r3u_bench.json; echo
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r3u_bench_ref.json 2> gpurun_out/r3u_bench_ref.err; head -c 400 gpurun_out/r3u_bench_ref.json
