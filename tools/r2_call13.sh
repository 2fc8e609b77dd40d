set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r13_tests.log 2>&1; tail -5 gpurun_out/r13_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r13_smoke.log 2>&1; tail -2 gpurun_out/r13_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err; tail -2 gpurun_out/r13_bench.err | cut -c1-300; cut -c1-1200 gpurun_out/r13_bench.json
