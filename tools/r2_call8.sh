set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; tail -3 gpurun_out/r8_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r8_ref.json 2> gpurun_out/r8_ref.err; tail -5 gpurun_out/r8_ref.err
cat gpurun_out/r8_bench.json gpurun_out/r8_ref.json
