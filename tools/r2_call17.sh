set -x
N=$1
if [ "$N" = "1" ]; then
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r17_bench1.json 2> gpurun_out/r17_bench1.err; tail -2 gpurun_out/r17_bench1.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_synthetic.py tests/test_gpu_runner.py -x -q > gpurun_out/r17_tests.log 2>&1; tail -3 gpurun_out/r17_tests.log
else
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 10 --warmup 3 --no-hard > gpurun_out/r17_bench${N}.json 2> gpurun_out/r17_bench${N}.err; tail -3 gpurun_out/r17_bench${N}.err | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r17_multi_tests.log 2>&1; tail -3 gpurun_out/r17_multi_tests.log
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r17_bench*.json')):
    for ln in open(f):
        if ln.startswith('{'):
            d=json.loads(ln); print(f, 'value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G frac %.3f'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, d['roofline']['frac']), d['check']['state_checksum'], d['check']['stage_histogram'], [round(x,3) for x in d['step_ms']])
PY
