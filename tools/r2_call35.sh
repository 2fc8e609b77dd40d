set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r35_multi_tests.log 2>&1; tail -3 gpurun_out/r35_multi_tests.log | cut -c1-300
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-hard > gpurun_out/r35_bench2.json 2> gpurun_out/r35_bench2.err; tail -2 gpurun_out/r35_bench2.err | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/r35_bench2.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('N=%d value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G frac %.3f launches %d'%(d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, d['roofline']['frac'], d['gpu_launches']), d['check']['state_checksum'])
PY
