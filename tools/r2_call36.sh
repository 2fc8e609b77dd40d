set -x
timeout 700 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r36_multi_tests.log 2>&1; tail -3 gpurun_out/r36_multi_tests.log | cut -c1-300
for fam in rand comm; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 10 --warmup 3 --no-hard --family $fam > gpurun_out/r36_bench2_$fam.json 2> gpurun_out/r36_bench2_$fam.err; tail -2 gpurun_out/r36_bench2_$fam.err | cut -c1-200
done
python - <<'PY'
import json
for fam in ('rand','comm'):
  for ln in open('gpurun_out/r36_bench2_%s.json'%fam):
    if ln.startswith('{'):
        d=json.loads(ln); print(fam,'N=%d value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G frac %.3f launches %d'%(d['n_gpus'], d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, d['roofline']['frac'], d['gpu_launches']), d['check']['state_checksum'], d.get('halo'))
PY
