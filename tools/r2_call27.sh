set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r27_tests.log 2>&1; tail -4 gpurun_out/r27_tests.log | cut -c1-300
python tools/probe_host.py --agents 1050000 > gpurun_out/r27_host_1m.log 2>&1; head -4 gpurun_out/r27_host_1m.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-hard > gpurun_out/r27_bench1.json 2> gpurun_out/r27_bench1.err; tail -2 gpurun_out/r27_bench1.err | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/r27_bench1.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G frac %.3f launches %d'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, d['roofline']['frac'], d['gpu_launches']), d['check']['state_checksum'])
PY
