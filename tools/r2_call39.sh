set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r39_tests.log 2>&1; tail -3 gpurun_out/r39_tests.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r39_smoke.log 2>&1; tail -2 gpurun_out/r39_smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r39_bench1.json 2> gpurun_out/r39_bench1.err; tail -2 gpurun_out/r39_bench1.err | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/r39_bench1.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value %.1f G  ms %.3f  e2e %.1f G (%.3f ms) fwd %.1f G frac %.3f worst %.3f launches %d'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['e2e']['ms_per_step'], d['fwd_only']['value']/1e9, d['roofline']['frac'], d['roofline']['worst_active_frac'], d['gpu_launches']), d['check']['state_checksum'], 'hard %.1f G frac %.3f'%(d['hard']['value']/1e9, d['hard']['roofline']['frac']), d.get('cpu_baseline'))
PY
