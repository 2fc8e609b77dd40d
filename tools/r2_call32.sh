set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r32_tests.log 2>&1; tail -4 gpurun_out/r32_tests.log | cut -c1-300
python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 2 --brief 1 > gpurun_out/r32_soft.log 2>&1
grep -E "hot steps|encode us|episode|bwd|checksum|Error|error" gpurun_out/r32_soft.log gpurun_out/r32_hard.log | cut -c1-200
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r32_bench1.json 2> gpurun_out/r32_bench1.err; tail -2 gpurun_out/r32_bench1.err | cut -c1-200
python - <<'PY'
import json
for ln in open('gpurun_out/r32_bench1.json'):
    if ln.startswith('{'):
        d=json.loads(ln); print('value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G frac %.3f worst %.3f launches %d'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, d['roofline']['frac'], d['roofline']['worst_active_frac'], d['gpu_launches']), d['check']['state_checksum'], 'hard %.1f G frac %.3f'%(d['hard']['value']/1e9, d['hard']['roofline']['frac']))
PY
