set -x
N=$1
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r15_bench${N}.json 2> gpurun_out/r15_bench${N}.err; tail -3 gpurun_out/r15_bench${N}.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r15_bench*.json')):
    for ln in open(f):
        if ln.startswith('{'):
            d=json.loads(ln); print(f, 'value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G hard %.1f G'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9, (d['hard'] or {'value':0})['value']/1e9), d['check']['state_checksum'], d['check']['stage_histogram'], [round(x,3) for x in d['step_ms']])
PY
