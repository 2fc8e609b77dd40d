set -x
N=$1
AGENTTORCH_B200_OVERLAP_EXCHANGE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-hard > gpurun_out/r16_bench${N}_overlap.json 2> gpurun_out/r16_bench${N}_overlap.err; tail -3 gpurun_out/r16_bench${N}_overlap.err | cut -c1-300
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r16_bench*.json')):
    for ln in open(f):
        if ln.startswith('{'):
            d=json.loads(ln); print(f, 'value %.1f G  ms %.3f  e2e %.1f G  fwd %.1f G'%(d['value']/1e9, d['ms_per_step'], d['e2e']['value']/1e9, d['fwd_only']['value']/1e9), d['check']['state_checksum'])
PY
