set -x

timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r9_bench2.json 2> gpurun_out/r9_bench2.err; tail -5 gpurun_out/r9_bench2.err; cat gpurun_out/r9_bench2.json

