set -x
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r11_multi_tests.log 2>&1; tail -5 gpurun_out/r11_multi_tests.log
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-hard > gpurun_out/r11_bench2.json 2> gpurun_out/r11_bench2.err; tail -3 gpurun_out/r11_bench2.err; cut -c1-1500 gpurun_out/r11_bench2.json
