timeout 800 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r38_multi_tests.log 2>&1; tail -5 gpurun_out/r38_multi_tests.log | cut -c1-300
