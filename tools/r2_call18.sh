set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r18_tests.log 2>&1; tail -8 gpurun_out/r18_tests.log | cut -c1-300
