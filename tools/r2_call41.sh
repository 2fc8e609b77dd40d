python tools/probe_host.py --agents 1050000 > gpurun_out/r41_host_1m.log 2>&1; head -3 gpurun_out/r41_host_1m.log
