set -x
python tools/probe_host.py --agents 1050000 > gpurun_out/r20_host_1m.log 2>&1; head -3 gpurun_out/r20_host_1m.log
python tools/probe_host.py --agents 1050000 --no-grad > gpurun_out/r20_host_1m_fwd.log 2>&1; head -3 gpurun_out/r20_host_1m_fwd.log
python tools/probe_host.py --agents 8400000 > gpurun_out/r20_host_8m.log 2>&1; head -3 gpurun_out/r20_host_8m.log
