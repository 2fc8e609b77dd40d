set -x
python tools/probe.py --n 8400000 --iso soft --reps 2 > gpurun_out/r7_soft.log 2>&1
python tools/probe.py --n 8400000 --iso hard --reps 2 > gpurun_out/r7_hard.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r7_1m.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r7_launches.csv python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r7.log 2>&1
grep -E "layout|t= ?[0-9]+ |episode|checksum|Error|error" gpurun_out/r7_*.log | cut -c1-120
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r7_tests.log 2>&1; tail -15 gpurun_out/r7_tests.log
