set -x
python tools/probe.py --n 8400000 --iso soft --reps 2 > gpurun_out/r14_soft_u2.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --tpg 3 > gpurun_out/r14_soft_u3.log 2>&1
python tools/probe.py --n 8400000 --iso hard --reps 2 > gpurun_out/r14_hard_u2.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r14_launches_hard.csv python tools/probe.py --n 8400000 --iso hard --reps 1 > gpurun_out/ncu_r14b.log 2>&1
grep -E "t= ?[0-9]+ |episode|checksum|Error|error" gpurun_out/r14_*.log | cut -c1-120
