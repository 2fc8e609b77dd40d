set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
python tools/probe.py --n 8400000 --iso soft --reps 2 > gpurun_out/r10_soft.log 2>&1
python tools/probe.py --n 8400000 --iso hard --reps 2 > gpurun_out/r10_hard.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 --grad 0 > gpurun_out/r10_1m_fwd.log 2>&1
ncu --metrics $M --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r10_launches_soft.csv python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r10a.log 2>&1
ncu --metrics $M --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r10_launches_hard.csv python tools/probe.py --n 8400000 --iso hard --reps 1 > gpurun_out/ncu_r10b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_gather' -s 3 -c 1 -o gpurun_out/r10_gather_full python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r10c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_update' -s 3 -c 1 -o gpurun_out/r10_update_full python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r10d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_gather' -s 10 -c 1 -o gpurun_out/r10_gather_hard_full python tools/probe.py --n 8400000 --iso hard --reps 1 > gpurun_out/ncu_r10e.log 2>&1
grep -E "layout|t= ?[0-9]+ |episode|checksum|Error|error" gpurun_out/r10_*.log | cut -c1-120
