set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hard > gpurun_out/r40_bench_plain.json 2> gpurun_out/r40_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r40_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hard > gpurun_out/ncu_r40f.log 2>&1
ncu --metrics $M --clock-control none -k regex:'seirm|stream|init_encode' -c 48 --csv --log-file gpurun_out/r40_launches_soft.csv python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 1 --brief 1 > gpurun_out/ncu_r40a.log 2>&1
ls -la gpurun_out | grep r40
