set -x
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 2 > gpurun_out/r29_soft.log 2>&1
python tools/probe.py --n 8400000 --iso hard --f32 1 --reps 2 > gpurun_out/r29_hard.log 2>&1
python tools/probe.py --n 1000000 --iso soft --f32 1 --reps 2 --grad 0 > gpurun_out/r29_1m_fwd.log 2>&1
ncu --metrics $M --clock-control none -k regex:'seirm|stream|init_encode' -c 48 --csv --log-file gpurun_out/r29_launches_soft.csv python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 1 > gpurun_out/ncu_r29a.log 2>&1
ncu --metrics $M --clock-control none -k regex:'seirm|stream|init_encode' -c 48 --csv --log-file gpurun_out/r29_launches_hard.csv python tools/probe.py --n 8400000 --iso hard --f32 1 --reps 1 > gpurun_out/ncu_r29b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_gather' -s 3 -c 1 -o gpurun_out/r29_gather_full python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 1 > gpurun_out/ncu_r29c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_update' -s 3 -c 1 -o gpurun_out/r29_update_full python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 1 > gpurun_out/ncu_r29d.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'seirm_episode_bwd' -c 1 -o gpurun_out/r29_bwd_full python tools/probe.py --n 8400000 --iso soft --f32 1 --reps 1 > gpurun_out/ncu_r29e.log 2>&1
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hard > gpurun_out/r29_bench_plain.json 2> gpurun_out/r29_bench_plain.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r29_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-hard > gpurun_out/ncu_r29f.log 2>&1
grep -E "hot steps|encode us|episode|bwd|checksum|Error|error" gpurun_out/r29_*.log | cut -c1-150
ls -la gpurun_out/ | grep r29
