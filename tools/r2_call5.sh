set -x
python tools/probe.py --n 8400000 --iso soft --reps 2 --bank 0 > gpurun_out/r5_w32u3_nobank.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 > gpurun_out/r5_w32u3.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --tpg 2 > gpurun_out/r5_w32u2.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --warps 24 > gpurun_out/r5_w24u4.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --warps 16 > gpurun_out/r5_w16u6.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 1050624 > gpurun_out/r5_w32u3_b8.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 1050624 --warps 24 > gpurun_out/r5_w24u4_b8.log 2>&1
python tools/probe.py --n 8400000 --iso hard --reps 2 > gpurun_out/r5_hard.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r5_1m.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r5_launches.csv python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r5.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r5_launches_hard.csv python tools/probe.py --n 8400000 --iso hard --reps 1 > gpurun_out/ncu_r5h.log 2>&1
grep -E "layout|t= ?[1-9] |episode|checksum|Error|error" gpurun_out/r5_*.log | cut -c1-120
