set -x
python tools/probe.py --n 8400000 --iso soft --reps 2 > gpurun_out/r6_b6.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 1050624 > gpurun_out/r6_b8.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 700416 > gpurun_out/r6_b12.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 525312 > gpurun_out/r6_b16.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 350208 > gpurun_out/r6_b24.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 700416 --warps 24 > gpurun_out/r6_b12_w24.log 2>&1
python tools/probe.py --n 8400000 --iso hard --reps 2 --block 700416 > gpurun_out/r6_b12_hard.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r6_1m.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 --block 250880 > gpurun_out/r6_1m_b4.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'seirm|stream' -c 48 --csv --log-file gpurun_out/r6_launches_b12.csv python tools/probe.py --n 8400000 --iso soft --reps 1 --block 700416 > gpurun_out/ncu_r6.log 2>&1
grep -E "layout|t= ?[1-9] |episode|checksum|Error|error" gpurun_out/r6_*.log | cut -c1-120
