set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for iso in soft hard; do
python tools/probe.py --n 8400000 --iso $iso --reps 2 > gpurun_out/r3_stream_$iso.log 2>&1
python tools/probe.py --n 8400000 --iso $iso --reps 2 --layout bcsr > gpurun_out/r3_hashed_$iso.log 2>&1
done
python tools/probe.py --n 8400000 --iso soft --reps 2 --warps 24 > gpurun_out/r3_stream_w24_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --warps 16 > gpurun_out/r3_stream_w16_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --tile 1024 > gpurun_out/r3_stream_t1024_soft.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r3_1m_stream_soft.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 --layout bcsr > gpurun_out/r3_1m_bcsr_soft.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'seirm|gather' -c 48 --csv --log-file gpurun_out/r3_launches_stream_soft.csv python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r3.log 2>&1
grep -E "layout|t= ?[0-9]+ |episode|checksum|Error|error" gpurun_out/r3_*.log | cut -c1-120
