set -x
for iso in soft hard; do
python tools/probe.py --n 8400000 --iso $iso --reps 2 > gpurun_out/r4_stream_$iso.log 2>&1
done
python tools/probe.py --n 8400000 --iso soft --reps 2 --warps 24 > gpurun_out/r4_stream_w24_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --block 1050624 > gpurun_out/r4_stream_b8_soft.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r4_1m_stream_soft.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum --clock-control none -k regex:'seirm|gather' -c 48 --csv --log-file gpurun_out/r4_launches_stream_soft.csv python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r4.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'stream_gather' -s 3 -c 1 -o gpurun_out/r4_gather_full python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r4b.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'seirm_kernel' -s 4 -c 1 -o gpurun_out/r4_tupdate_full python tools/probe.py --n 8400000 --iso soft --reps 1 > gpurun_out/ncu_r4c.log 2>&1
grep -E "layout|t= ?[0-9]+ |episode|checksum|Error|error" gpurun_out/r4_*.log | cut -c1-120
