set -x
for iso in soft hard; do
python tools/probe.py --n 8400000 --iso $iso --reps 2 --layout packed > gpurun_out/r2_pk_default_$iso.log 2>&1
done
python tools/probe.py --n 8400000 --iso soft --reps 2 --layout packed --epc 768 --pw 10 > gpurun_out/r2_pk_e768_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --layout packed --epc 512 --pw 12 --block 933888 > gpurun_out/r2_pk_b9_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 2 --layout packed --epc 1024 --pw 8 > gpurun_out/r2_pk_e1024_soft.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 > gpurun_out/r2_1m_auto_soft.log 2>&1
python tools/probe.py --n 1000000 --iso soft --reps 2 --layout packed > gpurun_out/r2_1m_pk_soft.log 2>&1
python tools/probe.py --n 8400000 --iso soft --reps 1 --layout packed > gpurun_out/plain_pk.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'seirm|gather' -c 48 --csv --log-file gpurun_out/r2_launches_pk_soft.csv python tools/probe.py --n 8400000 --iso soft --reps 1 --layout packed > gpurun_out/ncu_pk.log 2>&1
grep -E "layout|t= ?[0-9]+ |episode" gpurun_out/r2_pk_*.log gpurun_out/r2_1m_*.log | cut -c1-110
