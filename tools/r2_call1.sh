set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
for iso in soft hard; do
python tools/probe.py --n 8400000 --iso $iso --reps 2 > gpurun_out/r2_p_hashed_$iso.log 2>&1
for blk in 1050624 840704 700416; do
python tools/probe.py --n 8400000 --iso $iso --reps 2 --block $blk > gpurun_out/r2_p_blk${blk}_$iso.log 2>&1
done
done
tail -n 30 gpurun_out/r2_p_*.log
