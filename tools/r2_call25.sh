for v in base age pre cb agecb all all6; do
  for f in 1 0; do
    echo "== variant $v f32=$f"
    python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_$v.so --f32 $f --brief 1 2>&1 | grep -E "hot steps|episode fwd|bwd|checksum|Error|error" | cut -c1-160
  done
done > gpurun_out/r25_variants.log 2>&1
cat gpurun_out/r25_variants.log
