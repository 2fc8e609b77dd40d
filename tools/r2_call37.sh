for v in t256c6 t512c3 t384c4 t128c12; do
  echo "== variant $v f32=1"
  python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_$v.so --f32 1 --brief 1 2>&1 | grep -E "hot steps|episode fwd|checksum|Error|error" | cut -c1-160
done > gpurun_out/r37_variants.log 2>&1
cat gpurun_out/r37_variants.log
