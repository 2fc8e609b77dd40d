for v in stat8 stat6 dyn8 dyn6 dyn5 dyn4; do
  echo "== variant $v f32=1"
  python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_$v.so --f32 1 --brief 1 2>&1 | grep -E "hot steps|episode fwd|bwd|checksum|Error|error" | cut -c1-160
done > gpurun_out/r26_variants.log 2>&1
echo "== variant dyn6 tile 1024" >> gpurun_out/r26_variants.log
python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_dyn6.so --f32 1 --brief 1 --tile 1024 2>&1 | grep -E "hot steps|episode fwd|bwd|checksum|Error|error" | cut -c1-160 >> gpurun_out/r26_variants.log
echo "== variant dyn6 hard" >> gpurun_out/r26_variants.log
python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_dyn6.so --f32 1 --brief 1 --iso hard 2>&1 | grep -E "hot steps|episode fwd|bwd|checksum|Error|error" | cut -c1-160 >> gpurun_out/r26_variants.log
echo "== variant stat8 hard" >> gpurun_out/r26_variants.log
python tools/probe.py --n 8400000 --lib agenttorch_b200/lib/variants/libat_seirm_stat8.so --f32 1 --brief 1 --iso hard 2>&1 | grep -E "hot steps|episode fwd|bwd|checksum|Error|error" | cut -c1-160 >> gpurun_out/r26_variants.log
grep -E "variant|hot steps|checksum" gpurun_out/r26_variants.log
