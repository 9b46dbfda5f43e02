mkdir -p gpurun_out
for v in base W Wseq base W; do
  echo "variant $v"
  KEX_LIB=/root/repo/kleenexlang_b200/exp/libkexcuda_$v.so timeout 200 python scripts/emit_only_gpu.py 8
done > gpurun_out/s3_var2.log 2>&1
cat gpurun_out/s3_var2.log
