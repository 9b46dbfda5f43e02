mkdir -p gpurun_out
(timeout 200 python scripts/emit_only_gpu.py 8; timeout 200 python scripts/emit_only_gpu.py 8 fastq2fasta) > gpurun_out/s3_var13.log 2>&1
cat gpurun_out/s3_var13.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q 2>&1 | tail -4) > gpurun_out/s3_tests13.log
cat gpurun_out/s3_tests13.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
