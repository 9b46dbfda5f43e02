mkdir -p gpurun_out
(for i in 1 2; do timeout 200 python scripts/emit_only_gpu.py 8; done; timeout 200 python scripts/emit_only_gpu.py 8 fastq2fasta) > gpurun_out/s3_var9.log 2>&1
cat gpurun_out/s3_var9.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "synthetic or golden or reject or forced or shard or v4 or monoid" 2>&1 | tail -4) > gpurun_out/s3_tests9.log
cat gpurun_out/s3_tests9.log
