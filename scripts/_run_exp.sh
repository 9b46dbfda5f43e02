mkdir -p gpurun_out
(timeout 200 python scripts/emit_only_gpu.py 8; timeout 200 python scripts/emit_only_gpu.py 4 iso_datetime_to_json; timeout 200 python scripts/emit_only_gpu.py 8 fastq2fasta; timeout 200 python scripts/emit_only_gpu.py 4 add-commas) > gpurun_out/s3_var8.log 2>&1
cat gpurun_out/s3_var8.log
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) > gpurun_out/s3_tests8.log
cat gpurun_out/s3_tests8.log
