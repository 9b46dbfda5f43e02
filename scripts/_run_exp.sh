mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6) > gpurun_out/s3_tests12.log
cat gpurun_out/s3_tests12.log
