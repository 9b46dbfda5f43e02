mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "oracle_code" 2>&1 | tail -25) > gpurun_out/s3_tests2.log
cat gpurun_out/s3_tests2.log
