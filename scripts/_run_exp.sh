mkdir -p gpurun_out
(timeout 200 python scripts/emit_only_gpu.py 4 iso_datetime_to_json; KEX_V3_NORMW=1 timeout 200 python scripts/emit_only_gpu.py 4 iso_datetime_to_json; KEX_NO_V4=1 timeout 200 python scripts/emit_only_gpu.py 8; KEX_NO_V4=1 KEX_V3_NORMW=1 timeout 200 python scripts/emit_only_gpu.py 8) > gpurun_out/s3_var10.log 2>&1
cat gpurun_out/s3_var10.log
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "v3_paths or v4_paths or synthetic or reject or golden" 2>&1 | tail -4) > gpurun_out/s3_tests10.log
cat gpurun_out/s3_tests10.log
