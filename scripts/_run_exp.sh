mkdir -p gpurun_out
python bench.py > gpurun_out/s3_bench_1gpu.json 2> gpurun_out/s3_bench_1gpu.err
tail -c 300 gpurun_out/s3_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s3_launches_bench_16gib.csv python bench.py --steps 2 --warmup 1 --no-e2e > gpurun_out/s3_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k3_fwd|k4_emit' -s 2 -c 2 -o gpurun_out/s3_k4_full_16gib -f python scripts/emit_only_gpu.py 16 > gpurun_out/s3_k4_full.log 2>&1
tail -3 gpurun_out/s3_k4_full.log
cat gpurun_out/s3_bench_1gpu.json | cut -c1-400
