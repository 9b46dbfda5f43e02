mkdir -p gpurun_out
( echo "== memcheck csv2json"; timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_v4_gpu.py csv2json 2>&1 | tail -6
  echo "== racecheck csv2json"; timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_v4_gpu.py csv2json 2>&1 | tail -6
  echo "== synccheck csv2json fastq2fasta"; timeout 600 compute-sanitizer --tool synccheck python scripts/sanitize_v4_gpu.py 2>&1 | tail -6
) > gpurun_out/s3_sanitizer.txt 2>&1
cat gpurun_out/s3_sanitizer.txt
