cd /root/repo
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_frames.py tests/test_gpu_jump_decode.py tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/r02_dec_tests2.txt 2>&1
tail -2 gpurun_out/r02_dec_tests2.txt
for cfg in "128 4" "64 4" "32 4" "256 4" "64 2" "96 4"; do
  set -- $cfg
  echo "== chunk MiB $1 lanes $2"
  DLZ4_CHUNK_MIB=$1 DLZ4_LANES=$2 timeout 200 python divortio-lz4_b200/tools/e2e_bench.py 1024 2>&1 | tail -3
done > gpurun_out/r02_e2e_sweep.txt 2>&1
cat gpurun_out/r02_e2e_sweep.txt
