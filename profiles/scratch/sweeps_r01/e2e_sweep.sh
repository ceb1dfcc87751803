mkdir -p gpurun_out
for cfg in "128 2" "64 4" "32 4" "16 4" "32 2" "8 4"; do
  set -- $cfg
  echo "== chunk MiB $1 lanes $2"
  DLZ4_CHUNK_MIB=$1 DLZ4_LANES=$2 timeout 200 python divortio-lz4_b200/tools/e2e_bench.py 1024 2>&1 | tail -3
done > gpurun_out/e2e_sweep.log 2>&1
cat gpurun_out/e2e_sweep.log
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
