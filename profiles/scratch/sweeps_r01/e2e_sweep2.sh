mkdir -p gpurun_out
for cfg in "128 4" "64 4" "64 8" "32 4" "32 8" "16 8"; do
  set -- $cfg
  echo "== chunk MiB $1 lanes $2"
  DLZ4_CHUNK_MIB=$1 DLZ4_LANES=$2 timeout 200 python divortio-lz4_b200/tools/e2e_bench.py 1024 2>&1 | tail -2
done > gpurun_out/e2e_sweep2.log 2>&1
cat gpurun_out/e2e_sweep2.log
