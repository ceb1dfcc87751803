cd /root/repo
for cfg in "64 2" "32 2" "48 2" "64 1" "64 3" "32 3" "96 2" "128 2"; do
  set -- $cfg
  echo "== chunk MiB $1 lanes $2"
  DLZ4_CHUNK_MIB=$1 DLZ4_LANES=$2 timeout 200 python divortio-lz4_b200/tools/e2e_bench.py 1024 2>&1 | tail -2
done > gpurun_out/r02_e2e_sweep2.txt 2>&1
cat gpurun_out/r02_e2e_sweep2.txt
