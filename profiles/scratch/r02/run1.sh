set -x
cd /root/repo
./profiles/scratch/ubench/chains > gpurun_out/r02_chains.txt 2>&1
cat gpurun_out/r02_chains.txt
# correctness of the wide window on the all-shared-memory kernel
DLZ4_HYBRID=0 DLZ4_WIDE=1 timeout 900 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/r02_wide_tests.txt 2>&1
tail -15 gpurun_out/r02_wide_tests.txt
for cfg in "DLZ4_HYBRID=0 DLZ4_WIDE=0" "DLZ4_HYBRID=0 DLZ4_WIDE=1" "DLZ4_HYBRID=1"; do
  echo "== $cfg" >> gpurun_out/r02_kbench1.txt
  env $cfg timeout 600 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,rand,zero >> gpurun_out/r02_kbench1.txt 2>&1
done
cat gpurun_out/r02_kbench1.txt
