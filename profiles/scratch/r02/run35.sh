cd /root/repo
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -2
for cfg in "DLZ4_PW_ORDER=1" "DLZ4_PW_ORDER=0"; do
  echo "== $cfg"
  env $cfg timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand 2>&1 | cut -c1-100
done > gpurun_out/r02_pw_order.txt 2>&1
cat gpurun_out/r02_pw_order.txt
