cd /root/repo
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests3.txt 2>&1
tail -4 gpurun_out/r02_final_tests3.txt
for cfg in "DLZ4_SPLIT=1" "DLZ4_SPLIT=0" "DLZ4_PW=0" "DLZ4_SPLIT=0 DLZ4_HYBRID=0"; do
  echo "== $cfg"
  env $cfg timeout 200 python divortio-lz4_b200/tools/kbench.py 256 65536 log,mixed 2>&1 | cut -c1-130
done
