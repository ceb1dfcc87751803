cd /root/repo
for cfg in "DLZ4_PW_SLEEP=200" "DLZ4_PW_SLEEP=400" "DLZ4_PW_SLEEP=800" "DLZ4_PW_SLEEP=1500" "DLZ4_PW_SLEEP=100"; do
  echo "== $cfg"
  env $cfg timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | cut -c1-100
done > gpurun_out/r02_pw_sleep.txt 2>&1
cat gpurun_out/r02_pw_sleep.txt
