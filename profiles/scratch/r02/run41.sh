cd /root/repo
timeout 300 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -3 | cut -c1-200
for cfg in "DLZ4_PW=2 DLZ4_PW_LEAD=7" "DLZ4_PW=3 DLZ4_PW_LEAD=8" "DLZ4_PW=2 DLZ4_PW_LEAD=9"; do
  echo "== $cfg"
  env $cfg timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand 2>&1 | cut -c1-100
done > gpurun_out/r02_pw_scout.txt 2>&1
DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so timeout 120 python divortio-lz4_b200/tools/pw_phases.py log 1024 >> gpurun_out/r02_pw_scout.txt 2>&1
cat gpurun_out/r02_pw_scout.txt
