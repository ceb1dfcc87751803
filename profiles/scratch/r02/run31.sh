cd /root/repo
DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_cap32.so timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu 2>&1 | tail -2
for cfg in "DLZ4_PW=2 DLZ4_PW_LEAD=7" "DLZ4_PW=2 DLZ4_PW_LEAD=6" "DLZ4_PW=2 DLZ4_PW_LEAD=5" "DLZ4_PW=3 DLZ4_PW_LEAD=7"; do
  echo "== cap32 $cfg"
  env $cfg DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_cap32.so timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | cut -c1-100
done > gpurun_out/r02_pw_cap32.txt 2>&1
DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_cap32p.so timeout 300 python divortio-lz4_b200/tools/pw_phases.py log 1024 >> gpurun_out/r02_pw_cap32.txt 2>&1
cat gpurun_out/r02_pw_cap32.txt
