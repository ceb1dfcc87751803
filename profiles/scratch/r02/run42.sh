cd /root/repo
for v in c5 c4; do for cfg in "DLZ4_PW=2" "DLZ4_PW=3 DLZ4_PW_LEAD=8"; do
  echo "== chains $v $cfg"
  env $cfg DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_$v.so timeout 120 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | cut -c1-100
done; done > gpurun_out/r02_pw_chains.txt 2>&1
cat gpurun_out/r02_pw_chains.txt
