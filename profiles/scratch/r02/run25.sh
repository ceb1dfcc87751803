set -x
cd /root/repo
rm -f gpurun_out/r02_dec_kbench2.txt
for v in "" _vB _vC _vD; do
  echo "== variant '$v' (''=minb6 q2, vB=minb1 q2, vC=minb5 q4, vD=minb5 q3)" >> gpurun_out/r02_dec_kbench2.txt
  DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200$v.so timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand,bench 2>&1 | cut -c1-20,75- >> gpurun_out/r02_dec_kbench2.txt
done
cat gpurun_out/r02_dec_kbench2.txt
