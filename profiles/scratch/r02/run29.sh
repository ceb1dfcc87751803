cd /root/repo
for pw in 2 3; do
DLZ4_PW=$pw DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so timeout 300 python divortio-lz4_b200/tools/pw_phases.py log 1024
done > gpurun_out/r02_pw_phases5.txt 2>&1
cat gpurun_out/r02_pw_phases5.txt
