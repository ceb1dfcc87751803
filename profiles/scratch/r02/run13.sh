set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/r02_pw_tests.txt 2>&1
tail -3 gpurun_out/r02_pw_tests.txt
grep -q passed gpurun_out/r02_pw_tests.txt || exit 1
timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand > gpurun_out/r02_pw_kbench7.txt 2>&1
cat gpurun_out/r02_pw_kbench7.txt
DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so timeout 300 python divortio-lz4_b200/tools/pw_phases.py log 1024 > gpurun_out/r02_pw_phases.txt 2>&1
DLZ4_LIB=divortio-lz4_b200/csrc/libdlz4_b200_prof.so timeout 300 python divortio-lz4_b200/tools/pw_phases.py mixed 1024 >> gpurun_out/r02_pw_phases.txt 2>&1
cat gpurun_out/r02_pw_phases.txt
free -g; df -h /dev/shm /tmp; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv
python divortio-lz4_b200/tools/prof_one.py log 1024 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_parse_pw -s 1 -c 1 -o gpurun_out/r02_pw5_log1024 -f python divortio-lz4_b200/tools/prof_one.py log 1024 > gpurun_out/r02_ncu_pw.log 2>&1
tail -3 gpurun_out/r02_ncu_pw.log
