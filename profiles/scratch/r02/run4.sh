set -x
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/r02_split_tests.txt 2>&1
tail -15 gpurun_out/r02_split_tests.txt
timeout 600 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,rand,zero > gpurun_out/r02_kbench3.txt 2>&1
cat gpurun_out/r02_kbench3.txt
python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_parse_fresh16\|k_encode_blocks -s 2 -c 2 -o gpurun_out/r02_split_log128 -f python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_ncu_split.log 2>&1
tail -3 gpurun_out/r02_ncu_split.log
