set -x
cd /root/repo
export DLZ4_HYBRID=0 DLZ4_WIDE=1
timeout 600 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/r02_wide_tests.txt 2>&1
tail -3 gpurun_out/r02_wide_tests.txt
timeout 600 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed > gpurun_out/r02_kbench2.txt 2>&1
cat gpurun_out/r02_kbench2.txt
python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16 -s 1 -c 1 -o gpurun_out/r02_wide2_log128 -f python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_ncu_wide.log 2>&1
tail -3 gpurun_out/r02_ncu_wide.log
