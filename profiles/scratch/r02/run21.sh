set -x
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_s4_tests.txt 2>&1
tail -8 gpurun_out/r02_s4_tests.txt
timeout 600 python divortio-lz4_b200/tools/config_bench.py 1024 262144 > gpurun_out/r02_config_bench.txt 2>&1
cat gpurun_out/r02_config_bench.txt
