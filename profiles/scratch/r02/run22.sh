set -x
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_s4_tests.txt 2>&1
tail -8 gpurun_out/r02_s4_tests.txt
