set -x
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_s3_tests.txt 2>&1
tail -5 gpurun_out/r02_s3_tests.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_s3_bench.txt 2>&1
tail -2 gpurun_out/r02_s3_bench.txt
