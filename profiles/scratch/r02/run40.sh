cd /root/repo
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests2.txt 2>&1
tail -4 gpurun_out/r02_final_tests2.txt
timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed 2>&1 | cut -c1-130
timeout 300 python divortio-lz4_b200/tools/stream_bench.py 512 64 2>&1 | tail -4
