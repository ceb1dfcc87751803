cd /root/repo
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/r02_stream_tests2.txt 2>&1
tail -15 gpurun_out/r02_stream_tests2.txt
timeout 600 python divortio-lz4_b200/tools/stream_bench.py 512 64 > gpurun_out/r02_stream_bench.txt 2>&1
cat gpurun_out/r02_stream_bench.txt
