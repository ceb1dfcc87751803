set -x
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_frames.py tests/test_gpu_stream.py -x -q -m gpu > gpurun_out/r02_dec_tests.txt 2>&1
tail -3 gpurun_out/r02_dec_tests.txt
grep -q passed gpurun_out/r02_dec_tests.txt || exit 1
timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,zero,rand,bench > gpurun_out/r02_dec_kbench1.txt 2>&1
cat gpurun_out/r02_dec_kbench1.txt
