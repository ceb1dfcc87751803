mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jump_decode.py -m gpu -x -q 2>&1 | tail -25
(timeout 300 python divortio-lz4_b200/tools/frame_bench.py log 64; timeout 600 python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --no-cpu) > gpurun_out/frame_bench2.log 2>&1
cat gpurun_out/frame_bench2.log
