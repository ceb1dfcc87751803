mkdir -p gpurun_out
(timeout 300 python divortio-lz4_b200/tools/frame_bench.py log 64; timeout 600 python divortio-lz4_b200/tools/frame_bench.py mixed 1024 --no-cpu) > gpurun_out/frame_bench1.log 2>&1
cat gpurun_out/frame_bench1.log
