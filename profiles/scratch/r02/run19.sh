set -x
cd /root/repo
./profiles/scratch/ubench/xxh_chain > gpurun_out/r02_xxh_chain.txt 2>&1; cat gpurun_out/r02_xxh_chain.txt
timeout 900 python -m pytest tests/test_gpu_blocks.py tests/test_gpu_frames.py tests/test_gpu_jump_decode.py -x -q -m gpu > gpurun_out/r02_pg_tests.txt 2>&1
tail -5 gpurun_out/r02_pg_tests.txt
(time timeout 900 python bench.py --steps 5 --warmup 3 --no-configs) > gpurun_out/r02_bench_new2.txt 2>&1
tail -c 3500 gpurun_out/r02_bench_new2.txt
