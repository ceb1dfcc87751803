set -x
cd /root/repo
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r02_final_tests.txt 2>&1
tail -4 gpurun_out/r02_final_tests.txt
(time timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/r02_bench_final_n1.txt 2>&1
tail -c 1500 gpurun_out/r02_bench_final_n1.txt
(time timeout 600 python bench.py --impl reference --steps 5 --warmup 3) > gpurun_out/r02_bench_final_ref.txt 2>&1
tail -c 1200 gpurun_out/r02_bench_final_ref.txt
python divortio-lz4_b200/tools/prof_one.py mixed 1024 > gpurun_out/r02_prof_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:k_parse_pw -s 1 -c 1 -o gpurun_out/r02_bench_compress -f python divortio-lz4_b200/tools/prof_one.py mixed 1024 > gpurun_out/r02_ncu_c.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_decompress_blocks -s 1 -c 1 -o gpurun_out/r02_bench_decompress -f python divortio-lz4_b200/tools/prof_one.py mixed 1024 > gpurun_out/r02_ncu_d.log 2>&1
python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r02_bench_for_launches.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_l.log 2>&1
tail -3 gpurun_out/r02_launches.csv
