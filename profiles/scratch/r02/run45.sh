cd /root/repo
(time timeout 900 python bench.py --steps 5 --warmup 3) > gpurun_out/r02_bench_final_n1.txt 2>&1
tail -c 600 gpurun_out/r02_bench_final_n1.txt
python -c "import __graft_entry__ as g; g.smoke()"
