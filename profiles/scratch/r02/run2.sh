set -x
cd /root/repo
export DLZ4_HYBRID=0 DLZ4_WIDE=1
python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_compress_fresh16 -s 1 -c 1 -o gpurun_out/r02_wide_log128 -f python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_ncu_wide.log 2>&1
tail -3 gpurun_out/r02_ncu_wide.log
df -h /dev/shm /tmp | cat
free -g | cat
nproc
unset DLZ4_HYBRID DLZ4_WIDE
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_frames_tests.txt 2>&1
tail -5 gpurun_out/r02_frames_tests.txt
