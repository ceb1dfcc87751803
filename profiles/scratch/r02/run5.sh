set -x
cd /root/repo
nvidia-smi -L
timeout 120 python divortio-lz4_b200/tools/prof_one.py log 16 > gpurun_out/r02_pw_first.txt 2>&1; echo rc=$? >> gpurun_out/r02_pw_first.txt
cat gpurun_out/r02_pw_first.txt
grep -q "^ok" gpurun_out/r02_pw_first.txt || exit 1
timeout 120 python divortio-lz4_b200/tools/prof_one.py mixed 64 >> gpurun_out/r02_pw_first.txt 2>&1; echo rc=$? >> gpurun_out/r02_pw_first.txt
tail -3 gpurun_out/r02_pw_first.txt
timeout 900 python -m pytest tests/test_gpu_blocks.py -x -q -m gpu > gpurun_out/r02_pw_tests.txt 2>&1
tail -15 gpurun_out/r02_pw_tests.txt
for cfg in "DLZ4_PW=3 DLZ4_PW_LEAD=6" "DLZ4_PW=3 DLZ4_PW_LEAD=4" "DLZ4_PW=3 DLZ4_PW_LEAD=8" "DLZ4_PW=2 DLZ4_PW_LEAD=6" "DLZ4_PW=2 DLZ4_PW_LEAD=4"; do
  echo "== $cfg" >> gpurun_out/r02_pw_kbench1.txt
  env $cfg timeout 300 python divortio-lz4_b200/tools/kbench.py 1024 65536 log,mixed,rand,zero >> gpurun_out/r02_pw_kbench1.txt 2>&1
done
cat gpurun_out/r02_pw_kbench1.txt
