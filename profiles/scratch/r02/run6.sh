set -x
cd /root/repo
export DLZ4_PW=${DLZ4_PW:-2} DLZ4_PW_LEAD=${DLZ4_PW_LEAD:-4}
python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_prof_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:k_parse_pw -s 1 -c 1 -o gpurun_out/r02_pw_log128 -f python divortio-lz4_b200/tools/prof_one.py log 128 > gpurun_out/r02_ncu_pw.log 2>&1
tail -3 gpurun_out/r02_ncu_pw.log
