cd /root/repo
B="--no-cpu-baseline --no-hbm-kernels --no-extra"
python bench.py --steps 2 --warmup 3 $B > gpurun_out/plain.log 2>&1 &&
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k 'regex:dense_tma|igemm_fwd_kernel' -c 6 -f -o gpurun_out/r2_fc2 python bench.py --steps 2 --warmup 3 $B --ncu-step 1 > gpurun_out/ncu_fc2.log 2>&1; echo "rc=$?"; ls -la gpurun_out/r2_fc2.ncu-rep
