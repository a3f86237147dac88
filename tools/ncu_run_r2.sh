set -x
cd $GRAFT_REPO_ROOT
for k in mhsa_fused_kernel conv_fused_kernel ffn_fused_kernel; do
  CFM_B200_CUDA_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 26 -c 2 -f -o gpurun_out/r2_ncu_$k python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2_ncu_$k.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:subsample_fused_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_subsample_fused_kernel python tools/prof_kernels.py frontend > gpurun_out/r2_ncu_fe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_ncu_ctc_gemm_tc_argmax python tools/prof_kernels.py ctc > gpurun_out/r2_ncu_ctc.log 2>&1
CFM_B200_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2_launch_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
