# round 2, second half: textual evidence kept under profiles/ (clock traces, host/device profiles, micro-benchmarks)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out/ev
CFM_B200_FFN_PAIR=0 timeout 200 python tools/ffn_chain_trace.py > gpurun_out/ev/r2b_ffn_chain_trace_single.txt 2>&1
timeout 200 python tools/ffn_chain_trace.py > gpurun_out/ev/r2b_ffn_chain_trace_pair.txt 2>&1
timeout 200 python tools/attn_pp_trace.py > gpurun_out/ev/r2b_attn_pp_trace.txt 2>&1
timeout 300 python tools/opt_step_profile.py > gpurun_out/ev/r2b_opt_step_profile.txt 2>&1
timeout 300 python tools/train_kernel_diff.py > gpurun_out/ev/r2b_train_kernels.txt 2>&1
timeout 300 python tools/frontend_bench.py > gpurun_out/ev/r2b_frontend_bench.txt 2>&1
for v in 0 1; do CFM_B200_GEMM_PAIR=$v timeout 200 python tools/gemm_shapes_bench.py; done > gpurun_out/ev/r2b_gemm_shapes_bench.txt 2>&1
timeout 200 python tools/wgrad_splits_bench.py > gpurun_out/ev/r2b_wgrad_splits_bench.txt 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:subsample_fused_kernel -s 1 -c 1 -f -o gpurun_out/r2b_ncu_subsample_fused python tools/prof_kernels.py frontend > gpurun_out/ev/ncu_fe.log 2>&1
ls gpurun_out/ev
