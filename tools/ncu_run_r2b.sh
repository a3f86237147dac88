# round 2, second half: captures of the CTA-pair FFN kernel (in the C2 step) and of the split-softmax attention_pp (C4 shapes),
# plus the launch list of the C2 step.  Run under gpurun; summaries go to profiles/ via tools/ncu_metrics.py.
set -x
cd $GRAFT_REPO_ROOT
CFM_B200_CUDA_GRAPHS=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:ffn_fused_kernel -s 26 -c 2 -f -o gpurun_out/r2b_ncu_ffn_fused_pair python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2b_ncu_ffn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_pp_kernel -s 1 -c 1 -f -o gpurun_out/r2b_ncu_attention_pp python tools/prof_kernels.py attn_long > gpurun_out/r2b_ncu_attn.log 2>&1
CFM_B200_CUDA_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches.csv python bench.py --profile --steps 2 --warmup 3 > gpurun_out/r2b_launch_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
