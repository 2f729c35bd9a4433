set -x
mkdir -p gpurun_out/prof
python tools/attn_bench.py --bwd > gpurun_out/prof/attn_bench.log 2>&1
python tools/profile_step.py > gpurun_out/prof/profile_step.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/launches_step.csv python tools/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_fwd_sm100|attn_bwd_kv|attn_bwd_dq" --launch-skip 22 --launch-count 5 -o gpurun_out/prof/attn_step -f python tools/profile_step.py > gpurun_out/prof/ncu_attn.log 2>&1
tail -4 gpurun_out/prof/attn_bench.log
