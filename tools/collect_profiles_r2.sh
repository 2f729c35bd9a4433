#!/bin/bash
# Round-2 evidence run on ONE B200 (gpurun): tests, bench (both arms, default flags), launch lists of the det and the --stochastic step,
# GEMM DRAM traffic, ncu --set full of the tcgen05 Wasserstein-attention kernels and of the det attention + GEMM kernels.
# Every ncu pass runs only after the same program has exited 0 without ncu.
set -x
O=gpurun_out/prof2
mkdir -p $O
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $O/tests_gpu.log
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err
python bench.py --stochastic --steps 10 --warmup 3 --no-cpu-baseline --no-mc --no-sub > $O/bench_stochastic_n1.json 2> $O/bench_sto.err
python tools/attn_bench.py --bwd > $O/attn_bench.log 2>&1
python tools/gemm_bench.py --quick > $O/gemm_shapes.log 2>&1
python tools/profile_step.py > $O/profile_step.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step.csv python tools/profile_step.py > /dev/null 2>&1
PROFILE_STOCHASTIC=1 python tools/profile_step.py > $O/profile_step_sto.log 2>&1 || exit 1
PROFILE_STOCHASTIC=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_step_stochastic.csv python tools/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 --csv --log-file $O/gemm_dram.csv python tools/profile_step.py > /dev/null 2>&1
PROFILE_STOCHASTIC=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wattn_fwd_kernel|wattn_prep" --launch-skip 26 --launch-count 3 -o $O/wattn_fwd_step -f python tools/profile_step.py > $O/ncu_wattn_fwd.log 2>&1
PROFILE_STOCHASTIC=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wattn_bwd_kv|wattn_bwd_dx" --launch-skip 4 --launch-count 2 -o $O/wattn_bwd_step -f python tools/profile_step.py > $O/ncu_wattn_bwd.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_fwd_sm100" --launch-skip 14 --launch-count 2 -o $O/attn_fwd_step -f python tools/profile_step.py > $O/ncu_attn_fwd.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_bwd_kv|attn_bwd_dq" --launch-skip 4 --launch-count 2 -o $O/attn_step -f python tools/profile_step.py > $O/ncu_attn.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"ln_bwd_bulk|ln_fwd_kernel" --launch-skip 20 --launch-count 2 -o $O/rows_step -f python tools/profile_step.py > $O/ncu_rows.log 2>&1
python tools/row_bench.py > $O/row_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 --launch-skip 60 --launch-count 8 -o $O/gemm_step -f python tools/profile_step.py > $O/ncu_gemm.log 2>&1
tail -2 $O/tests_gpu.log
cut -c1-300 $O/bench_n1.json
