#!/bin/bash
# Round-1 evidence run on ONE B200 (gpurun): tests, bench (both arms), launch list, GEMM dram traffic, ncu --set full of the attention + GEMM kernels.
set -x
mkdir -p gpurun_out/prof
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/prof/tests_gpu.log
python bench.py > gpurun_out/prof/bench_n1.json 2> gpurun_out/prof/bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/prof/bench_reference_arm.json 2> gpurun_out/prof/bench_ref.err
python bench.py --stochastic --steps 10 --warmup 3 --no-cpu-baseline --no-mc > gpurun_out/prof/bench_stochastic_n1.json 2> gpurun_out/prof/bench_sto.err
python tools/attn_bench.py --bwd > gpurun_out/prof/attn_bench.log 2>&1
python tools/gemm_bench.py > gpurun_out/prof/gemm_shapes.log 2>&1
python tools/aux_bench.py > gpurun_out/prof/aux_bench.json 2> gpurun_out/prof/aux_bench.err          # Mixup / CutMix, uint8 normalise, mask generator
python tools/bench_finetune.py --mixup > gpurun_out/prof/finetune_large_n1_mixup.json 2> gpurun_out/prof/finetune.err
# every ncu pass below runs only after the same program exited 0 without ncu
python tools/profile_step.py > gpurun_out/prof/profile_step.log 2>&1 || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/launches_step.csv python tools/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16 --csv --log-file gpurun_out/prof/gemm_dram.csv python tools/profile_step.py > /dev/null 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attn_fwd_sm100|attn_bwd_kv|attn_bwd_dq" --launch-skip 36 --launch-count 4 -o gpurun_out/prof/attn_step -f python tools/profile_step.py > gpurun_out/prof/ncu_attn.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_bf16 --launch-skip 60 --launch-count 8 -o gpurun_out/prof/gemm_step -f python tools/profile_step.py > gpurun_out/prof/ncu_gemm.log 2>&1
tail -2 gpurun_out/prof/tests_gpu.log
cut -c1-300 gpurun_out/prof/bench_n1.json
