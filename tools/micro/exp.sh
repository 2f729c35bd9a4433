for d in 0 1 2 4 7; do
  B200VIT_ATTN_DEBUG=$d timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/exp_$d.csv python tools/attn_bench.py --ncu --bwd > /dev/null 2>&1
  echo "debug=$d: $(grep attn_bwd_kv gpurun_out/exp_$d.csv | tail -1 | awk -F'","' '{print $NF}')"
done
