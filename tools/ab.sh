# same-box A/B of an environment switch: tools/ab.sh VAR  -> runs bench (main step only) with VAR=0 and VAR=1, twice each, interleaved
v=$1
for rep in 1 2; do for val in 0 1; do
  env $v=$val python bench.py --steps 20 --warmup 5 --no-mc ${AB_SUB:---no-sub} --no-cpu-baseline > gpurun_out/ab_${v}_${val}_$rep.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/ab_${v}_${val}_$rep.json')); print('$v=$val rep $rep', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1), round(d['roofline']['frac'],4))"
done; done
