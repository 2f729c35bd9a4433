run() { # name envs...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-mc --no-sub --no-cpu-baseline > gpurun_out/n2_$name.json 2> gpurun_out/n2_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/n2_$name.json')); print('$name', round(d['value'],1), round(d['ms_per_step'],3), d['final_loss'])" || tail -3 gpurun_out/n2_$name.err
}
run off B200VIT_AR_OVERLAP=0
run cut1r0 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=0 B200VIT_AR_CUT=1
run cut1r16 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=16 B200VIT_AR_CUT=1
run cut1r32 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=32 B200VIT_AR_CUT=1
run cut2r0 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=0 B200VIT_AR_CUT=2
run cut2r24 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=24 B200VIT_AR_CUT=2
run cut3r0 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=0 B200VIT_AR_CUT=3
run cut8r0 B200VIT_AR_OVERLAP=1 B200VIT_AR_SM_RESERVE=0 B200VIT_AR_CUT=8
