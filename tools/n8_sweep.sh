# all-reduce overlap variants at N GPUs (main step only): tools/n8_sweep.sh N
N=${1:-8}
run() {
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus $N --steps 20 --warmup 5 --no-mc --no-sub --no-cpu-baseline 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', round(d['value'],1), round(d['ms_per_step'],3), round(d['e2e']['value'],1))"
}
run B200VIT_AR_OVERLAP=0
run B200VIT_AR_OVERLAP=1 B200VIT_AR_CUT=2
run B200VIT_AR_OVERLAP=1 B200VIT_AR_CUT=4
run B200VIT_AR_OVERLAP=1 B200VIT_AR_CUT=6
run B200VIT_AR_OVERLAP=1 B200VIT_AR_CUT=2
run B200VIT_AR_OVERLAP=0
