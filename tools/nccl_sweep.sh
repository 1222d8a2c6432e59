# 2-GPU experiments on where the data-parallel overhead comes from (run under gpurun --gpus 2)
run() { echo "== $1 | $2"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; }
echo "== single GPU"; python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
run "A=1" "--stages-per-group 4" 29541
run "ADP_TC_SMS=140" "--stages-per-group 4" 29542
run "ADP_TC_SMS=132" "--stages-per-group 4" 29543
run "ADP_TC_SMS=140 ADP_TC_DYNAMIC=1" "--stages-per-group 4" 29544
run "ADP_TC_SMS=132 NCCL_MAX_CTAS=16" "--stages-per-group 4" 29545
run "NCCL_MIN_CTAS=32" "--stages-per-group 4" 29546
