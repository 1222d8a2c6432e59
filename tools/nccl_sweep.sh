run() { echo "== $1 | $2"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; }
run "A=1" "" 29541
run "NCCL_MAX_CTAS=8" "" 29542
run "NCCL_MAX_CTAS=4" "" 29543
run "A=1" "--stages-per-group 4" 29544
run "A=1" "--stages-per-group 1" 29545
run "NCCL_MAX_CTAS=8" "--stages-per-group 4" 29546
