# 2-GPU experiments on where the data-parallel overhead comes from (run under gpurun --gpus 2; results in DESIGN.md section 6)
run() { echo "== $1 | $2"; env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"; }
echo "== single GPU"; python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
run "A=1" "--stages-per-group 2" 29541
run "A=1" "--stages-per-group 4" 29542
run "A=1" "--stages-per-group 16" 29543            # one all-reduce at the end: no overlap
run "ADP_SKIP_GRAD_ALLREDUCE=1" "" 29544           # everything but the gradient all-reduce (wrong training, timing only)
run "ADP_SIDE_STREAM=0" "" 29545                   # weight gradients back on the main stream
run "ADP_TC_SMS=132" "" 29546                      # leave 16 SMs to NCCL
run "NCCL_MAX_CTAS=8" "" 29547
