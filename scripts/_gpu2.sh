P=29511
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P scripts/run_strips.py --native --transport $2 --reps 200 $3 2>&1 | grep -E '^\{|Error|error' | tail -3 | cut -c1-400; P=$((P+1)); }
run 8 peer "--check --graph"
run 8 peer ""
run 8 nccl ""
run 4 peer "--check --graph"
