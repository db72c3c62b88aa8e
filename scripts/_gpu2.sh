P=29511
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $P scripts/run_strips.py --native --transport $2 --reps 200 $3 2>&1 | grep -E '^\{|Error|error' | tail -3 | cut -c1-400; P=$((P+1)); }
timeout 300 python -m pytest tests/test_dist.py -x -q -m gpu -k "peer_memory" 2>&1 | tail -5
run $NG peer --check
run $NG peer "--check --graph"
run $NG nccl ""
