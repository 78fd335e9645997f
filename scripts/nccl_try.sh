run() { echo "== $*"; env "$@" FQD_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench_seq.py --mode tight --pairs 20000000 --steps 2 2>&1 | grep -E "fqd trace" | grep -o '"records all-to-all": [0-9.]*'; }
run NCCL_MIN_P2P_NCHANNELS=16
run NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run NCCL_P2P_NET_CHUNKSIZE=4194304 NCCL_MIN_P2P_NCHANNELS=32
