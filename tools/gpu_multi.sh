# N-GPU run of bench.py (our arm) the way the driver launches it; usage: bash tools/gpu_multi.sh N [extra bench args]
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@"
