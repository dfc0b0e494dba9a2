# usage: N=<gpus> TAG=<tag> bash tools/run_scale.sh
set -x
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
grep -c metric gpurun_out/${TAG}_bench_n$N.json
