python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n8.json
cut -c1-330 gpurun_out/bench_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 4 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_n4.json
cut -c1-330 gpurun_out/bench_n4.json
