set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_n8_replicas.json 2> gpurun_out/bench_n8_replicas.err; tail -c 600 gpurun_out/bench_n8_replicas.json; tail -2 gpurun_out/bench_n8_replicas.err
timeout 300 $TR --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 3 --mg-mode shard > gpurun_out/bench_n8_shard.json 2> gpurun_out/bench_n8_shard.err; tail -c 600 gpurun_out/bench_n8_shard.json; tail -2 gpurun_out/bench_n8_shard.err
