set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 420 $TR --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_n4_replicas.json 2> gpurun_out/bench_n4_replicas.err; tail -c 1200 gpurun_out/bench_n4_replicas.json; tail -3 gpurun_out/bench_n4_replicas.err
timeout 300 $TR --master-port 29512 bench_extra.py --which gicp --gicp-pairs 32 --no-cpu > gpurun_out/extra_gicp_n4.log 2> gpurun_out/extra_gicp_n4.err; tail -1 gpurun_out/extra_gicp_n4.log; tail -3 gpurun_out/extra_gicp_n4.err
timeout 420 $TR --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 --mg-mode shard > gpurun_out/bench_n4_shard.json 2> gpurun_out/bench_n4_shard.err; tail -c 1200 gpurun_out/bench_n4_shard.json; tail -3 gpurun_out/bench_n4_shard.err
