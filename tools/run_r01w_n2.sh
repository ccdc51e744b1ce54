set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_w.json 2> gpurun_out/bench_n2_w.err
head -c 200 gpurun_out/bench_n2_w.json; echo; wc -l gpurun_out/bench_n2_w.json; grep -c "NCCL version" gpurun_out/bench_n2_w.err
