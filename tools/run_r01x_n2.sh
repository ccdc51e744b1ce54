set -x
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 200 python -m pytest tests/test_gpu_map_knn.py -q -k "box_delete or bit_exact" > gpurun_out/pytest_r01x.log 2>&1; tail -6 gpurun_out/pytest_r01x.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 --streams 0 > gpurun_out/bench_n2_x.json 2> gpurun_out/bench_n2_x.err
head -c 120 gpurun_out/bench_n2_x.json; echo; wc -l gpurun_out/bench_n2_x.json
