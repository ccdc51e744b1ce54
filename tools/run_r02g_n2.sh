# Round 2, GPU call 7 (2 GPUs): brick k-NN v6 probe, LRU test on hardware, the N = 2 multi-GPU legs.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv
timeout 900 python -m pytest tests/test_gpu_lru.py tests/test_gpu_brick_knn.py tests/test_gpu_shard.py tests/test_gpu_zz_fastlio_seam.py -m gpu -q > gpurun_out/r02g_pytest.log 2>&1; tail -8 gpurun_out/r02g_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02g_knn_probe.jsonl 2> gpurun_out/r02g_knn_probe.err; cat gpurun_out/r02g_knn_probe.jsonl; tail -3 gpurun_out/r02g_knn_probe.err
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/r02g_bench_n2.err; tail -c 1800 gpurun_out/r02g_bench_n2.json; tail -5 gpurun_out/r02g_bench_n2.err
timeout 900 $TR --master-port 29512 bench_extra.py --which ndt > gpurun_out/r02g_extra_ndt_n2.jsonl 2> gpurun_out/r02g_extra_ndt_n2.err; cat gpurun_out/r02g_extra_ndt_n2.jsonl; tail -5 gpurun_out/r02g_extra_ndt_n2.err
timeout 900 $TR --master-port 29513 bench_extra.py --which gicp --gicp-pairs 32 > gpurun_out/r02g_extra_gicp_n2.jsonl 2> gpurun_out/r02g_extra_gicp_n2.err; cat gpurun_out/r02g_extra_gicp_n2.jsonl; tail -3 gpurun_out/r02g_extra_gicp_n2.err
