# Round 2, GPU call 10 (8 GPUs): the multi-GPU legs at N = 8.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv | sort | uniq -c
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r02j_bench_n8.json 2> gpurun_out/r02j_bench_n8.err; tail -c 2000 gpurun_out/r02j_bench_n8.json; tail -5 gpurun_out/r02j_bench_n8.err
timeout 900 $TR --master-port 29522 bench_extra.py --which gicp --gicp-pairs 256 > gpurun_out/r02j_extra_gicp_n8.jsonl 2> gpurun_out/r02j_extra_gicp_n8.err; cat gpurun_out/r02j_extra_gicp_n8.jsonl; tail -3 gpurun_out/r02j_extra_gicp_n8.err
timeout 900 $TR --master-port 29523 bench_extra.py --which ndt > gpurun_out/r02j_extra_ndt_n8.jsonl 2> gpurun_out/r02j_extra_ndt_n8.err; cat gpurun_out/r02j_extra_ndt_n8.jsonl; tail -5 gpurun_out/r02j_extra_ndt_n8.err
