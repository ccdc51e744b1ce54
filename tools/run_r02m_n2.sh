# Round 2, GPU call 13 (2 GPUs): the multi-GPU legs at N = 2 on the final state (exchange counter over the timed steps only).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_bench_n2.json 2> gpurun_out/r02m_bench_n2.err; tail -c 1800 gpurun_out/r02m_bench_n2.json; tail -5 gpurun_out/r02m_bench_n2.err
timeout 600 $TR --master-port 29532 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02m_bench_ref_n2.json 2> gpurun_out/r02m_bench_ref_n2.err; tail -c 300 gpurun_out/r02m_bench_ref_n2.json
timeout 600 $TR --master-port 29533 bench_extra.py --which ndt > gpurun_out/r02m_extra_ndt_n2.jsonl 2> gpurun_out/r02m_extra_ndt_n2.err; cut -c1-1200 gpurun_out/r02m_extra_ndt_n2.jsonl; tail -3 gpurun_out/r02m_extra_ndt_n2.err
timeout 300 python -m pytest tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -3
