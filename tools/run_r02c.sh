# Round 2, GPU call 3: warp-per-item brick k-NN + cluster/DSMEM reuse evaluation.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_lio.py tests/test_gpu_lio_config1.py tests/test_gpu_zz_pdl.py tests/test_gpu_zz_sequence.py tests/test_gpu_shard.py -m gpu -q -s > gpurun_out/r02c_pytest.log 2>&1; tail -12 gpurun_out/r02c_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02c_knn_probe.jsonl 2> gpurun_out/r02c_knn_probe.err; cat gpurun_out/r02c_knn_probe.jsonl; tail -3 gpurun_out/r02c_knn_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_ --launch-skip 4 -c 4 -f -o gpurun_out/r02c_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02c_ncu.log 2>&1; tail -3 gpurun_out/r02c_ncu.log
ncu -i gpurun_out/r02c_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02c_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02c_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02c_brick_knn_summary.txt 2>&1; cat gpurun_out/r02c_brick_knn_summary.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02c_bench_cluster.json 2> gpurun_out/r02c_bench_cluster.err; tail -c 1200 gpurun_out/r02c_bench_cluster.json; tail -3 gpurun_out/r02c_bench_cluster.err
LSD_REUSE_CLUSTER=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02c_bench_nocluster.json 2> /dev/null; tail -c 1200 gpurun_out/r02c_bench_nocluster.json
