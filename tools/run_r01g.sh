set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01g.log 2>&1; tail -5 gpurun_out/pytest_r01g.log
timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_g32.log 2>&1; tail -1 gpurun_out/knn_probe_g32.log
LSD_L2_FETCH_GRANULARITY=128 timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_g128.log 2>&1; tail -1 gpurun_out/knn_probe_g128.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; tail -c 1500 gpurun_out/bench_r01g.json
LSD_L2_FETCH_GRANULARITY=128 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch > gpurun_out/bench_r01g_128.json 2> gpurun_out/bench_r01g_128.err; tail -c 1500 gpurun_out/bench_r01g_128.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_query_kernel -c 2 -f -o gpurun_out/knn_batch_r01g python tools/knn_probe.py 1048576 > gpurun_out/knn_probe_ncu.log 2>&1
