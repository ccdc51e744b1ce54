set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01k.log 2>&1; tail -15 gpurun_out/pytest_r01k.log
timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_k.log 2>&1; tail -1 gpurun_out/knn_probe_k.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01k.json 2> gpurun_out/bench_r01k.err; tail -c 2500 gpurun_out/bench_r01k.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch --no-prefetch > gpurun_out/bench_r01k_noprefetch.json 2> gpurun_out/bench_r01k_noprefetch.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_query_thread -c 1 -f -o gpurun_out/knn_thread_r01k python tools/knn_probe.py 1048576 > gpurun_out/knn_probe_ncu_k.log 2>&1
timeout 600 python bench_extra.py --which gicp --gicp-pairs 4 > gpurun_out/extra_gicp_k.log 2>&1; tail -1 gpurun_out/extra_gicp_k.log
timeout 600 python bench_extra.py --which gicp --gicp-pairs 4 --gicp-method FAST_VGICP > gpurun_out/extra_vgicp_k.log 2>&1; tail -1 gpurun_out/extra_vgicp_k.log
