set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r01q.log 2>&1; tail -3 gpurun_out/pytest_r01q.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01q.json 2> gpurun_out/bench_r01q.err; tail -c 600 gpurun_out/bench_r01q.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 --no-prefetch > gpurun_out/bench_r01q_noprefetch.json 2> /dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01q_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lio_knn_kernel|lio_hmodel_kernel" --launch-skip 8 -c 6 -f -o gpurun_out/r01q_lio_kernels python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/bench_under_ncu2.log 2>&1
