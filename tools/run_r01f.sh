set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err
tail -c 3000 gpurun_out/bench_r01f.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_query_kernel -c 3 -f -o gpurun_out/knn_batch_r01f python tools/knn_probe.py 1048576 > gpurun_out/knn_probe.log 2>&1
tail -5 gpurun_out/knn_probe.log
