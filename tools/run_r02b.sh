# Round 2, GPU call 2:  brick k-NN on hardware (tests, timing, ncu), the VGICP comparator, the bench pair with the sweep fix.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_zz_ref_cuda_vgicp.py tests/test_gpu_voxelgrid.py tests/test_gpu_map_knn.py -m gpu -q -s > gpurun_out/r02b_pytest.log 2>&1; tail -30 gpurun_out/r02b_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02b_knn_probe.jsonl 2> gpurun_out/r02b_knn_probe.err; cat gpurun_out/r02b_knn_probe.jsonl; tail -3 gpurun_out/r02b_knn_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_ --launch-skip 4 -c 4 -f -o gpurun_out/r02b_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02b_ncu.log 2>&1; tail -3 gpurun_out/r02b_ncu.log
ncu -i gpurun_out/r02b_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02b_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02b_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02b_brick_knn_summary.txt 2>&1; cat gpurun_out/r02b_brick_knn_summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err; tail -c 300 gpurun_out/r02b_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 2500 gpurun_out/r02b_bench.json; tail -5 gpurun_out/r02b_bench.err
