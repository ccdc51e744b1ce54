# Round 2, GPU call 11 (1 GPU): brick k-NN v7 (ordered-batch gather) A/B against the single-buffer build, the lean bench loop, ncu of the brick kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_lio.py tests/test_gpu_map_knn.py -q -x > gpurun_out/r02k_pytest.log 2>&1; tail -5 gpurun_out/r02k_pytest.log
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02k_knn_probe.jsonl 2> gpurun_out/r02k_knn_probe.err; cut -c1-400 gpurun_out/r02k_knn_probe.jsonl; tail -3 gpurun_out/r02k_knn_probe.err
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 5 --lib lidar-slam-detection_b200/liblsdreg_ab_single.so > gpurun_out/r02k_knn_probe_single.jsonl 2> gpurun_out/r02k_knn_probe_single.err; cut -c1-400 gpurun_out/r02k_knn_probe_single.jsonl; tail -3 gpurun_out/r02k_knn_probe_single.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_ --launch-skip 4 -c 8 -f -o gpurun_out/r02k_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02k_ncu.log 2>&1; tail -3 gpurun_out/r02k_ncu.log
ncu -i gpurun_out/r02k_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02k_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02k_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02k_brick_knn_summary.txt 2>&1; cat gpurun_out/r02k_brick_knn_summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err; tail -c 300 gpurun_out/r02k_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; tail -c 1500 gpurun_out/r02k_bench.json; tail -5 gpurun_out/r02k_bench.err
