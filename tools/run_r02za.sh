# Round 2, GPU call 26 (1 GPU): batched k-NN with the entry format chosen per batch (16-byte entries for spatially ordered batches).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_map_knn.py tests/test_gpu_lru.py -m gpu -q -x > gpurun_out/r02za_pytest.log 2>&1; tail -3 gpurun_out/r02za_pytest.log
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02za_knn_probe.jsonl 2> gpurun_out/r02za_knn_probe.err; cut -c1-330 gpurun_out/r02za_knn_probe.jsonl; tail -2 gpurun_out/r02za_knn_probe.err
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 5 > gpurun_out/r02za_knn_probe_b.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r02za_knn_probe_b.jsonl
