# Round 2, GPU call 28 (1 GPU): cell-major ordered gather (the list is the reference's candidate sequence as gathered) — tests, cost, bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lio.py tests/test_gpu_zz_sequence.py tests/test_gpu_lio_config1.py tests/test_gpu_map_knn.py tests/test_gpu_lru.py tests/test_gpu_shard.py tests/test_gpu_zz_fastlio_seam.py -m gpu -q -x > gpurun_out/r02zc_pytest.log 2>&1; tail -3 gpurun_out/r02zc_pytest.log
timeout 600 python tools/lio_probe.py "" "LSD_REF_ORDER=0" > gpurun_out/r02zc_lio_probe.jsonl 2> gpurun_out/r02zc_lio_probe.err; cut -c1-500 gpurun_out/r02zc_lio_probe.jsonl; tail -3 gpurun_out/r02zc_lio_probe.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02zc_bench.json 2> gpurun_out/r02zc_bench.err; tail -c 500 gpurun_out/r02zc_bench.json; tail -3 gpurun_out/r02zc_bench.err
