# Round 2, GPU call 23 (1 GPU): sanity of the last source change (serial replay: one code path) — LIO tests, one short bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lio.py tests/test_gpu_zz_sequence.py tests/test_gpu_lio_config1.py tests/test_gpu_shard.py -m gpu -q -x > gpurun_out/r02x_pytest.log 2>&1; tail -3 gpurun_out/r02x_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; tail -c 600 gpurun_out/r02x_bench.json; tail -3 gpurun_out/r02x_bench.err
