# Round 2, GPU call 13 (1 GPU): reference neighbour order + Eigen-order plane solve on hardware — parity tests, what the mode costs, the bench with it.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lio.py tests/test_gpu_zz_sequence.py tests/test_gpu_lio_config1.py tests/test_gpu_shard.py tests/test_gpu_zz_fastlio_seam.py -m gpu -q -x -s > gpurun_out/r02n_pytest.log 2>&1; grep -E "config\[1\]|passed|failed|Error" gpurun_out/r02n_pytest.log | tail -8
timeout 600 python tools/lio_probe.py "" "LSD_REF_ORDER=1" > gpurun_out/r02n_lio_probe.jsonl 2> gpurun_out/r02n_lio_probe.err; cut -c1-700 gpurun_out/r02n_lio_probe.jsonl; tail -3 gpurun_out/r02n_lio_probe.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err; tail -c 300 gpurun_out/r02n_bench_ref.json
LSD_REF_ORDER=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-knn-batch --streams 0 > gpurun_out/r02n_bench_reforder.json 2> gpurun_out/r02n_bench_reforder.err; tail -c 1500 gpurun_out/r02n_bench_reforder.json; tail -5 gpurun_out/r02n_bench_reforder.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-knn-batch --streams 0 > gpurun_out/r02n_bench_default.json 2> gpurun_out/r02n_bench_default.err; tail -c 1200 gpurun_out/r02n_bench_default.json; tail -5 gpurun_out/r02n_bench_default.err
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest_all.log 2>&1; tail -5 gpurun_out/r02n_pytest_all.log
