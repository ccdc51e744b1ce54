set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_r01p.log 2>&1; tail -6 gpurun_out/pytest_r01p.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r01p.json 2> gpurun_out/bench_r01p.err; tail -c 1500 gpurun_out/bench_r01p.json
