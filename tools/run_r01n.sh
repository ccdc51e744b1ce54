set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01n.log 2>&1; tail -8 gpurun_out/pytest_r01n.log
