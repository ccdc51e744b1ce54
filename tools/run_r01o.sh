set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01o.log 2>&1; tail -6 gpurun_out/pytest_r01o.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01o.log 2>&1; tail -3 gpurun_out/smoke_r01o.log
