set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01i.log 2>&1; tail -15 gpurun_out/pytest_r01i.log
timeout 600 python bench_extra.py --which gicp --gicp-pairs 4 > gpurun_out/extra_gicp_i.log 2>&1; tail -1 gpurun_out/extra_gicp_i.log
timeout 600 python bench_extra.py --which gicp --gicp-pairs 4 --gicp-method FAST_VGICP > gpurun_out/extra_vgicp_i.log 2>&1; tail -1 gpurun_out/extra_vgicp_i.log
