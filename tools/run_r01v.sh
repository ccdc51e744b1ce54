set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_r01v.log 2>&1; tail -4 gpurun_out/pytest_r01v.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r01v.log 2>&1; tail -2 gpurun_out/smoke_r01v.log
timeout 400 python bench_extra.py --which ndt > gpurun_out/extra_ndt_v.log 2>&1; tail -1 gpurun_out/extra_ndt_v.log
timeout 300 python bench_extra.py --which gicp,vfe --gicp-pairs 8 > gpurun_out/extra_gicp_vfe_v.log 2>&1; tail -2 gpurun_out/extra_gicp_vfe_v.log
timeout 300 python bench_extra.py --which gicp --gicp-method FAST_VGICP --gicp-pairs 8 > gpurun_out/extra_vgicp_v.log 2>&1; tail -1 gpurun_out/extra_vgicp_v.log
