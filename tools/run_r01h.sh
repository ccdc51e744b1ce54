set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r01h.log 2>&1; tail -5 gpurun_out/pytest_r01h.log
timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_h.log 2>&1; tail -1 gpurun_out/knn_probe_h.log
timeout 600 python bench_extra.py --which gicp --gicp-pairs 2 --no-cpu > gpurun_out/extra_gicp_h.log 2>&1; tail -2 gpurun_out/extra_gicp_h.log
timeout 600 python bench_extra.py --which gicp --gicp-pairs 2 --gicp-method FAST_VGICP > gpurun_out/extra_vgicp_h.log 2>&1; tail -2 gpurun_out/extra_vgicp_h.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/gicp_launches_h.csv python bench_extra.py --which gicp --gicp-pairs 1 --no-cpu > gpurun_out/extra_gicp_ncu.log 2>&1
