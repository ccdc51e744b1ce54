set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reg.py -x -q > gpurun_out/pytest_r01l.log 2>&1; tail -5 gpurun_out/pytest_r01l.log
timeout 300 python tools/reg_build_probe.py > gpurun_out/reg_probe_l.log 2>&1; cat gpurun_out/reg_probe_l.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/vgicp_launches_l.csv python tools/reg_build_probe.py FAST_VGICP > gpurun_out/reg_probe_ncu.log 2>&1
timeout 600 python bench_extra.py --which gicp --gicp-pairs 8 --no-cpu > gpurun_out/extra_gicp_l.log 2>&1; tail -1 gpurun_out/extra_gicp_l.log
