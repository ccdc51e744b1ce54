set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ref_cuda.py tests/test_gpu_vfe.py -q > gpurun_out/pytest_r01u.log 2>&1; tail -12 gpurun_out/pytest_r01u.log
timeout 600 python tools/ref_cuda_probe.py 3 > gpurun_out/ref_cuda_probe_u.log 2>&1; tail -2 gpurun_out/ref_cuda_probe_u.log
