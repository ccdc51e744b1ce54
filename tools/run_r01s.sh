set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ref_cuda.py -q > gpurun_out/pytest_r01s.log 2>&1; tail -30 gpurun_out/pytest_r01s.log
timeout 600 python tools/ref_cuda_probe.py 3 > gpurun_out/ref_cuda_probe.log 2>&1; tail -3 gpurun_out/ref_cuda_probe.log
