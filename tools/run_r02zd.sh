# Round 2, GPU call 29 (1 GPU): the library as committed at the end of the round — LIO / k-NN tests and smoke().
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lio.py tests/test_gpu_map_knn.py tests/test_gpu_brick_knn.py tests/test_gpu_zz_sequence.py -m gpu -q -x > gpurun_out/r02zd_pytest.log 2>&1; tail -3 gpurun_out/r02zd_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("SMOKE_OK")' 2>&1 | tail -2
