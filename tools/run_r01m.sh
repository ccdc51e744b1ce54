set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reg.py tests/test_gpu_map_knn.py -x -q > gpurun_out/pytest_r01m.log 2>&1; tail -5 gpurun_out/pytest_r01m.log
for G in 1 2 4; do LSD_KNN_GROUP=$G timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_m_g$G.log 2>&1; echo G=$G; tail -1 gpurun_out/knn_probe_m_g$G.log; done
LSD_KNN_GROUP=2 LSD_L2_FETCH_GRANULARITY=64 timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_m_g2_f64.log 2>&1; tail -1 gpurun_out/knn_probe_m_g2_f64.log
LSD_KNN_GROUP=2 LSD_L2_FETCH_GRANULARITY=32 timeout 300 python tools/knn_probe.py 2097152 > gpurun_out/knn_probe_m_g2_f32.log 2>&1; tail -1 gpurun_out/knn_probe_m_g2_f32.log
