# Round 2, GPU call 4: lane-group brick k-NN, cluster shapes of the reuse evaluation, the GICP k-NN fix.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_reg.py tests/test_gpu_zz_ref_cuda_vgicp.py tests/test_gpu_lio.py tests/test_gpu_zz_pdl.py -m gpu -q -s > gpurun_out/r02d_pytest.log 2>&1; tail -14 gpurun_out/r02d_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02d_knn_probe.jsonl 2> gpurun_out/r02d_knn_probe.err; cat gpurun_out/r02d_knn_probe.jsonl; tail -3 gpurun_out/r02d_knn_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_knn --launch-skip 1 -c 1 -f -o gpurun_out/r02d_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02d_ncu.log 2>&1; tail -3 gpurun_out/r02d_ncu.log
ncu -i gpurun_out/r02d_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02d_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02d_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02d_brick_knn_summary.txt 2>&1; cat gpurun_out/r02d_brick_knn_summary.txt
timeout 900 python tools/lio_probe.py "LSD_REUSE_CLUSTER=0" "LSD_REUSE_CLUSTER=8x256" "LSD_REUSE_CLUSTER=8x512" "LSD_REUSE_CLUSTER=4x512" "LSD_REUSE_CLUSTER=8x128" "LSD_REUSE_CLUSTER=16x256" "LSD_REUSE_CLUSTER=2x1024" "LSD_REUSE_CLUSTER=0,LSD_PDL=0" > gpurun_out/r02d_lio_probe.jsonl 2> gpurun_out/r02d_lio_probe.err; cat gpurun_out/r02d_lio_probe.jsonl; tail -3 gpurun_out/r02d_lio_probe.err
