# Round 2, GPU call 5: brick k-NN v4, NDT shard test (ranks emulated on one device), defaults after the r02d measurements.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_brick_knn.py tests/test_gpu_shard.py tests/test_gpu_reg.py -m gpu -q -s > gpurun_out/r02e_pytest.log 2>&1; tail -14 gpurun_out/r02e_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02e_knn_probe.jsonl 2> gpurun_out/r02e_knn_probe.err; cat gpurun_out/r02e_knn_probe.jsonl; tail -3 gpurun_out/r02e_knn_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_knn --launch-skip 1 -c 1 -f -o gpurun_out/r02e_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02e_ncu.log 2>&1; tail -3 gpurun_out/r02e_ncu.log
ncu -i gpurun_out/r02e_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02e_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02e_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02e_brick_knn_summary.txt 2>&1; cat gpurun_out/r02e_brick_knn_summary.txt
timeout 900 python tools/lio_probe.py "" "LSD_PDL=1" "LSD_PIPELINE_VG=0" > gpurun_out/r02e_lio_probe.jsonl 2> gpurun_out/r02e_lio_probe.err; cat gpurun_out/r02e_lio_probe.jsonl; tail -3 gpurun_out/r02e_lio_probe.err
timeout 600 python bench_extra.py --which gicp --gicp-pairs 16 > gpurun_out/r02e_extra_gicp.jsonl 2> gpurun_out/r02e_extra_gicp.err; cat gpurun_out/r02e_extra_gicp.jsonl; tail -3 gpurun_out/r02e_extra_gicp.err
