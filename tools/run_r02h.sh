# Round 2, GPU call 8: slam_wrapper.process on hardware, NDT config 3 with the iteration log, ncu of the LIO kernels.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_slam_wrapper.py tests/test_gpu_reg.py tests/test_gpu_lru.py tests/test_gpu_brick_knn.py tests/test_gpu_shard.py -m gpu -q > gpurun_out/r02h_pytest.log 2>&1; tail -12 gpurun_out/r02h_pytest.log
timeout 900 python bench_extra.py --which ndt,vfe > gpurun_out/r02h_extra.jsonl 2> gpurun_out/r02h_extra.err; cat gpurun_out/r02h_extra.jsonl; tail -3 gpurun_out/r02h_extra.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lio_|vg_" --launch-skip 60 -c 40 -f -o gpurun_out/r02h_lio python tools/lio_probe.py "" > gpurun_out/r02h_ncu.log 2>&1; tail -3 gpurun_out/r02h_ncu.log
ncu -i gpurun_out/r02h_lio.ncu-rep --page raw --csv > gpurun_out/r02h_lio_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02h_lio_ncu_raw.csv > gpurun_out/r02h_lio_summary.txt 2>&1; grep -E "^kernel|duration|issue slots|achieved occ|top stalls|dram read  " gpurun_out/r02h_lio_summary.txt | head -120
