# Round 2, GPU call 1:  gpurun --timeout 1800 -- 'bash tools/run_r02a.sh'
# New defaults (stale rows, PDL, pipelined voxel grid), strict markers, config[1] parity, degenerate scene, the bench line.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
# 1. the test that failed in round 1, on its own, full output
timeout 600 python -m pytest tests/test_gpu_zz_ref_cuda_vgicp.py -m gpu -q -x > gpurun_out/r02a_vgicp.log 2>&1; tail -40 gpurun_out/r02a_vgicp.log
# 2. the whole GPU suite
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02a_pytest.log 2>&1; tail -15 gpurun_out/r02a_pytest.log
# 3. the bench pair as the driver runs it: reference arm first, then ours
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; tail -c 800 gpurun_out/r02a_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 3000 gpurun_out/r02a_bench.json; tail -5 gpurun_out/r02a_bench.err
# 4. the same with the launch-side defaults off (what PDL + pipelined voxel grid are worth)
LSD_PDL=0 LSD_PIPELINE_VG=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02a_bench_flags_off.json 2> /dev/null; tail -c 600 gpurun_out/r02a_bench_flags_off.json
# 5. launch list of the bench step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02a_bench_under_ncu.log 2>&1
python tools/launch_list.py gpurun_out/r02a_launches.csv > gpurun_out/r02a_launches_summary.txt 2>&1; tail -30 gpurun_out/r02a_launches_summary.txt
