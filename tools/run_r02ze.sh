# Round 2, GPU call 30 (1 GPU, the last of the budget): voxel grid with pcl::VoxelGrid's sequential fp32 sums in input order
# (bit-exact vs the oracle) — tests, then the bench pair for pose_parity against the reference arm.
set -x
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_voxelgrid.py tests/test_gpu_lio_config1.py tests/test_gpu_lio.py tests/test_gpu_zz_sequence.py -m gpu -q -x -s > gpurun_out/r02ze_pytest.log 2>&1; grep -E "config\[1\]|passed|failed|Error" gpurun_out/r02ze_pytest.log | tail -5
timeout 150 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02ze_bench_ref.json 2> gpurun_out/r02ze_bench_ref.err; tail -c 200 gpurun_out/r02ze_bench_ref.json
timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-knn-batch --streams 0 > gpurun_out/r02ze_bench.json 2> gpurun_out/r02ze_bench.err; tail -c 400 gpurun_out/r02ze_bench.json; tail -3 gpurun_out/r02ze_bench.err
