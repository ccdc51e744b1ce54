# Round 2, first GPU call:  gpurun --timeout 1500 -- 'bash tools/run_r02a.sh'
# Everything written after round 1's GPU budget was spent, on its own, before anything else depends on it.
set -x
mkdir -p gpurun_out
# 1. the six never-run GPU tests, each reported separately (they are non-strict xfails: look for XPASS / the tail on XFAIL)
for t in flat_knn scancontext sequence fastlio_seam pdl ref_cuda_vgicp; do
  timeout 900 python -m pytest tests/test_gpu_zz_$t.py -m gpu -q -rxX --runxfail > gpurun_out/r02a_$t.log 2>&1; tail -15 gpurun_out/r02a_$t.log
done
# 2. the validated suite, to see that nothing moved
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02a_pytest.log 2>&1; tail -4 gpurun_out/r02a_pytest.log
# 3. A/B of the three k-NN shapes (bit identity, batch timing random / sorted, per-scan search inside the LIO stream)
timeout 600 python tools/knn_shapes_probe.py > gpurun_out/r02a_knn_shapes.jsonl 2> gpurun_out/r02a_knn_shapes.err; cat gpurun_out/r02a_knn_shapes.jsonl; tail -3 gpurun_out/r02a_knn_shapes.err
LSD_L2_FETCH_GRANULARITY=64 timeout 600 python tools/knn_shapes_probe.py --no-lio --shapes 2,3 > gpurun_out/r02a_knn_shapes_l2fetch64.jsonl 2>&1; tail -3 gpurun_out/r02a_knn_shapes_l2fetch64.jsonl
# 4. ncu: the flat kernel next to the thread kernel, 1 M queries (one launch each is enough: -c bounds the replay cost)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_query_flat_kernel -c 2 -o gpurun_out/r02a_knn_flat \
  python tools/knn_shapes_probe.py 1048576 --no-lio > gpurun_out/r02a_ncu_flat.log 2>&1; tail -2 gpurun_out/r02a_ncu_flat.log
timeout 600 ncu --set full --clock-control none -k regex:knn_query_thread_kernel -c 2 -o gpurun_out/r02a_knn_thread \
  python tools/knn_shapes_probe.py 1048576 --no-lio > gpurun_out/r02a_ncu_thread.log 2>&1; tail -2 gpurun_out/r02a_ncu_thread.log
for k in flat thread; do
  ncu -i gpurun_out/r02a_knn_$k.ncu-rep --page raw --csv > gpurun_out/r02a_knn_${k}_ncu_raw.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r02a_knn_${k}_ncu_raw.csv --items 1048576 --alg-bytes 680 > gpurun_out/r02a_knn_${k}_summary.txt 2>&1; cat gpurun_out/r02a_knn_${k}_summary.txt
done
# 5. (LIO rows of step 3 carry the PDL A/B: shape_0 / shape_4 with and without lsd_lio_set_pdl, device and wall ms per scan;
#    LSD_PDL=1 python bench.py --no-experimental is the whole bench line with it on)
# 6. the bench line of the unchanged default path (reference arm first, as the driver does)
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02a_bench_ref.json 2> gpurun_out/r02a_bench_ref.err; tail -c 600 gpurun_out/r02a_bench_ref.json
timeout 900 python bench.py > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err; tail -c 1500 gpurun_out/r02a_bench.json
LSD_PDL=1 timeout 600 python bench.py --no-experimental --no-knn-batch > gpurun_out/r02a_bench_pdl.json 2> gpurun_out/r02a_bench_pdl.err; tail -c 1500 gpurun_out/r02a_bench_pdl.json
LSD_PDL=1 LSD_PIPELINE_VG=1 timeout 600 python bench.py --no-experimental --no-knn-batch > gpurun_out/r02a_bench_pdl_pipe.json 2> gpurun_out/r02a_bench_pdl_pipe.err; tail -c 1500 gpurun_out/r02a_bench_pdl_pipe.json
