# Round 2, GPU call 9: whole GPU suite + all benches on the state that goes to the 8-GPU run.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02i_pytest.log 2>&1; tail -8 gpurun_out/r02i_pytest.log
timeout 900 python bench_extra.py --which ndt,gicp,vfe --gicp-pairs 16 > gpurun_out/r02i_extra.jsonl 2> gpurun_out/r02i_extra.err; cat gpurun_out/r02i_extra.jsonl | cut -c1-1500; tail -3 gpurun_out/r02i_extra.err
timeout 600 python bench_extra.py --which gicp --gicp-pairs 16 --gicp-method FAST_VGICP > gpurun_out/r02i_extra_vgicp.jsonl 2> /dev/null; cat gpurun_out/r02i_extra_vgicp.jsonl | cut -c1-1200
timeout 600 python tools/ref_cuda_probe.py > gpurun_out/r02i_ref_cuda_probe.json 2> gpurun_out/r02i_ref_cuda_probe.err; cat gpurun_out/r02i_ref_cuda_probe.json | cut -c1-1500; tail -2 gpurun_out/r02i_ref_cuda_probe.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; tail -c 300 gpurun_out/r02i_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 1200 gpurun_out/r02i_bench.json; tail -5 gpurun_out/r02i_bench.err
