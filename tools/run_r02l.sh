# Round 2, GPU call 12 (1 GPU): the state that ends the round — whole GPU suite, both bench arms as the driver runs them, configs 3-5,
# the reference's own CUDA kernels beside ours, the launch list of the bench command and an ncu --set full pass over the LIO kernels.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest.log 2>&1; tail -6 gpurun_out/r02l_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("SMOKE_OK")' 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02l_bench_ref.json 2> gpurun_out/r02l_bench_ref.err; tail -c 400 gpurun_out/r02l_bench_ref.json; tail -3 gpurun_out/r02l_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; tail -c 1500 gpurun_out/r02l_bench.json; tail -5 gpurun_out/r02l_bench.err
timeout 900 python bench_extra.py --which ndt,gicp,vfe --gicp-pairs 16 > gpurun_out/r02l_extra.jsonl 2> gpurun_out/r02l_extra.err; cut -c1-700 gpurun_out/r02l_extra.jsonl; tail -3 gpurun_out/r02l_extra.err
timeout 600 python bench_extra.py --which gicp --gicp-pairs 16 --gicp-method FAST_VGICP > gpurun_out/r02l_extra_vgicp.jsonl 2> /dev/null; cut -c1-600 gpurun_out/r02l_extra_vgicp.jsonl
timeout 600 python tools/ref_cuda_probe.py > gpurun_out/r02l_ref_cuda_probe.json 2> gpurun_out/r02l_ref_cuda_probe.err; cut -c1-900 gpurun_out/r02l_ref_cuda_probe.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02l_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 0 > gpurun_out/r02l_bench_under_ncu.log 2>&1
python tools/launch_list.py gpurun_out/r02l_launches.csv > gpurun_out/r02l_launches_summary.txt 2>&1; tail -30 gpurun_out/r02l_launches_summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lio_|vg_" --launch-skip 60 -c 40 -f -o gpurun_out/r02l_lio python tools/lio_probe.py "" > gpurun_out/r02l_ncu.log 2>&1; tail -3 gpurun_out/r02l_ncu.log
ncu -i gpurun_out/r02l_lio.ncu-rep --page raw --csv > gpurun_out/r02l_lio_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02l_lio_ncu_raw.csv > gpurun_out/r02l_lio_summary.txt 2>&1; grep -E "^kernel|duration|dram read  |dram write" gpurun_out/r02l_lio_summary.txt | head -80
rm -f gpurun_out/r02l_lio.ncu-rep
