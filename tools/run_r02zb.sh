# Round 2, GPU call 27 (1 GPU): the state that ends the round (reference order the default, bench sorted_order leg):
# whole GPU suite, smoke, both bench arms as the driver runs them, launch list.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02zb_pytest.log 2>&1; tail -6 gpurun_out/r02zb_pytest.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke(); print("SMOKE_OK")' 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02zb_bench_ref.json 2> gpurun_out/r02zb_bench_ref.err; tail -c 300 gpurun_out/r02zb_bench_ref.json; tail -3 gpurun_out/r02zb_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02zb_bench.json 2> gpurun_out/r02zb_bench.err; tail -c 1800 gpurun_out/r02zb_bench.json; tail -5 gpurun_out/r02zb_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02zb_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 0 --no-reference-order-leg > gpurun_out/r02zb_bench_under_ncu.log 2>&1
python tools/launch_list.py gpurun_out/r02zb_launches.csv > gpurun_out/r02zb_launches_summary.txt 2>&1; tail -30 gpurun_out/r02zb_launches_summary.txt
