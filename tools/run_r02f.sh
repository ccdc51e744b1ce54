# Round 2, GPU call 6: whole GPU suite on the current state, brick k-NN v5 (prefetched items, queries in brick order), bench pair.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest.log 2>&1; tail -12 gpurun_out/r02f_pytest.log
timeout 600 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02f_knn_probe.jsonl 2> gpurun_out/r02f_knn_probe.err; cat gpurun_out/r02f_knn_probe.jsonl; tail -3 gpurun_out/r02f_knn_probe.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:brick_ --launch-skip 4 -c 4 -f -o gpurun_out/r02f_brick_knn python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 2 > gpurun_out/r02f_ncu.log 2>&1; tail -3 gpurun_out/r02f_ncu.log
ncu -i gpurun_out/r02f_brick_knn.ncu-rep --page raw --csv > gpurun_out/r02f_brick_knn_ncu_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r02f_brick_knn_ncu_raw.csv --items 2097152 --alg-bytes 680 > gpurun_out/r02f_brick_knn_summary.txt 2>&1; cat gpurun_out/r02f_brick_knn_summary.txt
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02f_bench_ref.json 2> gpurun_out/r02f_bench_ref.err; tail -c 300 gpurun_out/r02f_bench_ref.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; tail -c 1500 gpurun_out/r02f_bench.json; tail -5 gpurun_out/r02f_bench.err
