# Round 2, GPU call 24 (1 GPU): batched k-NN A/B — 32-byte vs 16-byte sorted query entries.
set -x
mkdir -p gpurun_out
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 > gpurun_out/r02y_knn_probe.jsonl 2> gpurun_out/r02y_knn_probe.err; cut -c1-330 gpurun_out/r02y_knn_probe.jsonl; tail -2 gpurun_out/r02y_knn_probe.err
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3,2 --reps 5 --lib lidar-slam-detection_b200/liblsdreg_ab_q16.so > gpurun_out/r02y_knn_probe_q16.jsonl 2> gpurun_out/r02y_knn_probe_q16.err; cut -c1-330 gpurun_out/r02y_knn_probe_q16.jsonl; tail -2 gpurun_out/r02y_knn_probe_q16.err
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 5 > gpurun_out/r02y_knn_probe_b.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r02y_knn_probe_b.jsonl
timeout 300 python tools/knn_batch_probe.py 2097152 --shapes 3 --reps 5 --lib lidar-slam-detection_b200/liblsdreg_ab_q16.so > gpurun_out/r02y_knn_probe_q16_b.jsonl 2>/dev/null; cut -c1-200 gpurun_out/r02y_knn_probe_q16_b.jsonl
