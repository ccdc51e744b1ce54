# Round 2, GPU call 20 (1 GPU): reference order with the everything inside the search kernel (no export, no replay kernel) — parity tests and what the mode costs now.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lio.py tests/test_gpu_zz_sequence.py tests/test_gpu_lio_config1.py -m gpu -q -x -s > gpurun_out/r02u_pytest.log 2>&1; grep -E "config\[1\]|passed|failed|Error" gpurun_out/r02u_pytest.log | tail -8
timeout 600 python tools/lio_probe.py "" "LSD_REF_ORDER=1" > gpurun_out/r02u_lio_probe.jsonl 2> gpurun_out/r02u_lio_probe.err; cut -c1-700 gpurun_out/r02u_lio_probe.jsonl; tail -3 gpurun_out/r02u_lio_probe.err
LSD_REF_ORDER=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"lio_" --launch-skip 40 -c 120 --csv --log-file gpurun_out/r02u_launches.csv python tools/lio_probe.py "" > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r02u_launches.csv 2>&1 | tail -8
