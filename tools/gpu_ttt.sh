python -m pytest tests/test_time_to_tolerance.py -q -m gpu 2>&1 | tail -2
python tools/time_to_tolerance.py > gpurun_out/r49_ttt.jsonl 2> gpurun_out/r49_ttt.err; echo "ttt rc=$?"; cut -c1-150 gpurun_out/r49_ttt.jsonl; tail -2 gpurun_out/r49_ttt.err
