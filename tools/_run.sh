( timeout 900 python -m pytest tests -x -q -m gpu --durations=5 ) > gpurun_out/gputest_r02_s2.log 2>&1
tail -8 gpurun_out/gputest_r02_s2.log
( timeout 600 python bench.py ) > gpurun_out/bench_r02_s2_n1.json 2> gpurun_out/bench_r02_s2_n1.err
tail -c 1500 gpurun_out/bench_r02_s2_n1.json; tail -3 gpurun_out/bench_r02_s2_n1.err
( timeout 300 python bench.py --impl reference ) > gpurun_out/bench_r02_s2_n1_ref.json 2> gpurun_out/bench_r02_s2_n1_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_s2.csv python bench.py --steps 2 --warmup 1 --no-rmat > gpurun_out/ncu_l2.log 2>&1
tail -2 gpurun_out/ncu_l2.log | cut -c1-300
