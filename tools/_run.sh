( timeout 300 python -m pytest tests/test_dist_gpu.py -x -q -m gpu ) > gpurun_out/gputest_r02_s2_dist.log 2>&1
tail -3 gpurun_out/gputest_r02_s2_dist.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 ) > gpurun_out/bench_r02_s2_n2.json 2> gpurun_out/bench_r02_s2_n2.err
tail -c 2500 gpurun_out/bench_r02_s2_n2.json; tail -3 gpurun_out/bench_r02_s2_n2.err
