export BHB200_PATTERN=off
( timeout 600 python -m pytest tests/test_spgemm_gpu.py -x -q -m gpu -k "bucket or wide or heavy or rmat or spill" ) > gpurun_out/b3_pytest.log 2>&1
tail -3 gpurun_out/b3_pytest.log
( BHB200_DEBUG_FORCE_HEAVY=1 timeout 300 python tools/bucket_dev.py check 17 19 ) > gpurun_out/b3_check.log 2>&1
grep -c OK gpurun_out/b3_check.log; grep FAIL gpurun_out/b3_check.log
( BHB200_DEBUG_FORCE_HEAVY=1 timeout 200 python tools/bucket_dev.py bins 21 ) > gpurun_out/b3_bins_v3e_heavy2.txt 2>&1
( timeout 300 python tools/block_report.py 8 0 ) > gpurun_out/b3_block_s24_r0_e.txt 2>&1
( timeout 300 python tools/block_report.py 8 3 ) > gpurun_out/b3_block_s24_r3_e.txt 2>&1
for f in gpurun_out/b3_bins_v3e*.txt gpurun_out/b3_block_s24_r*_e.txt; do echo $f; grep -B1 -A3 "^total" $f | cut -c1-900; done
