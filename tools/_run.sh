export BHB200_PATTERN=off
( timeout 600 python -m pytest tests/test_spgemm_gpu.py -x -q -m gpu -k "bucket or wide or direct or rmat" ) > gpurun_out/b3w_pytest.log 2>&1
tail -3 gpurun_out/b3w_pytest.log
( timeout 300 python tools/bin_report.py rect ) > gpurun_out/b3w_bins_rect2.txt 2>&1
( timeout 200 python tools/bucket_dev.py bins 21 ) > gpurun_out/b3w_bins_rmat21_2.txt 2>&1
for f in gpurun_out/b3w_bins_rect2.txt gpurun_out/b3w_bins_rmat21_2.txt; do echo $f; grep -A9 "^total" $f | cut -c1-700; done
