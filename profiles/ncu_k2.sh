O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'fast_parse|block_emit' --launch-skip 6 -c 2 -f -o $O/r1_full_deflate2_b python bench.py --workload deflate2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/k2b.log 2>&1
ls -la $O/r1_full_deflate2_b.ncu-rep
