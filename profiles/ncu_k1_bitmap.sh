O=gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'quick_parse' --launch-skip 3 -c 1 -f -o $O/r1_k1_bitmap_c24 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ab3.log 2>&1
ls -la $O/r1_k1_bitmap_c24.ncu-rep
