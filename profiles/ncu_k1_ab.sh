O=gpurun_out
ZNG_B200_FLAGS=8 ZNG_B200_CHAINS=12 ncu --set full --clock-control none --import-source on -k regex:'quick_parse' --launch-skip 3 -c 1 -f -o $O/r1_k1_narrow_c12 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ab1.log 2>&1
ZNG_B200_FLAGS=$((4*256+8)) ZNG_B200_CHAINS=12 ncu --set full --clock-control none --import-source on -k regex:'quick_parse' --launch-skip 3 -c 1 -f -o $O/r1_k1_wide4_c12 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/ab2.log 2>&1
ls -la $O/*c12*
