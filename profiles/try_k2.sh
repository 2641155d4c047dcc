timeout 900 python -m pytest tests -m gpu -x -q -k "deflate_fast" 2>&1 | tail -3
echo "L2: $(python bench.py --workload deflate2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/k2_launches.csv python bench.py --workload deflate2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
grep -E "fast_parse|block_emit" gpurun_out/k2_launches.csv | tail -2 | awk -F'","' '{print substr($5,1,40), $NF}'
