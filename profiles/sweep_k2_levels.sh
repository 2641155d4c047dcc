# levels 4-6 (deep hash chains): chains per SM.  prev[] + head[] + chunk = 256 KiB per chain; 16 chains/SM ~ L2 capacity
for lvl in ${LEVELS:-6 5 4}; do for c in ${CHAINS:-8 12 16 20 24 32 48}; do
  echo "level $lvl chains $c: $(ZNG_B200_CHAINS_L2=$c timeout 120 python bench.py --workload deflate$lvl --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-parity 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["roofline"]["kernel_ms"],1))')"
done; done
