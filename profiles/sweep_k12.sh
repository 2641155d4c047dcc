timeout 900 python -m pytest tests -m gpu -x -q -k "deflate or host" 2>&1 | tail -2
for c in 16 20 24; do
  echo "L1 chains=$c: $(ZNG_B200_CHAINS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done
for c in 8 12 16 20 24; do
  echo "L2 chains=$c: $(ZNG_B200_CHAINS_L2=$c python bench.py --workload deflate2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done
