for c in 16 24 32; do
  echo "L2 chains=$c: $(ZNG_B200_CHAINS_L2=$c python bench.py --workload deflate2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done
echo "L1: $(python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
