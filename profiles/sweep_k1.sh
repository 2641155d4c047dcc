for g in 64 32; do for c in 8 12 16 24 32 48; do
  echo "fetch=$g chains=$c: $(ZNG_B200_L2FETCH=$g ZNG_B200_CHAINS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done; done
