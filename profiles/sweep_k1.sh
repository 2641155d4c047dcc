# K1 chains per SM (each chain = one warp with a 128 KiB hash-head table + its 64 KiB chunk)
for c in 4 8 12 16 24 32 48; do
  echo "chains=$c: $(ZNG_B200_CHAINS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done
