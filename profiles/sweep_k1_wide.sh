# K1 wide-window parser: blocks per step (flags bits 8..10) x chains per SM; flags low bits 8 = touch prefetch
timeout 900 python -m pytest tests -m gpu -x -q -k "deflate_quick or host_path" 2>&1 | tail -2
for w in 2 4; do
  ZNG_B200_FLAGS=$((w*256+8)) timeout 900 python -m pytest tests -m gpu -x -q -k "deflate_quick or host_path" 2>&1 | tail -2
done
for w in 0 2 3 4; do for c in 6 8 12 16 24; do
  echo "wide=$w chains=$c: $(ZNG_B200_FLAGS=$((w*256+8)) ZNG_B200_CHAINS=$c python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), d["parity"][:9])')"
done; done
