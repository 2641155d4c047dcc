# K1a v5 (CTA per chain) against the warp-per-chain parser: warps per CTA x CTAs (chains) per SM
run() { python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["roofline"]["kernel_ms"],2), d["parity"][:9])'; }
echo "warp-per-chain (24 chains/SM): $(run)"
for w in 8 4 2; do
  for c in 1 2 3 4 6 8 12 16; do
    if [ $((w*c)) -le 32 ]; then
      echo "cta warps=$w chains=$c: $(ZNG_B200_K1=cta ZNG_B200_K1_WARPS=$w ZNG_B200_K1_CHAINS=$c run)"
    fi
  done
done
