# K1a v6 (CTA per chain, producers one window ahead of the walker) against the warp-per-chain parser:
# producer warps per CTA x CTAs (chains) per SM.  Output: GB/s device-resident, ms of the step's hot kernels, parity.
run() { python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["roofline"]["kernel_ms"],2), d["parity"][:40])'; }
echo "warp-per-chain (24 chains/SM): $(ZNG_B200_K1=warp run)"
for w in ${WARPS:-8 6 4 12}; do
  for c in ${CHAINS:-2 3 4 5 6}; do
    echo "cta producers=$w chains=$c: $(ZNG_B200_K1=cta ZNG_B200_K1_WARPS=$w ZNG_B200_K1_CHAINS=$c run)"
  done
done
