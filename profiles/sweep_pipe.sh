# host pipeline of zng_b200_deflate_host: slab size (chunks) x slabs in flight; prints device GB/s and e2e GB/s
for cfg in "2048 4" "2048 6" "1024 6" "4096 4"; do set -- $cfg
  echo "slab_chunks=$1 pipe=$2: $(ZNG_B200_SLAB_CHUNKS=$1 ZNG_B200_PIPE=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["e2e"]["value"],2))')"
done
