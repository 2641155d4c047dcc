for cfg in "1024 6" "2048 4" "4096 3" "4096 4" "512 12" "256 16"; do set -- $cfg
  echo "slab_chunks=$1 pipe=$2: $(ZNG_B200_SLAB_CHUNKS=$1 ZNG_B200_PIPE=$2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["value"],2), round(d["e2e"]["value"],2))')"
done
