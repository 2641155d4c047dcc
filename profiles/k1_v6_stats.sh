# K1a v7 debug counters per configuration "producer-warps blocks-per-producer-warp chains-per-SM": GB/s device-resident + the parser's own counters
for cfg in ${CFGS:-"3 1 10" "2 2 11" "2 2 13" "1 4 13" "1 2 13"}; do set -- $cfg; echo "== producers=$1 blocks/producer=$2 chains=$3"; ZNG_B200_K1_STATS=${STATS:-1} ZNG_B200_K1=cta ZNG_B200_K1_WARPS=$1 ZNG_B200_K1_BPW=$2 ZNG_B200_K1_CHAINS=$3 timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${PARITY:---no-parity} 2>&1 | python -c '
import sys, json
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line); print("GB/s", round(d["value"], 2), "kernel_ms", round(d["roofline"]["kernel_ms"], 2), d["parity"][:60])
    elif "K1 v6" in line: print(line.strip())
'; done
