# K1a v6/v7 debug counters per configuration (producer warps, chains per SM): GB/s device-resident + the parser's own counters
for cfg in ${CFGS:-"3 6" "3 8" "3 10" "2 10" "2 13" "4 8"}; do set -- $cfg; echo "== producers=$1 chains=$2"; ZNG_B200_K1_STATS=1 ZNG_B200_K1=cta ZNG_B200_K1_WARPS=$1 ZNG_B200_K1_CHAINS=$2 timeout 120 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline ${PARITY:---no-parity} 2>&1 | python -c '
import sys, json
for line in sys.stdin:
    if line.startswith("{"):
        d = json.loads(line); print("GB/s", round(d["value"], 2), "kernel_ms", round(d["roofline"]["kernel_ms"], 2), d["parity"][:60])
    elif "K1 v6" in line: print(line.strip())
'; done
