#!/bin/bash
# Round-1 measurement suite (run on the GPU box through gpurun): plain bench lines first, then ncu.
# A number printed by a run under ncu is never a bench value.
set -u
O=gpurun_out
mkdir -p $O
for w in deflate1 deflate2 checksum inflate; do
  python bench.py --workload $w --steps 10 --warmup 3 > $O/r1_bench_$w.json 2> $O/r1_bench_$w.err || echo "bench $w failed rc=$?"
  tail -c 300 $O/r1_bench_$w.err
done
python bench.py --impl reference --steps 3 --warmup 1 > $O/r1_bench_reference.json 2> $O/r1_bench_reference.err || echo "reference arm failed"
python profiles/measure_stream_inflate.py 1024 > $O/r1_host_api_roundtrip.txt 2>&1 || echo "round trip failed"
# launch lists (per-launch durations; cold-cache and serialised: use the SHARES)
for w in deflate1 deflate2 checksum inflate; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/r1_launches_$w.csv \
      python bench.py --workload $w --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/r1_ncu_list_$w.log 2>&1 || echo "ncu list $w failed"
done
# one full capture per hot kernel (the timed step's launches, after 3 warm-up steps)
ncu --set full --clock-control none --import-source on -k regex:'quick_parse|static_emit|checksum_tiles' --launch-skip 9 -c 3 -f -o $O/r1_full_deflate1 \
    python bench.py --workload deflate1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r1_ncu_full_deflate1.log 2>&1 || echo "ncu full deflate1 failed"
ncu --set full --clock-control none --import-source on -k regex:'fast_parse|block_emit' --launch-skip 6 -c 2 -f -o $O/r1_full_deflate2 \
    python bench.py --workload deflate2 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r1_ncu_full_deflate2.log 2>&1 || echo "ncu full deflate2 failed"
ncu --set full --clock-control none --import-source on -k regex:'checksum_tiles|crc32_fold' --launch-skip 6 -c 3 -f -o $O/r1_full_checksum \
    python bench.py --workload checksum --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r1_ncu_full_checksum.log 2>&1 || echo "ncu full checksum failed"
ncu --set full --clock-control none --import-source on -k regex:'inflate_members' --launch-skip 3 -c 1 -f -o $O/r1_full_inflate \
    python bench.py --workload inflate --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $O/r1_ncu_full_inflate.log 2>&1 || echo "ncu full inflate failed"
ls -la $O | tail -12
