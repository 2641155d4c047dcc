# slab pipeline geometry (chunks per slab x slabs in flight) for the slow levels: host->host GB/s of zng_b200_deflate_host on 1 GiB
for L in ${LEVELS:-6 4 2}; do for g in ${GEOMS:-2048:4 4096:4 8192:2 8192:3 16384:2}; do set -- ${g%:*} ${g#*:}
  echo "level $L slab_chunks $1 pipe $2: $(ZNG_B200_SLAB_CHUNKS=$1 ZNG_B200_PIPE=$2 timeout 120 python profiles/trace_host_level.py $L 2>/dev/null | tail -1)"
done; done
