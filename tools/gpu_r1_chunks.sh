for CH in 32 16 8 4; do
  SDMB200_CHUNK=$CH timeout 300 ncu --metrics gpu__time_duration.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pair_cluster_kernel -s 10 -c 3 --csv --log-file gpurun_out/r1_ch$CH.csv python bench.py --replicas 1 --steps 5 --warmup 3 --no-cpu-baseline --e2e-depth 1 > /dev/null 2>&1
  echo "chunk $CH"; grep -E "gpu__time_duration|warps_active|issue_active" gpurun_out/r1_ch$CH.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | head -9
done
