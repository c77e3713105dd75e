# round-2 session O (1 GPU): does the latency lane overlap the throughput kernel?  carveout A/B
set -x
for cv in "" 50 100; do
  GAB1_CARVEOUT=$cv timeout 300 python tools/duo_probe.py shards 2>&1 | grep probe | cut -c1-330
done
