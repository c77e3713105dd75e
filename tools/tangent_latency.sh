for S in 1 101; do
for cfg in "reg 1 0.2" "reg 2 0.2" "stream 2 0.2" "stream 4 0.2" "reg 1 0.1" "stream 2 0.1" "stream 4 0.1"; do set -- $cfg
  GAB1_TANGENT=$1 GAB1_TANGENT_NT=$2 timeout 300 python tools/bench_tangent.py --reps 2 --dr $3 --sets $S --tf 5.0 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('S=$S', d['family'], d['nt'], 'dr=$3', round(1e3*d['s_per_call'],1), 'ms per call;', 'primal', round(1e3*d['primal_s_per_call'],1))"
done; done
