# usage: bash tools/ab_env.sh VAR a b   -- runs the bench alternately with VAR=a and VAR=b on one box (in-box A/B)
for i in 1 2; do for v in $2 $3; do
  env $1=$v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-parity 2>/dev/null | tail -1 > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); print('$1=$v', round(d['ms_per_step'],4), round(d['in_flight']['ms_per_step'],4), round(d['e2e']['value']))"
done; done
