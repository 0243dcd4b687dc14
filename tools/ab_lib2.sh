# like tools/ab_lib.sh, with the kernel families to print as $2 (comma-separated): tools/ab_lib2.sh other.so hrn_knn,hrn_fps
cp pcd_reg_hregnet_b200/libhregnet_b200.so /tmp/base.so
for i in 1 2; do for v in base other; do
  if [ $v = other ]; then cp $1 pcd_reg_hregnet_b200/libhregnet_b200.so; else cp /tmp/base.so pcd_reg_hregnet_b200/libhregnet_b200.so; fi
  timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-gpu-reference --no-parity 2>/dev/null | tail -1 > /tmp/o.json
  python -c "import json; d=json.load(open('/tmp/o.json')); k=d['kernel_breakdown_ms_per_step']; print('$v', round(d['ms_per_step'],4), round(d['in_flight']['ms_per_step'],4), {n: k[n] for n in '$2'.split(',')})"
done; done
cp /tmp/base.so pcd_reg_hregnet_b200/libhregnet_b200.so
