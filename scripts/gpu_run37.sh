set -x
mkdir -p gpurun_out
free -g | head -2
python bench_closest_hit.py --sizes 100000000 --shapes tri_soup --no-oracle 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['config']['workload'], round(d['value'],1), 'Mrays/s frac', round(d['roofline']['frac'],3), 'create', round(d['config']['scene_create_s'],1), 'host', round(d['config']['host_scene_s'],1))
" | tee gpurun_out/r2_ch100m.log
python bench_closest_hit.py --sizes 100000000 --shapes tri_soup --no-oracle --device-lbvh 2>&1 | grep '^{' | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('LBVH', d['config']['workload'], round(d['value'],1), 'Mrays/s', 'create', round(d['config']['scene_create_s'],1))
" | tee -a gpurun_out/r2_ch100m.log
