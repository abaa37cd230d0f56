set -x
mkdir -p gpurun_out
python bench_closest_hit.py --sizes 1000000 --shapes tri_soup --no-oracle > gpurun_out/r2_ch_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_closest_hit -s 2 -c 1 -o gpurun_out/r2_ch_primary python bench_closest_hit.py --sizes 1000000 --shapes tri_soup --no-oracle > gpurun_out/r2_ch_ncu1.log 2>&1 && \
ncu --set full --clock-control none -k regex:k_closest_hit -s 10 -c 1 -o gpurun_out/r2_ch_incoherent python bench_closest_hit.py --sizes 1000000 --shapes tri_soup --no-oracle > gpurun_out/r2_ch_ncu2.log 2>&1
tail -2 gpurun_out/r2_ch_ncu2.log | cut -c1-300
ls -la gpurun_out
