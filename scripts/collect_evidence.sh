#!/bin/bash
# Copies what scripts/evidence.sh left under gpurun_out/ into profiles/ (run here, where the library was built: the
# per-line tables join ncu's source page with nvdisasm of the same build).
R=${1:-r02}; O=gpurun_out; P=profiles
cp $O/${R}_bench_final.json $O/${R}_bench_reference_arm.json $O/${R}_final_launches.csv $P/
grep '^{' $O/${R}_chunk_sweep_1gib.jsonl > $P/${R}_chunk_sweep_1gib.jsonl
ncu -i $O/prof_${R}_final_parse_v7.ncu-rep --page details > /tmp/d1.txt 2>/dev/null; ncu -i $O/prof_${R}_final_decompress_v7.ncu-rep --page details > /tmp/d2.txt 2>/dev/null
cat /tmp/d1.txt /tmp/d2.txt > $P/${R}_decompress_v7_ncu_details.txt
ncu -i $O/prof_${R}_final_parse_v7.ncu-rep --page raw --csv > /tmp/r1.csv 2>/dev/null; ncu -i $O/prof_${R}_final_decompress_v7.ncu-rep --page raw --csv 2>/dev/null | tail -n +3 > /tmp/r2.csv
cat /tmp/r1.csv /tmp/r2.csv > $P/${R}_decompress_v7_ncu_raw.csv
ncu -i $O/prof_${R}_final_compress_c3.ncu-rep --page details > $P/${R}_compress_c3_ncu_details.txt 2>/dev/null
ncu -i $O/prof_${R}_final_compress_c3.ncu-rep --page raw --csv > $P/${R}_compress_c3_ncu_raw.csv 2>/dev/null
python scripts/ncu_by_line.py $O/prof_${R}_final_parse_v7.ncu-rep parse_v7_kernel:parse_v7_kernelILj32ELb0 45 > $P/${R}_parse_v7_by_source_line.txt
python scripts/ncu_by_line.py $O/prof_${R}_final_decompress_v7.ncu-rep decompress_v7:decompress_v7_kernelILb0 60 > $P/${R}_decompress_v7_by_source_line.txt
python scripts/ncu_by_line.py $O/prof_${R}_final_compress_c3.ncu-rep compress_c3:compress_c3_kernelILb0 70 > $P/${R}_compress_c3_by_source_line.txt
cuobjdump -sass lzbench_b200/_lib/liblz4b200.so > /tmp/all.sass
python - <<'EOF'
import re
txt = open('/tmp/all.sass').read()
want = {'compress_c3_kernelILb0': 'compress_c3.sass', 'parse_v7_kernelILj32ELb0': 'parse_v7.sass', 'decompress_v7_kernelILb0': 'decompress_v7.sass',
        'order_v7_kernel': 'order_v7.sass', 'coarse_scan_kernel': 'coarse_scan.sass', 'coarse_copy_kernel': 'coarse_copy.sass'}
for p in re.split(r'(?=\n\s*Function : )', txt):
    m = re.search(r'Function : (\S+)', p)
    if not m: continue
    for k, f in want.items():
        if k in m.group(1): open('profiles/sass/' + f, 'w').write(p.strip() + '\n')
EOF
