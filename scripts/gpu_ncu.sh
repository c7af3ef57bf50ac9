# ncu on the headline attention shape: launch list (durations) and a full-set capture of the attention kernels
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 20 -c 12 --csv --log-file gpurun_out/launches_attn.csv python bench.py --workload attn --steps 1 --warmup 2 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/ncu_l.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_.*umma -s 4 -c 2 -f -o gpurun_out/prof_attn32_r02 python bench.py --workload attn --steps 1 --warmup 2 --batch 16 --lmax 8192 --attn-dim 32 --attn-heads 8 --no-cpu-baseline > gpurun_out/ncu32.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_attn.csv')) if len(r)>5]
hdr=rows[0]
ki=hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi=hdr.index('Metric Value'); ii=hdr.index('ID')
cur={}
for r in rows[1:]:
    cur.setdefault((r[ii], r[ki][:60]), {})[r[mi]]=r[vi]
for (i,k),m in cur.items():
    print(i, k, m.get('gpu__time_duration.sum'), m.get('dram__bytes_read.sum'), m.get('dram__bytes_write.sum'))
PY
ls -la gpurun_out/prof_attn32_r02.ncu-rep
