#!/bin/bash
# usage: gpu_n2prof.sh tag  -> 2-rank bench where rank 0 runs under ncu (few single-pass metrics on the ADVECT / sort kernels)
tag=$1
cat > /tmp/rankwrap.sh <<'W'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"advect_lean|scatter_kernel|cell_collide|hist|DeviceScan" -s 12 -c 8 --csv --log-file gpurun_out/n2prof_$TAG.csv python bench.py "$@"
else
  exec python bench.py "$@"
fi
W
chmod +x /tmp/rankwrap.sh
TAG=$tag timeout 900 python -m torch.distributed.run --no-python --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 /tmp/rankwrap.sh --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/n2prof_$tag.log 2>&1
tail -3 gpurun_out/n2prof_$tag.log | cut -c1-300
# the same metrics on one rank alone
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none -k regex:"advect_lean|scatter_kernel|cell_collide|hist|DeviceScan" -s 12 -c 8 --csv --log-file gpurun_out/n1prof_$tag.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python - <<P
import csv,collections
for f in ("n2prof_$tag","n1prof_$tag"):
    rows=[r for r in csv.reader(open(f"gpurun_out/{f}.csv")) if len(r)>10 and r[0].isdigit()]
    agg=collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        k=r[4].split("(")[0][-40:]; m=r[-3]; v=float(r[-1].replace(",",""))
        agg[k][m+" "+r[-2]].append(v)
    print(f)
    for k,d in agg.items():
        print("  ",k, {m:round(sum(v)/len(v),3) for m,v in d.items()})
P
