import csv,collections,sys
lines=[l for l in open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/launches.csv') if not l.startswith('==')]
r=csv.DictReader(lines)
rows=collections.OrderedDict()
for row in r:
    key=(row['ID'],row['Kernel Name'][:52])
    rows.setdefault(key,{})[row['Metric Name']]=(float(row['Metric Value'].replace(',','')),row['Metric Unit'])
tot=0
for (id,k),m in list(rows.items()):
    t,u=m['gpu__time_duration.sum']; t=t/1e3 if u=='ns' else (t*1e3 if u=='ms' else t)
    tot+=t
    def gb(x):
        v,u=m[x]; return v*{'byte':1e-9,'Kbyte':1e-6,'Mbyte':1e-3,'Gbyte':1}[u]
    print(f"{t:9.1f} us  rd {gb('dram__bytes_read.sum'):6.3f} wr {gb('dram__bytes_write.sum'):6.3f} GB  fp64 {m['sm__inst_executed_pipe_fp64.sum'][0]/1e6:7.1f}M inst {m['smsp__inst_executed.sum'][0]/1e6:7.1f}M  {k}")
print("total us", tot)
