import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]; data=rows[2:]
ci={n:i for i,n in enumerate(h)}
tot=sum(int(r[ci['# Samples']] or 0) for r in data)
print('total samples',tot)
stalls=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg={s:sum(int(r[ci[s]] or 0) for r in data) for s in stalls}
print({k:v for k,v in sorted(agg.items(), key=lambda kv:-kv[1]) if v})
top=sorted(data,key=lambda r:-int(r[ci['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
for r in top:
    st={s:int(r[ci[s]] or 0) for s in stalls if int(r[ci[s]] or 0)}
    print(r[ci['# Samples']].rjust(7), r[ci['Source']][:90].ljust(90), dict(sorted(st.items(), key=lambda kv:-kv[1])[:3]))
