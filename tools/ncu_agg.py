import csv,collections,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
for i,r in enumerate(rows):
    if "Kernel Name" in r: hdr=r; rows=rows[i+1:]; break
ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki].split("(")[0][-44:], float(r[vi].replace(",",""))) for r in rows]
agg=collections.OrderedDict()
for k,v in seq: agg.setdefault(k,[]).append(v)
tot=sum(v for _,v in seq)
for k,v in agg.items(): print("%-46s n=%4d total %9.1f us (%.0f%%) mean %8.1f us" % (k, len(v), sum(v)/1000, 100*sum(v)/tot, sum(v)/len(v)/1000))
print("TOTAL us", tot/1000)
