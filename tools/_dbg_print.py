import json,sys
d=json.load(open(sys.argv[1]))
out=[]
for l in d['layers']:
    if l['op']=='conv' and l['cin']>=32: out.append("%d>%d:%.3f"%(l['cin'],l['cout'],l['ms']))
print(sys.argv[1], ' '.join(out[::2]))
