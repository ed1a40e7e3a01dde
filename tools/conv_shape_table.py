import sys, csv, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from humanliff_b200 import factory
from humanliff_b200.unet import _StepPlan
m,d = factory.create_model_and_diffusion(**factory.production_flags(""))
cpu=torch.device("cpu")
m._pack(cpu)
plan=_StepPlan(m,cpu,4,256,256)
convs=[a for n,a,_b in plan.calls if n=="hl_conv2d"]
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get("Metric Name")=="gpu__time_duration.sum" and "k_conv_tc" in r["Kernel Name"]]
print(len(convs), len(rows))
from collections import defaultdict
agg=defaultdict(lambda:[0,0.0,0.0])
for a,r in zip(convs,rows):
    B,H,W,Cin,Cout,k,s = a[11:18]
    res = a[5] is not None; st = a[9] is not None
    us=float(r["Metric Value"].replace(",",""))/1e3
    Ho,Wo=H//s,W//s
    fl=2.0*B*Ho*Wo*Cout*Cin*k*k
    key=(H,Cin,Cout,k,s,res)
    agg[key][0]+=1; agg[key][1]+=us; agg[key][2]+=fl
tot=sum(v[1] for v in agg.values())
print("H Cin Cout k s res | n | us each | TFLOP/s | share")
for key,(n,us,fl) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print(key, n, "%.1f"%(us/n), "%.0f"%(fl/us/1e6), "%.1f%%"%(100*us/tot))
print("total us", tot)
