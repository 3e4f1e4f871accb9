import sys, time, numpy as np
sys.path.insert(0,'.')
from libldpc_b200 import api
from oracle import oracle as O
H="codes/ref_h_n1152_m1024.txt"
c = api.Context(H, "", 0)
oc = O.Code(H)
rng=np.random.default_rng(0)
llr = rng.normal(1.0,1.6,size=(64,1152)); llr[:,oc.puncture]=0
for dec in ("BP_MS","BP"):
    ro,rc,ri = oc.decode(llr,50,True,dec=="BP_MS")
    out,hard,its = c.decode_batch(llr,dec,50,True)
    print(dec,"iters eq",np.array_equal(its,ri),"hard eq",np.array_equal(hard,rc),"llr eq",np.array_equal(out,ro), "maxrel", (np.abs(out-ro)/np.maximum(np.abs(ro),1e-9)).max(), its[:10], ri[:10])
for prec in (api.F64, api.F32):
    c.set_tuning(precision=prec)
    for dec,et in (("BP_MS",False),("BP_MS",True),("BP",False)):
        n = 200000 if dec=="BP_MS" else 40000
        c.sim_point("AWGN",-4.5,nframes=2000,decoding=dec,early_term=et)
        t=time.time(); r=c.sim_point("AWGN",-4.5,nframes=n,decoding=dec,iterations=50,early_term=et); dt=time.time()-t
        print("prec",prec,dec,"et",et,r,"frames/s %.0f"%(n/(r['device_ms']*1e-3)), "Gb/s %.3f"%(n*1024/(r['device_ms']*1e-3)/1e9), "wall",dt, c.stats(reset=True))
