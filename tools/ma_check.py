import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import amici_b200 as ab
from oracle.refsim import RefModel
from amici_b200.codegen.massaction import network
name=sys.argv[1]; N=int(sys.argv[2])
net=network(name)
model=ab.get_model(name)
ts=np.linspace(0,10,11)
model.set_timepoints(ts)
solver=model.create_solver()
solver.set_relative_tolerance(1e-8); solver.set_absolute_tolerance(1e-12)
solver.set_sensitivity_order(ab.SensitivityOrder.first)
rng=np.random.default_rng(1)
p=np.tile(net["p_nominal"],(N,1))*10**rng.uniform(-1,1,(N,net["nr"]))
k=np.ones((N,1))
my=rng.uniform(0.5,1.5,(N,11,net["ny"])); sig=np.full_like(my,0.1)
batch=ab.ExpDataBatch(ts,p,k,my=my,sigmay=sig)
t0=time.time(); res=ab.run_simulations(model,solver,batch); t1=time.time()
print("gpu wall",t1-t0,"kernel ms",res.kernel_ms,"status",res["status"][:8], "steps", res["numsteps"][:4,-1])
ref=RefModel(name)
t0=time.time(); r=ref.run(ts,p,k,my=my,sigmay=sig,sensi_order=1,rtol=1e-8,atol=1e-12,num_threads=16); print("ref wall",time.time()-t0, "steps", r["numsteps"][:4,-1])
for f in ("x","y","llh","sllh","sx","sy","res","sres","FIM","chi2","x0","xdot","J","sigmay"):
    a,b=np.asarray(res[f]),np.asarray(r[f]).reshape(np.asarray(res[f]).shape)
    err=np.abs(a-b); scale=10*1e-12+10*1e-8*np.abs(b)
    print(f, "max abs", err.max(), "max ratio to 10rtol/10atol", (err/scale).max())
