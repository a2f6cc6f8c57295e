import sys, math, numpy as np
sys.path.insert(0,'.')
from oracle import oracle
m=oracle.MASS["Ar"]; kB=oracle.K_B
for table,name,omega in ((oracle.PSEUDO_MAXWELL,'maxwell',1.0),(oracle.VHS,'vhs',0.81)):
    it=oracle.interaction("Ar","Ar",table); d=it[3]; Tref=273.0
    n_dens=1e23; n_p=40000; Fnum=n_dens/n_p; T0=273.0
    pv,pia=oracle.OPV(n_p),oracle.OPIA(1,1)
    oracle.sample_equal_weight_cell(oracle.Rng.seq(1),pv,pia,1,1,n_p,m,T0,Fnum)
    rows=pv.logical(1,n_p); rows[:,1]*=1.3; rows[:,2]*=0.8; pv.set_logical(1,rows)
    v=rows[:,1:4]; T=(m*(v**2).mean(0)/kB); Tm=T.mean()
    muref=15*math.sqrt(math.pi*m*kB*Tref)/(2*math.pi*d*d*(5-2*omega)*(7-2*omega))
    mu=muref*(Tm/Tref)**omega
    rate=n_dens*kB*Tm/mu
    cf=oracle.CF(1,oracle.estimate_sigma_g_w_max(it,m,m,Tm,Tm,Fnum)); rng=oracle.Rng.seq(5)
    dt=0.02/rate; a0=T[0]-Tm; out=[]
    for ts in range(1,101):
        oracle.ntc(rng,cf,it,pv,pia,1,1,1,dt,1.0,equal_weight=True)
        v=pv.logical(1,n_p)[:,1:4]; T=(m*(v**2).mean(0)/kB)
        out.append((T[0]-T.mean())/a0)
    t=np.arange(1,101)*0.02
    fit=-np.polyfit(t[:60],np.log(np.array(out[:60])),1)[0]
    print(name,'measured rate / theory (p/mu) =',fit, 'Tm',Tm)
