import sys, time
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, synth
from longtr_b200 import Engine
eng=Engine(0)
for kind in (0,1,2):
    r,ms=eng.fp64_issue_rate(kind); print("fp64 kind",kind,"lane-ops/s %.4g"%r,"ms %.2f"%ms, flush=True)
t=time.time(); b=synth.make_pair_batch(99,n_loci=1500,n_lo=50,n_hi=300,reads_lo=4,reads_hi=10,haps_lo=2,haps_hi=6,weird=0.0,sub=0.002,indel=0.002); print("gen",time.time()-t, flush=True)
job=eng.create_job(b)
for i in range(4):
    st=job.run(); print("run",i,"pairs",st.n_pairs,"cells %.3g"%st.n_cells,"ms %.3f"%st.kernel_ms,"GCUPS %.1f"%(st.n_cells/st.kernel_ms/1e6),"fallback",st.n_fallback,"launches",st.n_launches, flush=True)
