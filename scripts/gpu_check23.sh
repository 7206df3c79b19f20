#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke$i.log 2>&1; echo "smoke $i exit $?"; tail -1 gpurun_out/smoke$i.log; done
AMOE_STEM_FOLD=0 timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_nofold.log 2>&1; echo "smoke nofold exit $?"; tail -1 gpurun_out/smoke_nofold.log
python - <<'PY'
import sys; sys.path.insert(0,'.')
import torch
from automoe_b200.models.automoe import create_automoe_model
from oracle import automoe_oracle as O, synth
cfg=dict(synth.CONFIG_3EXPERT); m=create_automoe_model(cfg,"cpu"); sd=synth.synth_state_dict(m.state_dict(),0); m.load_state_dict(sd); m=m.to("cuda:0").eval()
sdd={k:v.cuda() for k,v in sd.items()}
def rel(a,b): return ((a.float()-b.float()).abs().max()/b.float().abs().max()).item()
import os
for seed in (1,2,3,4):
  for B,H in ((2,64),(8,256)):
    batch={k:v.cuda() for k,v in synth.synth_batch(B,H,H,seed=seed).items()}
    with torch.no_grad():
        ref=O.automoe_forward(sdd,batch,cfg)
        with torch.autocast("cuda",dtype=torch.bfloat16):
            r16=O.automoe_forward(sdd,batch,cfg)
            res={}
            for fold in ("1","0"):
                os.environ["AMOE_STEM_FOLD"]=fold
                o=m(batch); res[fold]={k:rel(o[k],ref[k]) for k in ("waypoints","speed_seq","gate_logits","combined_features")}
    print(seed,B,H,"fold1",{k:round(v,4) for k,v in res["1"].items()},"fold0",{k:round(v,4) for k,v in res["0"].items()},"ref16",{k:round(rel(r16[k],ref[k]),4) for k in res["1"]})
PY
