"""Parity diagnostic of the corrected mode: GPU and fp32 oracle against the fp64 oracle on a few shapes."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")]
import pv_oracle as po, pvb200
from signals import multitone, snr_db
SEMI7=float(np.float32(2**(7/12)))
for (N,Ha,Hs,betas,nf,noise) in [(2048,512,512,[SEMI7],60,0.0),(256,64,64,[1.0,float(np.float32(2**(4/12))),SEMI7,2.0],150,0.0),(2048,512,512,[SEMI7],60,1e-3),(2048,512,512,[SEMI7],60,1e-5),(2048,512,512,[SEMI7],60,1e-4)]:
    win=po.window(po.WIN_HANN_PERIODIC,N)
    pv=pvb200.PhaseVocoder(N,hop_in=Ha,hop_out=Hs,mode=1,window_type=2,pitch=tuple(betas))
    for s in range(3):
        x=multitone(N+(nf-1)*Ha-7,seed=50+s,noise=noise)
        got=pv.process(torch.from_numpy(x).cuda()[None,:],nf).cpu().numpy()[0]
        w64,_=po.process_corrected(x,N,Ha,Hs,win,betas,nf)
        w32,_=po.process_corrected(x,N,Ha,Hs,win,betas,nf,precision=32)
        print(N,Ha,Hs,noise,"stream",s,"gpu-vs-f64",[round(snr_db(w64[v],got[v]),1) for v in range(len(betas))],"f32-vs-f64",[round(snr_db(w64[v],w32[v]),1) for v in range(len(betas))])
