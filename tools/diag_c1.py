"""C1 'pitch x1.5' parity diagnostic: per-block SNR against the fp64 oracle, split and unsplit."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import pv_oracle as po, pvb200
from signals import snr_db
N, H, nf = 256, 64, 2000
x = (0.25 * np.sin(2 * np.pi * 440 * np.arange(N + (6890 - 1) * H) / 44100)).astype(np.float32)
win = po.window(po.WIN_HANN_PERIODIC, N)
w64, _ = po.process_corrected(x, N, H, H, win, [1.5], nf)
w32, _ = po.process_corrected(x, N, H, H, win, [1.5], nf, precision=32)
print("f32 oracle vs f64 oracle", round(snr_db(w64[0], w32[0]), 1))
for env in ({}, {"PV_NO_SPLIT": "1"}, {"PV_FORCE_GENERIC": "1", "PV_NO_SPLIT": "1"}):
    for k in ("PV_NO_SPLIT", "PV_FORCE_GENERIC"):
        os.environ.pop(k, None)
    os.environ.update(env)
    pv = pvb200.PhaseVocoder(N, hop_in=H, hop_out=H, mode=1, window_type=2, pitch=(1.5,))
    xd = torch.from_numpy(np.stack([x, x])).cuda()
    got = pv.process(xd, 6890).cpu().numpy()[0, 0, :nf * H]
    err = got - w64[0]
    blk = [round(snr_db(w64[0][i:i + 200 * H], got[i:i + 200 * H]), 1) for i in range(0, nf * H, 200 * H)]
    print(env, "total", round(snr_db(w64[0], got), 1), "per 200 frames", blk, "max |err|", float(np.abs(err).max()), "at frame", int(np.argmax(np.abs(err))) // H)
