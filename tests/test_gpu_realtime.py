"""Real-time block server (pv_rt_*): the reference's callback contract (src/main.cpp:45-59) for many channels.
Block-by-block output == offline output of the input delayed by the latency N-Ha, bit for bit."""
import numpy as np
import pytest
import torch

import pvb200
from signals import multitone

pytestmark = pytest.mark.gpu


def f32(x):
    return float(np.float32(x))


def offline(pv, x, latency, n_frames):
    xd = np.concatenate([np.zeros((x.shape[0], latency), np.float32), x], axis=1)
    xd = np.pad(xd, ((0, 0), (0, -xd.shape[1] % 4)))
    return pv.process(torch.from_numpy(xd).cuda(), n_frames, n_in=latency + x.shape[1]).cpu().numpy()


@pytest.mark.parametrize("mode,N,Ha,Hs,B,S", [
    ("compat", 256, 128, 128, 2, 3),          # the reference's head config: nBufferFrames = nSamps = 256
    ("compat", 1024, 256, 512, 1, 5),
    ("corrected", 256, 64, 64, 1, 7),         # C4's shape, one hop per block
    ("corrected", 2048, 512, 512, 3, 2),
    ("corrected", 512, 100, 100, 4, 2),       # hop not a multiple of 4: unaligned ring rows
    ("corrected", 256, 64, 64, 96, 1),        # one long block of one stream
    ("corrected4", 256, 64, 64, 1, 5),        # C4: four voices; a short block is ONE launch, the offline run two launches of two voices
    ("corrected4", 512, 128, 128, 20, 2),     # four voices, blocks long enough for the voice-pair launches inside the graph
])
def test_blocks_equal_offline_delayed_input(mode, N, Ha, Hs, B, S):
    blocks = 9
    if mode == "compat":
        pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_COMPAT)
    else:
        pitch = (1.0, f32(1.5)) if mode == "corrected" else (1.0, f32(2 ** (4 / 12)), f32(2 ** (7 / 12)), 2.0)
        pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED,
                                 window_type=pvb200.WIN_HANN_PERIODIC, pitch=pitch)
    x = np.stack([multitone(blocks * B * Ha, seed=70 + s) for s in range(S)])
    rt = pvb200.RealtimeServer(pv, S, B)
    assert rt.latency == N - Ha and rt.block_in == B * Ha and rt.block_out == B * Hs
    want = offline(pv, x, rt.latency, blocks * B)
    got = []
    for b in range(blocks):
        rt.input[:] = x[:, b * B * Ha:(b + 1) * B * Ha]
        got.append(rt.step().copy())
    got = np.concatenate(got, axis=2)
    assert np.array_equal(got, want)
    # a reset starts a new take; the callback form gives the same samples
    rt.reset()
    out = np.empty((S, pv.n_voices, B * Hs), np.float32)
    for b in range(2):
        assert rt.callback(out, np.ascontiguousarray(x[:, b * B * Ha:(b + 1) * B * Ha])) == 0
        assert np.array_equal(out, want[:, :, b * B * Hs:(b + 1) * B * Hs])
    with pytest.raises(pvb200.PvError):
        rt.callback(out, np.zeros((S, B * Ha + 1), np.float32))
    n0 = pv.launch_count()
    rt.step()
    assert pv.launch_count() == n0 + 1           # one fused kernel per block
    rt.close()


@pytest.mark.parametrize("mode", ["compat", "corrected"])
def test_server_survives_other_shapes_on_the_same_handle(mode):
    """the recorded graphs bake device pointers in, so the server must own everything they reference.
    Between two blocks the SAME handle runs 40 other shapes (more than the plan cache holds, incl. few-stream
    corrected runs that take the frame-range split and reallocate its scratch); the blocks must still equal the
    offline output bit for bit."""
    N, Ha, Hs, B, S, blocks = 512, 128, 128, 2, 3, 6
    if mode == "compat":
        pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_COMPAT)
    else:
        pv = pvb200.PhaseVocoder(N, hop_in=Ha, hop_out=Hs, mode=pvb200.MODE_CORRECTED,
                                 window_type=pvb200.WIN_HANN_PERIODIC, pitch=(f32(1.26),))
    x = np.stack([multitone(blocks * B * Ha, seed=90 + s) for s in range(S)])
    rt = pvb200.RealtimeServer(pv, S, B)
    want = offline(pv, x, rt.latency, blocks * B)
    noise = torch.randn((5, N + 4000 * Ha), device="cuda") * 0.1
    got, shape = [], 0
    for b in range(blocks):
        rt.input[:] = x[:, b * B * Ha:(b + 1) * B * Ha]
        got.append(rt.step().copy())
        for _ in range(8):                                   # 8 new shapes per block: 48 in all, cache holds 32
            shape += 1
            pv.process(noise[:1 + shape % 5], 20 + 3 * shape)
        pv.process(noise[:1], 1500 + 100 * b)                # one long stream: frame-range split in corrected mode
    torch.cuda.synchronize()
    assert np.array_equal(np.concatenate(got, axis=2), want)
    rt.close()
