"""Window-4096 measurements (C5 and batches) for the in-place large-window kernels."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "tools"), os.path.join(ROOT, "phase-vocoder_b200"), os.path.join(ROOT, "tests"), ROOT]
import run_configs as rc
from signals import multitone
noisy = lambda n, s: multitone(n, seed=s, noise=1e-3)
tone = lambda n, s: multitone(n, seed=s, noise=0.0)
semi = lambda k: rc.f32(2 ** (k / 12))
rc.run("C5 long file", 4096, 1024, 1024, "compat", [1.0], 2, 168750, 48000, noisy, check_frames=300)
rc.run("C5 long file", 4096, 1024, 1024, "corrected", [semi(7)], 2, 20000, 48000, tone, check_frames=100)
rc.run("C5 1h corrected", 4096, 1024, 1024, "corrected", [semi(7)], 2, 168750, 48000, noisy, check_frames=100)
rc.run("N=4096 batch", 4096, 1024, 1024, "compat", [1.0], 600, 430, 48000, noisy, check_frames=60)
rc.run("N=4096 batch", 4096, 1024, 1024, "corrected", [semi(7)], 600, 430, 48000, noisy, check_frames=60)
rc.run("N=4096 V=2", 4096, 1024, 1024, "corrected", [1.0, semi(7)], 600, 430, 48000, noisy, check_frames=60)
