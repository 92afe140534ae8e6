"""Generates tests/golden/*.npz from the reference checkout (run in the build container only).

    python tests/golden/make_golden.py [/root/reference]

The fixtures are SLICES of the reference's own committed input/output WAV pairs, kept as the
raw 16-bit integers so that nothing in this repo's code has touched them:
  testout_head / testout_tail : testtones/test.wav ch0  -> output/testout.wav   (HEAD config:
                                N=256, Ha=Hs=128, Hamming; src/main.cpp:84, phaseVocoder.h:85-89)
  sine1000_head               : testtones/1000sine.wav ch0 -> output/1000hzout.wav ch0
                                (N=256, Ha=1, Hs=128, symmetric Hann)
plus the sha256 of the four full files, which tests/test_oracle_golden.py re-checks (and then
compares the FULL files) whenever the reference checkout is present.
"""
import hashlib
import os
import struct
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def pcm16(path):
    data = open(path, "rb").read()
    d = data.find(b"data")
    f = data.find(b"fmt")
    ch = struct.unpack_from("<h", data, f + 10)[0]
    (size,) = struct.unpack_from("<i", data, d + 4)
    n = size // (2 * ch)
    raw = data[d + 8:d + 8 + n * 2 * ch]
    raw += b"\0" * (n * 2 * ch - len(raw))
    return np.frombuffer(raw, "<i2").reshape(n, ch).T.copy(), hashlib.sha256(data).hexdigest()


def main():
    tin, h_tin = pcm16(os.path.join(REF, "testtones/test.wav"))
    tout, h_tout = pcm16(os.path.join(REF, "output/testout.wav"))
    sin, h_sin = pcm16(os.path.join(REF, "testtones/1000sine.wav"))
    sout, h_sout = pcm16(os.path.join(REF, "output/1000hzout.wav"))
    N, H = 256, 128
    n = tin.shape[1]
    head_frames = 128
    tail_first = n // H - 64 - 1          # one warm-up frame, then 64 checked frames
    np.savez_compressed(
        os.path.join(HERE, "golden_compat.npz"),
        num_samples=np.int64(n),
        testout_head_in=tin[0, :head_frames * H + N],
        testout_head_out=tout[0, :head_frames * H],
        testout_head_out_ch1=tout[1, :head_frames * H],
        testout_tail_first_frame=np.int64(tail_first),
        testout_tail_in=tin[0, tail_first * H:],
        testout_tail_out=tout[0, (tail_first + 1) * H:],
        sine1000_head_in=sin[0, :head_frames + N],
        sine1000_head_out=sout[0, :head_frames * H],
        sha256=np.array([h_tin, h_tout, h_sin, h_sout]),
    )
    print("wrote golden_compat.npz", os.path.getsize(os.path.join(HERE, "golden_compat.npz")), "bytes")


if __name__ == "__main__":
    main()
