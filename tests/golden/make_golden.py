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

golden_wav.npz holds raw-integer slices of the two input WAVs BASELINE.json's configs name (the reference has
no output for them: pitch shift is not implemented there), so that C1 / C2 run on the reference's real
samples everywhere, also on the GPU box where the checkout does not exist:
  c1_440sine   : testtones/440sine.wav, int16, both channels, first 512 frames at window 256 / hop 64
  c2_matzo     : testtones/MAT_ZO_24_bit.wav, 24-bit as int32, both channels, 64 frames at window 2048 /
                 hop 512 starting 1 s into the file
and the total sample counts, so that the tests can check the slices against the full decode when the
checkout is present.
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


def pcm24(path):
    data = open(path, "rb").read()
    d = data.find(b"data")
    f = data.find(b"fmt")
    ch = struct.unpack_from("<h", data, f + 10)[0]
    (size,) = struct.unpack_from("<i", data, d + 4)
    n = size // (3 * ch)
    b = np.frombuffer(data[d + 8:d + 8 + n * 3 * ch], np.uint8).reshape(-1, 3).astype(np.int32)
    v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
    v = np.where(v & 0x800000, v | ~0xFFFFFF, v).astype(np.int32)
    return v.reshape(n, ch).T.copy(), hashlib.sha256(data).hexdigest()


def wav_slices():
    c1, h1 = pcm16(os.path.join(REF, "testtones/440sine.wav"))
    c2, h2 = pcm24(os.path.join(REF, "testtones/MAT_ZO_24_bit.wav"))
    n1 = 256 + 511 * 64
    o2, n2 = 44100, 2048 + 63 * 512
    np.savez_compressed(
        os.path.join(HERE, "golden_wav.npz"),
        c1_440sine=c1[:, :n1], c1_num_samples=np.int64(c1.shape[1]),
        c2_matzo=c2[:, o2:o2 + n2], c2_offset=np.int64(o2), c2_num_samples=np.int64(c2.shape[1]),
        sha256=np.array([h1, h2]),
    )
    print("wrote golden_wav.npz", os.path.getsize(os.path.join(HERE, "golden_wav.npz")), "bytes")


def main():
    wav_slices()
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
