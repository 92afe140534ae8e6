"""Rows (f1)/(f2) of SURVEY 8: the C++ WAV codec (host/pv_wav.h) against the numpy restatement of
AudioFile's rules, and the offline driver (host/pv_cli.cpp, semantics of src/main.cpp) end to end."""
import os
import struct
import subprocess

import numpy as np
import pytest

import wav_oracle as wo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "phase-vocoder_b200")
TOOL = os.path.join(ROOT, "tests", "cpp", "wav_tool")
CLI = os.path.join(ROOT, "tests", "cpp", "pv_cli")


@pytest.fixture(scope="module")
def tool():
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", TOOL, os.path.join(ROOT, "tests", "cpp", "wav_tool.cpp")], check=True)
    return TOOL


def _wav24(samples_i32, rate=48000):
    ch, n = samples_i32.shape
    raw = bytearray()
    for i in range(n):
        for c in range(ch):
            v = int(samples_i32[c, i]) & 0xFFFFFF
            raw += bytes((v & 255, (v >> 8) & 255, (v >> 16) & 255))
    hdr = b"RIFF" + struct.pack("<i", 36 + len(raw)) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, ch, rate, ch * rate * 3, ch * 3, 24) + b"data" + struct.pack("<i", len(raw))
    return hdr + bytes(raw) + b"LIST\x04\x00\x00\x00abcd"          # trailing chunk like MAT_ZO_24_bit.wav


def test_cpp_codec_matches_audiofile_rules(tool, tmp_path):
    rng = np.random.default_rng(0)
    x = rng.uniform(-1.3, 1.3, size=(2, 777)).astype(np.float32)
    x[0, :6] = [0.0, 1.0, -1.0, 0.99999, 1e-5, -1e-5]
    x.tofile(tmp_path / "x.f32")
    subprocess.run([tool, "encode", str(tmp_path / "x.f32"), "2", str(tmp_path / "x.wav")], check=True)
    blob = open(tmp_path / "x.wav", "rb").read()
    assert blob == wo.encode_wav16(x)                                  # byte-identical to the numpy restatement
    # decode 16 bit, including a file 2 bytes shorter than its header says (testtones/*sine.wav)
    for data in (blob, blob[:-2]):
        open(tmp_path / "y.wav", "wb").write(data)
        r = subprocess.run([tool, "decode", str(tmp_path / "y.wav"), str(tmp_path / "y.f32")], capture_output=True, text=True, check=True)
        ch, n, rate, bits = map(int, r.stdout.split())
        got = np.fromfile(tmp_path / "y.f32", np.float32).reshape(ch, n)
        want, wrate, wbits = wo.decode_wav(data)
        assert (rate, bits) == (wrate, wbits) and np.array_equal(got, want)
    # 24 bit with a trailing chunk
    s = rng.integers(-2 ** 23, 2 ** 23, size=(2, 100))
    data = _wav24(s)
    open(tmp_path / "z.wav", "wb").write(data)
    r = subprocess.run([tool, "decode", str(tmp_path / "z.wav"), str(tmp_path / "z.f32")], capture_output=True, text=True, check=True)
    got = np.fromfile(tmp_path / "z.f32", np.float32).reshape(2, 100)
    want, _, bits = wo.decode_wav(data)
    assert bits == 24 and np.array_equal(got, want) and np.array_equal(got, (s / 8388608.0).astype(np.float32))
    # float WAVs are rejected like AudioFile.h:454
    bad = bytearray(blob)
    bad[20] = 3
    open(tmp_path / "f.wav", "wb").write(bytes(bad))
    assert subprocess.run([tool, "decode", str(tmp_path / "f.wav"), str(tmp_path / "f.f32")], capture_output=True).returncode == 3


@pytest.mark.gpu
def test_cli_reproduces_golden_head(golden, tmp_path):
    """pv_cli in.wav t out.wav on the head of testtones/test.wav reproduces output/testout.wav +-1 LSB,
    with the reference driver's output format: 2 channels (ch1 := ch0), 44.1 kHz, 16 bit."""
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", CLI, os.path.join(LIBDIR, "host", "pv_cli.cpp"), "-L" + LIBDIR,
                    "-lpv_b200", "-Wl,-rpath," + LIBDIR], check=True)
    xi = golden["testout_head_in"].astype(np.int16)
    want = golden["testout_head_out"].astype(np.int32)
    n = len(want) + 128                      # numSamples = 129 hops -> 128 analysed and 129 synthesised frames
    pcm = np.stack([xi[:n], np.roll(xi[:n], -1)], axis=1).astype("<i2").tobytes()      # exact int16 input
    hdr = b"RIFF" + struct.pack("<i", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, 2, 44100, 2 * 44100 * 2, 4, 16) + b"data" + struct.pack("<i", len(pcm))
    open(tmp_path / "in.wav", "wb").write(hdr + pcm)
    r = subprocess.run([CLI, str(tmp_path / "in.wav"), "t", str(tmp_path / "out.wav")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Offline Vocoding" in r.stdout and "writing to file" in r.stdout
    blob = open(tmp_path / "out.wav", "rb").read()
    assert len(blob) == 44 + n * 2 * 2                                  # 2 channels x 16 bit x timeScale*numSamples
    y, rate, bits = wo.decode_wav(blob)
    assert y.shape == (2, n) and rate == 44100 and bits == 16
    assert np.array_equal(y[0], y[1])                                   # main.cpp:288-290
    got = np.round(y[0].astype(np.float64) * 32768).astype(np.int32)
    # frames 0..127 are identical to the full-file run; frame 128 here is the un-analysed last frame
    assert np.abs(got[:len(want) - 128] - want[:len(want) - 128]).max() <= 1
    assert not got[129 * 128:].any()
