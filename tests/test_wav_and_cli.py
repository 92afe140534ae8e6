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


AFTOOL = os.path.join(ROOT, "oracle", "_ref", "audiofile_tool")


@pytest.fixture(scope="module")
def audiofile_tool():
    """The reference's OWN codec (src/AudioFile.h, unmodified, compiled from where it lies by
    oracle/ref_harness/Makefile into git-ignored oracle/_ref/).  The binary travels to the GPU box; the
    reference checkout does not."""
    if os.path.isdir("/root/reference"):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle", "ref_harness"), "../_ref/audiofile_tool"], check=True)
    if not os.path.exists(AFTOOL):
        pytest.skip("oracle/_ref/audiofile_tool not built (needs the reference checkout at build time)")
    return AFTOOL


def _decode_both(tool, aftool, path, tmp_path):
    outs = []
    for t, name in ((tool, "ours.f32"), (aftool, "ref.f32")):
        r = subprocess.run([t, "decode", str(path), str(tmp_path / name)], capture_output=True, text=True)
        if r.returncode != 0:
            outs.append((r.returncode, None, None))
            continue
        ch, n, rate, bits = map(int, r.stdout.split())
        outs.append((0, (ch, n, rate, bits), np.fromfile(tmp_path / name, np.float32).reshape(ch, n)))
    return outs


def test_cpp_codec_matches_audiofile_itself(tool, audiofile_tool, tmp_path):
    """host/pv_wav.h against src/AudioFile.h ITSELF (:418-530 decode, :703-785 + :1045-1049 encode): same
    accept/reject decision, same header fields, bit-identical samples and byte-identical 16-bit files."""
    rng = np.random.default_rng(3)
    # encode: random floats incl. out-of-range values and the clamp / truncation corner cases
    x = rng.uniform(-1.3, 1.3, size=(2, 1201)).astype(np.float32)
    x[1, :8] = [0.0, 1.0, -1.0, 0.99999, -0.99999, 1e-5, -1e-5, 0.5]
    x.tofile(tmp_path / "x.f32")
    subprocess.run([tool, "encode", str(tmp_path / "x.f32"), "2", str(tmp_path / "ours.wav")], check=True)
    subprocess.run([audiofile_tool, "encode", str(tmp_path / "x.f32"), "2", str(tmp_path / "ref.wav")], check=True)
    assert open(tmp_path / "ours.wav", "rb").read() == open(tmp_path / "ref.wav", "rb").read()
    # decode: synthetic 16-bit stereo/mono, 24-bit with a trailing chunk, 8-bit
    files = {"s16.wav": open(tmp_path / "ours.wav", "rb").read(),
             "s24.wav": _wav24(rng.integers(-2 ** 23, 2 ** 23, size=(2, 333))),
             "m24.wav": _wav24(rng.integers(-2 ** 23, 2 ** 23, size=(1, 50)), rate=44100)}
    u8 = rng.integers(0, 256, size=200, dtype=np.uint8).tobytes()
    files["u8.wav"] = b"RIFF" + struct.pack("<i", 36 + len(u8)) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, 1, 8000, 8000, 1, 8) + b"data" + struct.pack("<i", len(u8)) + u8
    bad = bytearray(files["s16.wav"]); bad[20] = 3                       # IEEE float format tag: AudioFile.h:454
    files["float.wav"] = bytes(bad)
    bad = bytearray(files["s16.wav"]); bad[22] = 3                       # three channels: AudioFile.h:461
    files["three.wav"] = bytes(bad)
    for name, blob in files.items():
        open(tmp_path / name, "wb").write(blob)
        ours, ref = _decode_both(tool, audiofile_tool, tmp_path / name, tmp_path)
        assert (ours[0] == 0) == (ref[0] == 0), name
        if ref[0] == 0:
            assert ours[1] == ref[1], name
            assert np.array_equal(ours[2], ref[2]), name


def test_cpp_codec_on_the_reference_wavs(tool, audiofile_tool, reference_dir, tmp_path):
    """Every WAV of the reference's testtones/ (incl. the 24-bit MAT_ZO file with trailing LIST/id3 chunks, the
    2-bytes-short *sine/*square files and the float WAVs AudioFile rejects) through both codecs."""
    tdir = os.path.join(reference_dir, "testtones")
    seen = 0
    for name in sorted(os.listdir(tdir)):
        if not name.endswith(".wav"):
            continue
        ours, ref = _decode_both(tool, audiofile_tool, os.path.join(tdir, name), tmp_path)
        assert (ours[0] == 0) == (ref[0] == 0), name
        if ref[0] != 0:
            assert "FLOAT" in name or "float" in name, name               # only the fmt-3 files are rejected
            continue
        assert ours[1] == ref[1], name
        size = os.path.getsize(os.path.join(tdir, name))
        ch, n, _, bits = ref[1]
        short = size < 44 + n * ch * bits // 8                            # the reference reads past its buffer here (UB):
        k = n - 1 if short else n                                         # the frames that exist must agree; ours zero-fills
        assert np.array_equal(ours[2][:, :k], ref[2][:, :k]), name
        seen += 1
    assert seen >= 10


def test_wav_loader_rejects_malformed_files(tool, tmp_path):
    """the loader must not trust the file (no out-of-bounds header reads, no huge allocations)."""
    good = wo.encode_wav16(np.zeros((2, 64), np.float32))
    cases = {
        "trunc_fmt": good[:12] + b"fmt " + b"\x10\0\0\0" + b"\x01\0" + b"data" + b"\0" * 18,     # fmt fields cut short
        "fmt_at_end": good[:12] + b"data" + struct.pack("<I", 8) + b"\0" * 28 + b"fmt",
        "streaming_size": good[:40] + struct.pack("<I", 0xFFFFFFFF) + good[44:],                    # 4 GB claimed
        "huge_size": good[:40] + struct.pack("<I", 0x7FFFFFF0) + good[44:],
        "data_at_end": good[:36] + b"\0" * 8 + good[44:] + b"data\x01\0",
    }
    for name, blob in cases.items():
        open(tmp_path / (name + ".wav"), "wb").write(blob)
        r = subprocess.run([tool, "decode", str(tmp_path / (name + ".wav")), str(tmp_path / "o.f32")], capture_output=True, text=True)
        assert r.returncode == 3, (name, r.returncode, r.stderr)                                    # an error, not a crash
    # a file a little shorter than its header claims is still accepted and zero-filled (testtones/440sine.wav)
    open(tmp_path / "short.wav", "wb").write(good[:-2])
    r = subprocess.run([tool, "decode", str(tmp_path / "short.wav"), str(tmp_path / "o.f32")], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.split()[1] == "64"


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


@pytest.mark.gpu
def test_cli_with_the_reference_audiofile_writes_the_same_file(golden_wav, tmp_path):
    """North star: "WAV in and out via AudioFile".  oracle/_ref/pv_cli_audiofile is host/pv_cli.cpp compiled against the
    reference's OWN, unmodified src/AudioFile.h (used exactly as src/main.cpp:128-143,309 uses it); the default build uses
    the built-in codec.  Both must write byte-identical files -- 16-bit C1 input (440sine.wav slice, window 256, hop
    divisor 4, compat and pitch x1.5) and 24-bit C2 input with a trailing chunk (MAT_ZO slice, window 2048, +7 semitones)."""
    exe_af = os.path.join(ROOT, "oracle", "_ref", "pv_cli_audiofile")
    if not os.path.exists(exe_af):
        pytest.skip("oracle/_ref/pv_cli_audiofile not built (needs the reference checkout at build time)")
    subprocess.run(["g++", "-std=c++17", "-O1", "-o", CLI, os.path.join(LIBDIR, "host", "pv_cli.cpp"), "-L" + LIBDIR,
                    "-lpv_b200", "-Wl,-rpath," + LIBDIR], check=True)
    c1 = golden_wav["c1_raw"]                                              # int16 [2, n]
    pcm = c1.T.astype("<i2").tobytes()
    w16 = b"RIFF" + struct.pack("<i", 36 + len(pcm)) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, 2, 44100, 2 * 44100 * 2, 4, 16) + b"data" + struct.pack("<i", len(pcm)) + pcm
    open(tmp_path / "c1.wav", "wb").write(w16)
    open(tmp_path / "c2.wav", "wb").write(_wav24(golden_wav["c2_raw"], rate=44100))
    runs = [("c1.wav", "t", ["--window", "256", "--hop-div", "4"]), ("c1.wav", "p", ["--window", "256", "--hop-div", "4", "--scale", "1.5"]),
            ("c2.wav", "p", ["--window", "2048", "--hop-div", "4", "--scale", "1.4983071"])]
    for i, (wav, eff, extra) in enumerate(runs):
        outs = []
        for exe, tag in ((CLI, "own"), (exe_af, "af")):
            o = tmp_path / f"out{i}_{tag}.wav"
            r = subprocess.run([exe, str(tmp_path / wav), eff, str(o)] + extra, capture_output=True, text=True)
            assert r.returncode == 0, (exe, r.stdout, r.stderr)
            outs.append(open(o, "rb").read())
        assert outs[0] == outs[1], (wav, eff)
        y, rate, bits = wo.decode_wav(outs[0])
        assert y.shape[0] == 2 and rate == 44100 and bits == 16 and np.abs(y).max() > 0.01
