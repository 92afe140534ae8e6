"""WAV sample <-> float rules of the reference's AudioFile.h, restated in numpy.

TEST INFRASTRUCTURE ONLY.  Follows src/AudioFile.h:
  decodeWaveFile  :418-530  chunk discovery by FIRST substring match of "data"/"fmt"
                            (getIndexOfString :1017-1034), PCM 8/16/24, mono/stereo only
  sixteenBitIntToSample :1038-1042  s / 32768
  24-bit          :508-518  sign-extend, / 8388608
  saveToWaveFile  :703-785  44-byte header, interleaved
  sampleToSixteenBitInt :1045-1049  (int16) trunc(clamp(x,-1,1) * 32767)
"""
from __future__ import annotations

import struct

import numpy as np


def decode_wav(data: bytes):
    """-> (samples float32 [channels, n], sample_rate, bit_depth). Raises ValueError where
    AudioFile::load returns false."""
    if data[0:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError("not a RIFF/WAVE file")
    d = data.find(b"data")
    f = data.find(b"fmt")
    if d < 0 or f < 0:
        raise ValueError("missing chunk")
    audio_format, channels = struct.unpack_from("<hh", data, f + 8)
    rate, bytes_per_sec = struct.unpack_from("<ii", data, f + 12)
    block, bits = struct.unpack_from("<hh", data, f + 20)
    if audio_format != 1:
        raise ValueError("not PCM (AudioFile.h:454)")
    if channels < 1 or channels > 2:
        raise ValueError("neither mono nor stereo")
    nbytes = bits // 8
    if bytes_per_sec != channels * rate * bits // 8 or block != channels * nbytes:
        raise ValueError("inconsistent header")
    if bits not in (8, 16, 24, 32):
        raise ValueError("bad bit depth")
    (chunk,) = struct.unpack_from("<i", data, d + 4)
    n = chunk // (channels * bits // 8)
    start = d + 8
    need = start + n * block
    raw = data[start:need]
    if len(raw) < n * block:          # file shorter than its header says: the reference reads past
        raw = raw + b"\0" * (n * block - len(raw))   # the vector (UB); we define the missing bytes as 0
    if bits == 16:
        s = np.frombuffer(raw, "<i2").astype(np.float32) / np.float32(32768.0)
    elif bits == 24:
        b = np.frombuffer(raw, np.uint8).reshape(-1, 3).astype(np.int32)
        v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
        v = np.where(v & 0x800000, v | ~0xFFFFFF, v)
        s = v.astype(np.float32) / np.float32(8388608.0)
    elif bits == 8:
        s = (np.frombuffer(raw, np.uint8).astype(np.int32) - 128).astype(np.float32) / np.float32(128.0)
    else:
        raise ValueError("32-bit decode is a no-op in the reference (AudioFile.h:520-524)")
    return s.reshape(n, channels).T.copy(), rate, bits


def float_to_s16(x: np.ndarray) -> np.ndarray:
    """AudioFile::sampleToSixteenBitInt (AudioFile.h:1045-1049) on float samples."""
    x = np.asarray(x, np.float32)
    x = np.minimum(x, np.float32(1.0))
    x = np.maximum(x, np.float32(-1.0))
    return np.trunc(x.astype(np.float64) * 32767.0).astype(np.int16)


def s24_to_bytes(v: np.ndarray) -> np.ndarray:
    """int32 sample values (24 significant bits, sign-extended) -> [..., 3] little-endian bytes as stored in a WAV."""
    u = np.asarray(v, np.int64) & 0xFFFFFF
    return np.stack([u & 0xFF, (u >> 8) & 0xFF, (u >> 16) & 0xFF], axis=-1).astype(np.uint8)


def bytes_to_s24(b: np.ndarray) -> np.ndarray:
    """[..., 3] little-endian bytes -> sign-extended int32 (AudioFile.h:508-515)."""
    b = np.asarray(b, np.uint8).astype(np.int32)
    v = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16)
    return np.where(v & 0x800000, v | ~0xFFFFFF, v).astype(np.int32)


def float_to_s24(x: np.ndarray) -> np.ndarray:
    """The 24-bit branch of AudioFile::saveToWaveFile (AudioFile.h:755-757): (int32)(x * 8388608.), NO clamp; the three low
    bytes are what reaches the file.  Returns the int32 values (callers take `s24_to_bytes`)."""
    x = np.asarray(x, np.float32)
    return np.trunc(x.astype(np.float64) * 8388608.0).astype(np.int64).astype(np.int32)


def encode_wav16(samples: np.ndarray, rate: int = 44100) -> bytes:
    """samples [channels, n] float -> bytes exactly as AudioFile::saveToWaveFile, 16-bit."""
    ch, n = samples.shape
    pcm = float_to_s16(samples).T.reshape(-1)
    size = n * ch * 2
    hdr = b"RIFF" + struct.pack("<i", 4 + 24 + 8 + size) + b"WAVE" + b"fmt " + struct.pack(
        "<ihhiihh", 16, 1, ch, rate, ch * rate * 16 // 8, ch * 2, 16) + b"data" + struct.pack("<i", size)
    return hdr + pcm.astype("<i2").tobytes()
