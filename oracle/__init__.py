"""oracle/ -- CPU restatement of the reference's streaming-inference path (TEST INFRASTRUCTURE).

*** PARITY ***: the reference's arithmetic for this path lives in pymicro-features / TensorFlow-Lite
native code that is absent from /root/reference and not installable here, and the reference ships no
golden vectors.  Micro-frontend: pinned to the upstream library's own unit-test vectors
(tests/test_upstream_kat.py - all nine stages + consecutive frames reproduced exactly) and closed-form
known answers.  MixedNet / TFLite int8 kernels: UNPINNED (closed-form known answers only).

Import rules: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package.  The product package (microwakeword_b200/) never does.

    oracle.lib()          -> ctypes handle to oracle/_build/libmwwo.so (built by `make -C oracle`)
    oracle.Frontend       -> stateful C micro-frontend (frontend.c)
    oracle.MixedNet       -> C streaming MixedNet, fp32 or int8 (mixednet.c)
    oracle.run_pipeline   -> multi-threaded audio -> features -> probabilities over many streams
    oracle.mixednet_ref   -> NumPy restatement + synthetic weights + int8 quantiser
"""

from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmwwo.so")
_lib = None

NUM_CHANNELS = 40


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("frontend.c", "mixednet.c", "pipeline.c", "frontend.h", "mixednet.h", "Makefile")]
    stale = force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs if os.path.exists(s))
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        vp, sz, i32 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
        L.mwwo_frontend_create.restype = vp
        L.mwwo_frontend_create_cfg.restype = vp
        L.mwwo_frontend_create_cfg.argtypes = [i32, i32, i32, i32, ctypes.c_float, ctypes.c_float]
        L.mwwo_frontend_config.argtypes = [vp, vp]
        L.mwwo_frontend_window_tap.argtypes = [vp, vp, vp]
        L.mwwo_frontend_free.argtypes = [vp]
        L.mwwo_frontend_reset.argtypes = [vp]
        L.mwwo_frontend_process.restype = i32
        L.mwwo_frontend_process.argtypes = [vp, vp, sz, ctypes.POINTER(sz), vp]
        L.mwwo_generate_features.restype = sz
        L.mwwo_generate_features.argtypes = [vp, sz, vp, sz]
        L.mwwo_frontend_stream.restype = sz
        L.mwwo_frontend_stream.argtypes = [vp, vp, sz, vp, sz]
        L.mwwo_frontend_tables.argtypes = [vp] * 11
        L.mwwo_frontend_taps.argtypes = [vp] * 10
        L.mwwo_frontend_get_state.argtypes = [vp] * 4
        L.mwwo_sqrt64.restype = ctypes.c_uint32
        L.mwwo_sqrt64.argtypes = [ctypes.c_uint64]
        L.mwwo_wdf.restype = ctypes.c_int32
        L.mwwo_wdf.argtypes = [vp, ctypes.c_uint32]
        L.mwwo_pcan_shrink.restype = ctypes.c_uint32
        L.mwwo_pcan_shrink.argtypes = [ctypes.c_uint32]
        L.mwwo_log_scaled.restype = ctypes.c_uint32
        L.mwwo_log_scaled.argtypes = [vp, ctypes.c_uint32]
        L.mwwo_fftr.argtypes = [vp, vp, vp]
        L.mwwo_mixednet_create.restype = vp
        L.mwwo_mixednet_create.argtypes = [ctypes.c_char_p, sz]
        L.mwwo_mixednet_free.argtypes = [vp]
        L.mwwo_mixednet_reset.argtypes = [vp]
        L.mwwo_mixednet_is_quantized.argtypes = [vp]
        L.mwwo_mixednet_stride.argtypes = [vp]
        L.mwwo_mixednet_input_scale.restype = ctypes.c_float
        L.mwwo_mixednet_input_scale.argtypes = [vp]
        L.mwwo_mixednet_input_zero_point.argtypes = [vp]
        L.mwwo_mixednet_step_f32.restype = ctypes.c_float
        L.mwwo_mixednet_step_f32.argtypes = [vp, vp, vp]
        L.mwwo_mixednet_step_int8.restype = i32
        L.mwwo_mixednet_step_int8.argtypes = [vp, vp, vp]
        L.mwwo_mixednet_predict_u16.restype = sz
        L.mwwo_mixednet_predict_u16.argtypes = [vp, vp, sz, vp, sz]
        L.mwwo_pipeline_run.restype = i32
        L.mwwo_pipeline_run.argtypes = [ctypes.c_char_p, sz, vp, sz, sz, vp, sz, vp, sz, vp, vp, i32, i32]
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


class Frontend:
    """Stateful micro-frontend, one stream (pymicro_features.MicroFrontend stand-in)."""

    def __init__(self, config=None):
        """config=None: the okay_nabu configuration (audio_utils.py:71-78); otherwise a tuple
        (sample_rate, window_ms, step_ms, num_channels, lower_hz, upper_hz) -- used to run the upstream
        micro-frontend library's own unit-test configuration through the same code (tests/test_upstream_kat.py)."""
        self._L = lib()
        self._h = self._L.mwwo_frontend_create() if config is None else self._L.mwwo_frontend_create_cfg(*config)
        if not self._h:
            raise RuntimeError("oracle frontend: table construction failed")
        cfg = np.zeros(8, np.int32)
        self._L.mwwo_frontend_config(self._h, cfg.ctypes.data)
        self.sample_rate, self.window, self.step, self.fft_size, self.num_channels, self.start_index, self.end_index = (int(v) for v in cfg[:7])

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.mwwo_frontend_free(self._h)
            self._h = None

    def reset(self):
        self._L.mwwo_frontend_reset(self._h)

    def process_samples(self, samples: np.ndarray):
        """One ProcessSamples call -> (features uint16[40] or None, samples_read)."""
        samples = np.ascontiguousarray(samples, np.int16)
        out = np.zeros(NUM_CHANNELS, np.uint16)
        n_read = ctypes.c_size_t(0)
        got = self._L.mwwo_frontend_process(self._h, samples.ctypes.data, samples.size, ctypes.byref(n_read), out.ctypes.data)
        return (out[:self.num_channels] if got else None), n_read.value

    def stream(self, audio: np.ndarray) -> np.ndarray:
        audio = np.ascontiguousarray(audio, np.int16)
        out = np.zeros((audio.size // self.step + 1, self.num_channels), np.uint16)
        n = self._L.mwwo_frontend_stream(self._h, audio.ctypes.data, audio.size, out.ctypes.data, out.shape[0])
        return out[:n]

    def tables(self) -> dict:
        t = dict(window=np.zeros(512, np.int16), bin_channel=np.zeros(257, np.int16), bin_weight=np.zeros(257, np.int16),
                 bin_unweight=np.zeros(257, np.int16), chan_start=np.zeros(42, np.int16), gain_lut=np.zeros(125, np.int16),
                 log_lut=np.zeros(129, np.uint16), twiddles=np.zeros((256, 2), np.int16), super_twiddles=np.zeros((128, 2), np.int16),
                 scalars=np.zeros(8, np.int32))
        self._L.mwwo_frontend_tables(self._h, *[_ptr(t[k]) for k in ("window", "bin_channel", "bin_weight", "bin_unweight", "chan_start",
                                                                      "gain_lut", "log_lut", "twiddles", "super_twiddles", "scalars")])
        nb, nc = self.fft_size // 2 + 1, self.num_channels
        t["window"] = t["window"][:self.window]
        for k in ("bin_channel", "bin_weight", "bin_unweight"):
            t[k] = t[k][:nb]
        t["chan_start"] = t["chan_start"][:nc + 2]
        t["twiddles"] = t["twiddles"][:self.fft_size // 2]
        t["super_twiddles"] = t["super_twiddles"][:self.fft_size // 4]
        return t

    def taps(self) -> dict:
        t = dict(shift=np.zeros(1, np.int32), fft_in=np.zeros(512, np.int16), fft_out=np.zeros((257, 2), np.int16),
                 energy=np.zeros(257, np.uint32), work=np.zeros(41, np.uint64), sqrt=np.zeros(40, np.uint32),
                 nr=np.zeros(40, np.uint32), pcan=np.zeros(40, np.uint32), estimate=np.zeros(40, np.uint32))
        self._L.mwwo_frontend_taps(self._h, *[_ptr(t[k]) for k in ("shift", "fft_in", "fft_out", "energy", "work", "sqrt", "nr", "pcan", "estimate")])
        nb, nc = self.fft_size // 2 + 1, self.num_channels
        t["fft_in"] = t["fft_in"][:self.fft_size]
        t["fft_out"], t["energy"] = t["fft_out"][:nb], t["energy"][:nb]
        t["work"] = t["work"][:nc + 1]
        for k in ("sqrt", "nr", "pcan", "estimate"):
            t[k] = t[k][:nc]
        win = np.zeros(512, np.int16)
        mx = np.zeros(1, np.int32)
        self._L.mwwo_frontend_window_tap(self._h, win.ctypes.data, mx.ctypes.data)
        t["window_out"], t["max_abs"] = win[:self.window], int(mx[0])
        return t

    def state(self):
        buf, used, est = np.zeros(512, np.int16), np.zeros(1, np.int32), np.zeros(40, np.uint32)
        self._L.mwwo_frontend_get_state(self._h, buf.ctypes.data, used.ctypes.data, est.ctypes.data)
        return buf[:self.window], int(used[0]), est[:self.num_channels]

    def fftr(self, x512: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x512, np.int16)
        assert x.size == self.fft_size
        out = np.zeros((257, 2), np.int16)
        self._L.mwwo_fftr(self._h, x.ctypes.data, out.ctypes.data)
        return out[:self.fft_size // 2 + 1]

    def wdf(self, x: int) -> int:
        return self._L.mwwo_wdf(self._h, x)

    def log_scaled(self, x: int) -> int:
        return self._L.mwwo_log_scaled(self._h, x)


def generate_features_for_clip(audio: np.ndarray) -> np.ndarray:
    """audio_utils.generate_features_for_clip(use_c=True) restated; returns uint16 [T, 40]
    (the reference returns these values * 0.0390625 as float32)."""
    if audio.dtype in (np.float32, np.float64):
        audio = np.clip(audio * 32768, -32768, 32767).astype(np.int16)   # audio_utils.py:47-48
    audio = np.ascontiguousarray(audio, np.int16)
    out = np.zeros((audio.size // 160 + 1, NUM_CHANNELS), np.uint16)
    n = lib().mwwo_generate_features(audio.ctypes.data, audio.size, out.ctypes.data, out.shape[0])
    return out[:n]


class MixedNet:
    """C streaming MixedNet over an MWW container (bytes)."""

    def __init__(self, blob: bytes):
        self._L = lib()
        self._h = self._L.mwwo_mixednet_create(blob, len(blob))
        if not self._h:
            raise ValueError("oracle mixednet: malformed or incomplete model container")
        self.is_quantized = bool(self._L.mwwo_mixednet_is_quantized(self._h))
        self.stride = self._L.mwwo_mixednet_stride(self._h)
        self.input_scale = self._L.mwwo_mixednet_input_scale(self._h)
        self.input_zero_point = self._L.mwwo_mixednet_input_zero_point(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.mwwo_mixednet_free(self._h)
            self._h = None

    def reset(self):
        self._L.mwwo_mixednet_reset(self._h)

    def step_f32(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.float32)
        logit = ctypes.c_float(0)
        p = self._L.mwwo_mixednet_step_f32(self._h, x.ctypes.data, ctypes.addressof(logit))
        return np.float32(p), np.float32(logit.value)

    def step_int8(self, x: np.ndarray):
        x = np.ascontiguousarray(x, np.int8)
        logit = ctypes.c_int(0)
        out = self._L.mwwo_mixednet_step_int8(self._h, x.ctypes.data, ctypes.addressof(logit))
        return int(out), int(logit.value)

    def predict_u16(self, feat: np.ndarray) -> np.ndarray:
        feat = np.ascontiguousarray(feat, np.uint16)
        probs = np.zeros(feat.shape[0] // max(self.stride, 1) + 1, np.float32)
        n = self._L.mwwo_mixednet_predict_u16(self._h, feat.ctypes.data, feat.shape[0], probs.ctypes.data, probs.size)
        return probs[:n]


def run_pipeline(blob, audio: np.ndarray, want_features=True, want_probs=True, clip_loop=False, threads=1):
    """audio int16 [S, N] -> (features uint16 [S, T, 40], probs float32 [S, P]); every stream starts reset."""
    audio = np.ascontiguousarray(audio, np.int16)
    S, N = audio.shape
    rows = N // 160 + 1
    feats = np.zeros((S, rows, NUM_CHANNELS), np.uint16) if want_features else None
    probs = np.zeros((S, rows), np.float32) if (want_probs and blob is not None) else None
    n_rows = np.zeros(S, np.uint64)
    n_probs = np.zeros(S, np.uint64)
    rc = lib().mwwo_pipeline_run(blob, len(blob) if blob else 0, audio.ctypes.data, S, N, _ptr(feats), rows, _ptr(probs), rows,
                                 n_rows.ctypes.data, n_probs.ctypes.data, int(clip_loop), int(threads))
    if rc != 0:
        raise RuntimeError("oracle pipeline failed (%d)" % rc)
    r = int(n_rows[0]) if S else 0
    p = int(n_probs[0]) if S else 0
    return (feats[:, :r] if feats is not None else None), (probs[:, :p] if probs is not None else None)
