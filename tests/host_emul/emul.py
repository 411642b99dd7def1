"""ctypes front-end of tests/host_emul/*.cc -- TEST INFRASTRUCTURE.

Builds (g++) and loads a host executable version of the product's device PHASE functions
(microwakeword_b200/csrc/*_dev.cuh) so that the kernels' index math and bit-exactness can be
checked against the oracle without a GPU.  Nothing in the product imports this."""

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "microwakeword_b200", "csrc")
SO = os.path.join(HERE, "_build", "libemul.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, "emul_frontend.cc"), os.path.join(HERE, "emul_nn.cc"), os.path.join(HERE, "emul_generic.cc"),
                os.path.join(CSRC, "mww_tables.cc")]
        deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
        if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
            os.makedirs(os.path.dirname(SO), exist_ok=True)
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", SO] + srcs, check=True)
        L = ctypes.CDLL(SO)
        L.emul_features.restype = ctypes.c_int
        L.emul_features_fused.restype = ctypes.c_int
        L.emul_features_hop.restype = ctypes.c_int
        L.emul_fb_schedule.restype = None
        L.emul_nn_f32.restype = ctypes.c_int
        L.emul_nn_i8.restype = ctypes.c_int
        L.emul_nn_f32_live.restype = ctypes.c_int
        L.emul_nn_f32_live2.restype = ctypes.c_int
        L.emul_nn_f32_live3.restype = ctypes.c_int
        L.emul_nn_live_canonicalise.restype = None
        L.emul_nn_i8_live.restype = ctypes.c_int
        L.emul_nn_i8_live_canonicalise.restype = None
        L.emul_gen_arch.restype = ctypes.c_int
        L.emul_gen_f32.restype = ctypes.c_int
        L.emul_gen_i8.restype = ctypes.c_int
        L.emul_gen_fill_state_i8.restype = ctypes.c_int
        L.emul_isqrt64_round.restype = ctypes.c_uint32
        L.emul_isqrt64_round.argtypes = [ctypes.c_uint64]
        L.emul_isqrt64_round_fast.restype = ctypes.c_uint32
        L.emul_isqrt64_round_fast.argtypes = [ctypes.c_uint64]
        L.emul_isqrt_fast_mismatch.restype = ctypes.c_longlong
        L.emul_isqrt_fast_mismatch.argtypes = [ctypes.c_void_p, ctypes.c_longlong]
        _lib = L
    return _lib


def _p(a):
    return ctypes.c_void_p(a.ctypes.data)


def tables():
    t = dict(window=np.zeros(480, np.int16), bin_weight=np.zeros(257, np.int16), bin_unweight=np.zeros(257, np.int16),
             chan_start=np.zeros(42, np.int16), gain_lut=np.zeros(125, np.int16), log_lut=np.zeros(129, np.uint16),
             twiddles=np.zeros((256, 2), np.int16), super_twiddles=np.zeros((128, 2), np.int16), info=np.zeros(8, np.int32))
    lib().emul_tables(*[_p(t[k]) for k in ("window", "bin_weight", "bin_unweight", "chan_start", "gain_lut", "log_lut", "twiddles", "super_twiddles", "info")])
    return t


class Frontend:
    """Emulated mww_features state machine for S lockstep streams."""

    def __init__(self, n_streams):
        self.S = n_streams
        self.carry = np.zeros((n_streams, 480), np.int16)
        self.estimate = np.zeros((n_streams, 40), np.uint32)
        self.used = 0

    def features(self, audio, fused=False, order=0, hop=160):
        """fused=True: the phases of the one-launch clip kernel (filterbank -> estimate chain -> outputs from shared memory);
        hop != 160: the run-time window-step kernel"""
        audio = np.ascontiguousarray(audio, np.int16)
        S, N = audio.shape
        rows = (self.used + N) // min(hop, 160) + 1
        if hop != 160:
            feat = np.zeros((S, rows, 40), np.uint16)
            nu = ctypes.c_int(0)
            n = lib().emul_features_hop(_p(audio), S, N, _p(self.carry), self.used, _p(self.estimate), _p(feat), rows, ctypes.byref(nu), int(hop))
            assert n >= 0
            self.used = nu.value
            return feat[:, :n]
        feat = np.zeros((S, rows, 40), np.uint16)
        nu = ctypes.c_int(0)
        if fused:
            n = lib().emul_features_fused(_p(audio), S, N, _p(self.carry), self.used, _p(self.estimate), _p(feat), rows, ctypes.byref(nu), int(order))
        else:
            n = lib().emul_features(_p(audio), S, N, _p(self.carry), self.used, _p(self.estimate), _p(feat), rows, ctypes.byref(nu))
        assert n >= 0
        self.used = nu.value
        return feat[:, :n]


def fb_schedule():
    """(slots [16 lanes][4 slots][ch, word0, n, coef_off], coef int32 [800]) of the mel accumulation"""
    slots = np.zeros((16, 4, 4), np.int16)
    coef = np.zeros(800, np.int32)
    lib().emul_fb_schedule(_p(slots), _p(coef))
    return slots, coef


F32_NAMES = ["first_conv/w"] + ["b%d/dw/w" % i for i in range(4)] + ["b%d/dw/b" % i for i in range(4)] + \
            ["b%d/pw/w" % i for i in range(4)] + ["b%d/pw/b" % i for i in range(4)] + ["head/w", "head/b"]


class NnF32:
    def __init__(self, tensors, n_streams):
        self.arrs = [np.ascontiguousarray(tensors[n], np.float32) for n in F32_NAMES]
        self.wp = (ctypes.c_void_p * len(self.arrs))(*[a.ctypes.data for a in self.arrs])
        self.state = np.zeros((n_streams, 4176), np.float32)
        self.pend = np.zeros((n_streams, 80), np.float32)
        self.n_pend = 0

    def infer(self, rows):
        rows = np.ascontiguousarray(rows)
        S, n_rows = rows.shape[:2]
        probs = np.zeros((S, (self.n_pend + n_rows) // 3 + 1), np.float32)
        n = lib().emul_nn_f32(self.wp, _p(self.state), _p(self.pend), self.n_pend, _p(rows), n_rows, int(rows.dtype == np.float32), S,
                              _p(probs), probs.shape[1], None)
        self.n_pend = (self.n_pend + n_rows) % 3
        return probs[:, :n]


class NnF32Live(NnF32):
    """Same weights / state / pending layout as NnF32, stepped with the live-step kernel's phase functions (3 rows per call).
    Between live steps the rings are ROTATED (heads); `canonicalise` restores the clip kernels' layout, exactly as
    mww_capi.cu does before a clip call or mww_get_state."""

    RING_ROWS = (4, 10, 14, 22, 16)

    def __init__(self, *a, version=2, order=0, **kw):
        """version 1: the r01 kernel (every CTA alternates ring loads and layer chain); 2: the warp-specialised one (register-load
        streamers); 3: the bulk-copy kernel's phase functions (P threads, window warps, v2's chain) -- the GPU default"""
        super().__init__(*a, **kw)
        self.heads = np.zeros(5, np.int32)
        self.version, self.order = version, order

    def step(self, rows3):
        rows3 = np.ascontiguousarray(rows3)
        S = rows3.shape[0]
        assert rows3.shape[1:] == (3, 40)
        probs = np.zeros((S, 1), np.float32)
        if self.version == 3:
            lib().emul_nn_f32_live3(self.wp, _p(self.state), _p(self.pend), self.n_pend, _p(rows3), int(rows3.dtype == np.float32), S, _p(probs), 1,
                                    _p(self.heads), int(self.order))
        elif self.version == 2:
            lib().emul_nn_f32_live2(self.wp, _p(self.state), _p(self.pend), self.n_pend, _p(rows3), int(rows3.dtype == np.float32), S, _p(probs), 1,
                                    _p(self.heads), int(self.order))
        else:
            lib().emul_nn_f32_live(self.wp, _p(self.state), _p(self.pend), self.n_pend, _p(rows3), int(rows3.dtype == np.float32), S, _p(probs), 1,
                                   _p(self.heads))
        self.heads = ((self.heads + 1) % np.asarray(self.RING_ROWS, np.int32)).astype(np.int32)
        return probs[:, 0]

    def canonicalise(self):
        lib().emul_nn_live_canonicalise(_p(self.state), self.state.shape[0], _p(self.heads))
        self.heads[:] = 0

    def infer(self, rows):
        self.canonicalise()
        return super().infer(rows)


def i8_names():
    names = ["q/first_conv/w", "q/first_conv/bias", "q/first_conv/mult", "q/first_conv/shift"]
    for i in range(4):
        names += ["q/b%d/dw/%s" % (i, x) for x in ("w", "bias", "mult", "shift")] + ["q/b%d/pw/%s" % (i, x) for x in ("w", "bias", "mult", "shift")]
    return names + ["q/head/w", "q/logistic_lut"]


class NnI8:
    def __init__(self, q, n_streams):
        self.arrs = [np.ascontiguousarray(q[n]) for n in i8_names()]
        self.wp = (ctypes.c_void_p * len(self.arrs))(*[a.ctypes.data for a in self.arrs])
        self.zp = np.ascontiguousarray(q["q/zps"], np.int32)
        self.head3 = np.array([q["q/head/bias"][0], q["q/head/mult"][0], q["q/head/shift"][0]], np.int32)
        self.in_scale = float(q["q/scales"][0])
        self.state = np.zeros((n_streams, 4176), np.int8)
        self.pend = np.zeros((n_streams, 80), np.int8)
        self.n_pend = 0
        self.reset()

    def reset(self):
        lib().emul_fill_state_i8(_p(self.zp), _p(self.state), _p(self.pend), self.state.shape[0])
        self.n_pend = 0

    def infer(self, rows):
        rows = np.ascontiguousarray(rows)
        S, n_rows = rows.shape[:2]
        rt = {np.dtype(np.uint16): 0, np.dtype(np.float32): 1, np.dtype(np.int8): 2}[rows.dtype]
        probs = np.zeros((S, (self.n_pend + n_rows) // 3 + 1), np.float32)
        n = lib().emul_nn_i8(self.wp, _p(self.zp), _p(self.head3), ctypes.c_float(self.in_scale), _p(self.state), _p(self.pend), self.n_pend,
                             _p(rows), n_rows, rt, S, _p(probs), probs.shape[1])
        self.n_pend = (self.n_pend + n_rows) % 3
        return probs[:, :n]


class NnI8Live(NnI8):
    """NnI8 stepped with the int8 live-step kernel's phase functions (3 rows per call); rings stay rotated between
    live steps and are canonicalised before a clip-kernel call, as mww_capi.cu does."""

    RING_ROWS = (4, 10, 14, 22, 16)

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.heads = np.zeros(5, np.int32)

    def step(self, rows3):
        rows3 = np.ascontiguousarray(rows3)
        S = rows3.shape[0]
        assert rows3.shape[1:] == (3, 40)
        rt = {np.dtype(np.uint16): 0, np.dtype(np.float32): 1, np.dtype(np.int8): 2}[rows3.dtype]
        probs = np.zeros((S, 1), np.float32)
        lib().emul_nn_i8_live(self.wp, _p(self.zp), _p(self.head3), ctypes.c_float(self.in_scale), _p(self.state), _p(self.pend), self.n_pend,
                              _p(rows3), rt, S, _p(probs), 1, _p(self.heads))
        self.heads = ((self.heads + 1) % np.asarray(self.RING_ROWS, np.int32)).astype(np.int32)
        return probs[:, 0]

    def canonicalise(self):
        lib().emul_nn_i8_live_canonicalise(_p(self.state), self.state.shape[0], _p(self.heads))
        self.heads[:] = 0

    def infer(self, rows):
        self.canonicalise()
        return super().infer(rows)


# ---- run-time-geometry MixedNet (mww_nn_generic.cuh) --------------------------------------------------------------

def gen_arch_info(arch):
    arch = np.ascontiguousarray(arch, np.int32)
    out = np.zeros(8, np.int64)
    rc = lib().emul_gen_arch(_p(arch), arch.size, _p(out))
    keys = ("rc", "state_elems", "pend_cap", "stride", "sm_elems", "macs_per_step", "ring0", "c_last")
    return dict(zip(keys, (int(v) for v in out))) if rc == 0 else {"rc": rc}


class GenF32:
    """Any-architecture fp32 MixedNet stepped with the generic kernel's phase functions; `tensors` = container tensors."""

    def __init__(self, tensors, n_streams):
        self.arch = np.ascontiguousarray(tensors["arch"], np.int32)
        info = gen_arch_info(self.arch)
        assert info["rc"] == 0, info
        self.info = info
        nb = int(self.arch[4])
        names = ["first_conv/w"]
        for i in range(nb):
            names += ["b%d/dw/w" % i, "b%d/dw/b" % i, "b%d/pw/w" % i, "b%d/pw/b" % i]
        names += ["head/w", "head/b"]
        self.arrs = [np.ascontiguousarray(tensors[n], np.float32) for n in names]
        self.wp = (ctypes.c_void_p * len(self.arrs))(*[a.ctypes.data for a in self.arrs])
        self.state = np.zeros((n_streams, info["state_elems"]), np.float32)
        self.pend = np.zeros((n_streams, info["pend_cap"] * 40), np.float32)
        self.n_pend = 0
        self.stride = info["stride"]

    def infer(self, rows):
        rows = np.ascontiguousarray(rows)
        S, n_rows = rows.shape[:2]
        rt = {np.dtype(np.uint16): 0, np.dtype(np.float32): 1}[rows.dtype]
        probs = np.zeros((S, (self.n_pend + n_rows) // self.stride + 1), np.float32)
        n = lib().emul_gen_f32(_p(self.arch), self.arch.size, self.wp, _p(self.state), _p(self.pend), self.n_pend, _p(rows), n_rows, rt, S,
                               _p(probs), probs.shape[1])
        assert n >= 0, n
        self.n_pend = (self.n_pend + n_rows) % self.stride
        return probs[:, :n]


class GenI8:
    def __init__(self, q, n_streams):
        self.arch = np.ascontiguousarray(q["arch"], np.int32)
        info = gen_arch_info(self.arch)
        assert info["rc"] == 0, info
        self.info = info
        nb = int(self.arch[4])
        names = ["q/first_conv/w", "q/first_conv/bias", "q/first_conv/mult", "q/first_conv/shift"]
        for i in range(nb):
            names += ["q/b%d/dw/%s" % (i, x) for x in ("w", "bias", "mult", "shift")] + ["q/b%d/pw/%s" % (i, x) for x in ("w", "bias", "mult", "shift")]
        names += ["q/head/w", "q/logistic_lut"]
        self.arrs = [np.ascontiguousarray(q[n]) for n in names]
        self.wp = (ctypes.c_void_p * len(self.arrs))(*[a.ctypes.data for a in self.arrs])
        self.zp = np.ascontiguousarray(q["q/zps"], np.int32)
        assert self.zp.size == 4 + 2 * nb
        self.head3 = np.array([q["q/head/bias"][0], q["q/head/mult"][0], q["q/head/shift"][0]], np.int32)
        self.in_scale = float(q["q/scales"][0])
        self.state = np.zeros((n_streams, info["state_elems"]), np.int8)
        self.pend = np.zeros((n_streams, info["pend_cap"] * 40), np.int8)
        self.stride = info["stride"]
        self.reset()

    def reset(self):
        assert lib().emul_gen_fill_state_i8(_p(self.arch), self.arch.size, _p(self.zp), _p(self.state), _p(self.pend), self.state.shape[0]) == 0
        self.n_pend = 0

    def infer(self, rows):
        rows = np.ascontiguousarray(rows)
        S, n_rows = rows.shape[:2]
        rt = {np.dtype(np.uint16): 0, np.dtype(np.float32): 1, np.dtype(np.int8): 2}[rows.dtype]
        probs = np.zeros((S, (self.n_pend + n_rows) // self.stride + 1), np.float32)
        n = lib().emul_gen_i8(_p(self.arch), self.arch.size, self.wp, _p(self.zp), _p(self.head3), ctypes.c_float(self.in_scale), _p(self.state),
                              _p(self.pend), self.n_pend, _p(rows), n_rows, rt, S, _p(probs), probs.shape[1])
        assert n >= 0, n
        self.n_pend = (self.n_pend + n_rows) % self.stride
        return probs[:, :n]
