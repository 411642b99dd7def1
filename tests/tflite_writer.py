"""Test infrastructure: writes a TFL3 flatbuffer of the internal-state streaming MixedNet, shaped like
the file ``utils.convert_saved_model_to_tflite`` (utils.py:289-348) produces from ``mixednet.model``
(mixednet.py:278-386) -- ring-buffer variables (VAR_HANDLE / READ_VARIABLE / ASSIGN_VARIABLE with a CALL_ONCE
initialiser subgraph), CONCATENATION + STRIDED_SLICE ring updates (stream.py:586-590), MixConv as SPLIT_V +
StridedKeep STRIDED_SLICE + DEPTHWISE_CONV_2D groups + CONCATENATION (mixednet.py:186-231), 1x1 CONV_2D with the
BatchNorm folded and a fused RELU, RESHAPE + FULLY_CONNECTED + LOGISTIC, and for the int8 graph a final QUANTIZE
to uint8 (utils.py:337-338).

No TensorFlow exists here, so these files are NOT TensorFlow output: they exercise the flatbuffer reader and the
graph recogniser of ``microwakeword_b200/tflite_file.py`` and are executed op by op by
``oracle/tflite_interp.py``.  Field numbers / operator codes restate tensorflow/lite/schema/schema.fbs (v3).
"""

from __future__ import annotations

import struct

import numpy as np


class Builder:
    """Minimal flatbuffer builder (grows downwards like the official one; offsets are measured from the end)."""

    def __init__(self):
        self.buf = bytearray()
        self.minalign = 1
        self._fields = None
        self._obj_end = 0

    def _prep(self, size, additional):
        self.minalign = max(self.minalign, size)
        pad = (-(len(self.buf) + additional)) % size
        if pad:
            self.buf[0:0] = bytes(pad)

    def _put(self, fmt, *v):
        self.buf[0:0] = struct.pack("<" + fmt, *v)

    def string(self, s: str) -> int:
        b = s.encode()
        self._prep(4, len(b) + 1)
        self.buf[0:0] = b + b"\0"
        self._put("I", len(b))
        return len(self.buf)

    def vector(self, fmt: str, values, align=None) -> int:
        size = struct.calcsize(fmt)
        values = list(values)
        self._prep(4, len(values) * size)
        self._prep(align or size, len(values) * size)
        if values:
            self.buf[0:0] = struct.pack("<%d%s" % (len(values), fmt), *values)
        self._put("I", len(values))
        return len(self.buf)

    def bytes_vector(self, raw: bytes, align=16) -> int:
        self._prep(4, len(raw))
        self._prep(align, len(raw))
        self.buf[0:0] = raw
        self._put("I", len(raw))
        return len(self.buf)

    def offsets_vector(self, offs) -> int:
        offs = list(offs)
        self._prep(4, 4 * len(offs))
        for o in reversed(offs):
            self._put("I", len(self.buf) + 4 - o)
        self._put("I", len(offs))
        return len(self.buf)

    def start(self):
        self._fields = {}
        self._obj_end = len(self.buf)

    def scalar(self, idx, fmt, v, default=0):
        if v == default:
            return
        self._prep(struct.calcsize(fmt), 0)
        self._put(fmt, v)
        self._fields[idx] = len(self.buf)

    def offset(self, idx, target):
        if not target:
            return
        self._prep(4, 0)
        self._put("I", len(self.buf) + 4 - target)
        self._fields[idx] = len(self.buf)

    def end(self) -> int:
        self._prep(4, 0)
        self._put("i", 0)
        table = len(self.buf)
        n = (max(self._fields) + 1) if self._fields else 0
        vt = [4 + 2 * n, table - self._obj_end] + [(table - self._fields[i]) if i in self._fields else 0 for i in range(n)]
        if len(vt) % 2:
            self.buf[0:0] = b"\0\0"                      # keep 4-byte alignment below the vtable
        self.buf[0:0] = struct.pack("<%dH" % len(vt), *vt)
        vtab = len(self.buf)
        struct.pack_into("<i", self.buf, len(self.buf) - table, vtab - table)
        self._fields = None
        return table

    def finish(self, root, ident: bytes) -> bytes:
        self._prep(self.minalign, 8)
        self.buf[0:0] = ident
        self._put("I", len(self.buf) + 4 - root)
        return bytes(self.buf)


# schema constants (schema.fbs)
T_FLOAT32, T_INT32, T_UINT8, T_INT8, T_RESOURCE = 0, 2, 3, 9, 13
_NP2T = {np.dtype(np.float32): T_FLOAT32, np.dtype(np.int32): T_INT32, np.dtype(np.uint8): T_UINT8, np.dtype(np.int8): T_INT8}
OPS = {"CONCATENATION": 2, "CONV_2D": 3, "DEPTHWISE_CONV_2D": 4, "FULLY_CONNECTED": 9, "LOGISTIC": 14, "RESHAPE": 22,
       "STRIDED_SLICE": 45, "SPLIT_V": 102, "QUANTIZE": 114, "CALL_ONCE": 129, "VAR_HANDLE": 142, "READ_VARIABLE": 143,
       "ASSIGN_VARIABLE": 144}
# BuiltinOptions union member numbers
OPT = {"CONV_2D": 1, "DEPTHWISE_CONV_2D": 2, "FULLY_CONNECTED": 8, "CONCATENATION": 10, "RESHAPE": 17, "STRIDED_SLICE": 32,
       "LOGISTIC": 0, "SPLIT_V": 79, "QUANTIZE": 85, "CALL_ONCE": 103, "VAR_HANDLE": 111, "READ_VARIABLE": 112, "ASSIGN_VARIABLE": 113}


class GraphWriter:
    def __init__(self):
        self.tensors = []       # dicts
        self.buffers = [b""]    # buffer 0 is the empty sentinel
        self.ops = []           # (name, inputs, outputs, options dict)
        self.init_ops = []      # initialiser subgraph: (var name, const tensor)

    def tensor(self, name, shape, dtype, data=None, scale=None, zero_point=None, qdim=0, ttype=None):
        buf = 0
        if data is not None:
            self.buffers.append(np.ascontiguousarray(np.asarray(data, dtype)).tobytes())
            buf = len(self.buffers) - 1
        self.tensors.append(dict(name=name, shape=tuple(int(s) for s in shape), type=_NP2T[np.dtype(dtype)] if ttype is None else ttype,
                                 buffer=buf, scale=scale, zero_point=zero_point, qdim=qdim))
        return len(self.tensors) - 1

    def op(self, name, inputs, outputs, **options):
        self.ops.append((name, list(inputs), list(outputs), options))

    # ---- serialisation ----
    def _options(self, b: Builder, name, o):
        b.start()
        if name == "CONV_2D":
            b.scalar(0, "b", 1)                                    # padding VALID
            b.scalar(1, "i", o.get("stride_w", 1)); b.scalar(2, "i", o.get("stride_h", 1))
            b.scalar(3, "b", o.get("act", 0)); b.scalar(4, "i", 1, default=None); b.scalar(5, "i", 1, default=None)
        elif name == "DEPTHWISE_CONV_2D":
            b.scalar(0, "b", 1); b.scalar(1, "i", 1); b.scalar(2, "i", 1); b.scalar(3, "i", 1); b.scalar(4, "b", o.get("act", 0))
            b.scalar(5, "i", 1, default=None); b.scalar(6, "i", 1, default=None)
        elif name == "FULLY_CONNECTED":
            b.scalar(0, "b", 0)
        elif name == "CONCATENATION":
            b.scalar(0, "i", o["axis"])
        elif name == "STRIDED_SLICE":
            for i, k in enumerate(("begin_mask", "end_mask", "ellipsis_mask", "new_axis_mask", "shrink_axis_mask")):
                b.scalar(i, "i", o.get(k, 0))
        elif name == "SPLIT_V":
            b.scalar(0, "i", o["num_splits"])
        elif name == "CALL_ONCE":
            b.scalar(0, "i", o["init_subgraph_index"])
        elif name == "VAR_HANDLE":
            sn = b.string(o["shared_name"])
            b.start()
            b.offset(1, sn)
        return b.end()

    def _subgraph(self, b: Builder, tensors, ops, inputs, outputs, name):
        t_offs = []
        for t in tensors:
            q = 0
            if t["scale"] is not None:
                sc = b.vector("f", [float(x) for x in np.atleast_1d(t["scale"])])
                zp = b.vector("q", [int(x) for x in np.atleast_1d(t["zero_point"])], align=8)
                b.start(); b.offset(2, sc); b.offset(3, zp); b.scalar(6, "i", t["qdim"])
                q = b.end()
            nm = b.string(t["name"])
            sh = b.vector("i", t["shape"])
            b.start()
            b.offset(0, sh); b.scalar(1, "b", t["type"]); b.scalar(2, "I", t["buffer"]); b.offset(3, nm); b.offset(4, q)
            t_offs.append(b.end())
        o_offs = []
        for name_, ins, outs, o in ops:
            opt = self._options(b, name_, o) if name_ in ("CONV_2D", "DEPTHWISE_CONV_2D", "FULLY_CONNECTED", "CONCATENATION", "STRIDED_SLICE",
                                                           "SPLIT_V", "CALL_ONCE", "VAR_HANDLE") else 0
            iv, ov = b.vector("i", ins), b.vector("i", outs)
            b.start()
            b.scalar(0, "I", self._code_index[name_]); b.offset(1, iv); b.offset(2, ov)
            if opt:
                b.scalar(3, "B", OPT[name_]); b.offset(4, opt)
            o_offs.append(b.end())
        tv, ov_ = b.offsets_vector(t_offs), b.offsets_vector(o_offs)
        iv, outv, nm = b.vector("i", inputs), b.vector("i", outputs), b.string(name)
        b.start()
        b.offset(0, tv); b.offset(1, iv); b.offset(2, outv); b.offset(3, ov_); b.offset(4, nm)
        return b.end()

    def serialise(self, inputs, outputs, init=None) -> bytes:
        """init: optional (tensors, ops) of the CALL_ONCE initialiser subgraph"""
        b = Builder()
        used = []
        for name_, *_ in self.ops + (init[1] if init else []):
            if name_ not in used:
                used.append(name_)
        self._code_index = {n: i for i, n in enumerate(used)}
        buf_offs = []
        for raw in self.buffers:
            d = b.bytes_vector(raw) if raw else 0
            b.start(); b.offset(0, d)
            buf_offs.append(b.end())
        code_offs = []
        for n in used:
            c = OPS[n]
            b.start()
            b.scalar(0, "b", min(c, 127)); b.scalar(2, "i", 1, default=None); b.scalar(3, "i", c)
            code_offs.append(b.end())
        sgs = [self._subgraph(b, self.tensors, self.ops, inputs, outputs, "main")]
        if init:
            sgs.append(self._subgraph(b, init[0], init[1], [], [], "NoOp"))
        desc = b.string("written by tests/tflite_writer.py (not TensorFlow output)")
        cv, sv, bv = b.offsets_vector(code_offs), b.offsets_vector(sgs), b.offsets_vector(buf_offs)
        b.start()
        b.scalar(0, "I", 3); b.offset(1, cv); b.offset(2, sv); b.offset(3, desc); b.offset(4, bv)
        return b.finish(b.end(), b"TFL3")


def _sweights(w):
    return dict(scale=None, zero_point=None)


def write_streaming_mixednet(tensors: dict, shuffle_seed=None) -> bytes:
    """``model_file`` tensor dictionary (fp32 or int8) -> TFL3 bytes of the equivalent streaming graph.
    shuffle_seed: renumber the main subgraph's tensors and the model's buffers at random (the converter's numbering is not
    this writer's; a reader must follow the dataflow, not the indices)."""
    from microwakeword_b200 import model_file as MF
    arch = MF.Arch.decode(tensors["arch"])
    quant = MF.is_quantized(tensors)
    gw = GraphWriter()
    adt = np.int8 if quant else np.float32
    names = MF.act_names(arch)
    aq = {n: (float(tensors["q/scales"][i]), int(tensors["q/zps"][i])) for i, n in enumerate(names)} if quant else {}

    def act(name, shape, qname):
        if quant:
            return gw.tensor(name, shape, adt, scale=[aq[qname][0]], zero_point=[aq[qname][1]])
        return gw.tensor(name, shape, adt)

    def const_i32(name, vals):
        return gw.tensor(name, (len(vals),), np.int32, data=np.asarray(vals, np.int32))

    def wscales(prefix, s_in, s_out, n):
        """per-channel float32 weight scales recovered from the container's fixed-point multipliers"""
        m = tensors["q/" + prefix + "/mult"].astype(np.float64)
        sh = tensors["q/" + prefix + "/shift"].astype(np.float64)
        real = m / 2.0 ** 31 * 2.0 ** sh
        return (real * np.float64(s_out) / np.float64(s_in)).astype(np.float32)

    init_tensors, init_ops = [], []
    n_var = [0]

    def ring(prefix, x, rows, ch, qname):
        """Stream ring buffer: returns the memory tensor [1, rows + new, 1, ch]"""
        new_rows = gw.tensors[x]["shape"][1]
        var = gw.tensor(prefix + "/state", (), np.float32, ttype=13)
        gw.op("VAR_HANDLE", [], [var], shared_name=prefix + "/state")
        st = act(prefix + "/state_read", (1, rows, 1, ch), qname)
        gw.op("READ_VARIABLE", [var], [st])
        mem = act(prefix + "/memory", (1, rows + new_rows, 1, ch), qname)
        gw.op("CONCATENATION", [st, x], [mem], axis=1)
        keep = act(prefix + "/state_next", (1, rows, 1, ch), qname)
        gw.op("STRIDED_SLICE", [mem, const_i32(prefix + "/ss_begin", [0, new_rows, 0, 0]), const_i32(prefix + "/ss_end", [1, rows + new_rows, 1, ch]),
                                const_i32(prefix + "/ss_strides", [1, 1, 1, 1])], [keep])
        gw.op("ASSIGN_VARIABLE", [var, keep], [])
        # initialiser subgraph entry: real zero
        iv = len(init_tensors)
        init_tensors.append(dict(name=prefix + "/state", shape=(), type=13, buffer=0, scale=None, zero_point=None, qdim=0))
        zero = np.full((1, rows, 1, ch), aq[qname][1] if quant else 0, adt)
        gw.buffers.append(zero.tobytes())
        init_tensors.append(dict(name=prefix + "/zeros", shape=zero.shape, type=9 if quant else 0, buffer=len(gw.buffers) - 1,
                                 scale=[aq[qname][0]] if quant else None, zero_point=[aq[qname][1]] if quant else None, qdim=0))
        init_ops.append(("VAR_HANDLE", [], [iv], dict(shared_name=prefix + "/state")))
        init_ops.append(("ASSIGN_VARIABLE", [iv, iv + 1], [], {}))
        n_var[0] += 1
        return mem

    gw.op("CALL_ONCE", [], [], init_subgraph_index=1)
    k0, f0, stride = arch.first_conv_kernel_size, arch.first_conv_filters, arch.stride
    t_in = act("serving_default_input_audio:0", (1, stride, 40), "in")
    x = act("expand_dims", (1, stride, 1, 40), "in")
    gw.op("RESHAPE", [t_in, const_i32("expand_dims/shape", [1, stride, 1, 40])], [x])
    mem = ring("stream", x, arch.first_conv_ring_rows, 40, "in")
    if quant:
        w = np.transpose(tensors["q/first_conv/w"], (2, 0, 1))[:, :, None, :]              # [F, K, 1, 40]
        wt = gw.tensor("stream/conv2d/kernel", w.shape, np.int8, data=w, scale=wscales("first_conv", aq["in"][0], aq["c0"][0], f0),
                       zero_point=np.zeros(f0, np.int64), qdim=0)
        bt = gw.tensor("stream/conv2d/bias", (f0,), np.int32, data=tensors["q/first_conv/bias"],
                       scale=wscales("first_conv", aq["in"][0], aq["c0"][0], f0) * np.float32(aq["in"][0]), zero_point=np.zeros(f0, np.int64))
    else:
        w = np.transpose(tensors["first_conv/w"], (2, 0, 1))[:, :, None, :]
        wt = gw.tensor("stream/conv2d/kernel", w.shape, np.float32, data=w)
        bt = gw.tensor("stream/conv2d/bias", (f0,), np.float32, data=np.zeros(f0, np.float32))
    cur = act("activation/Relu", (1, 1, 1, f0), "c0")
    gw.op("CONV_2D", [mem, wt, bt], [cur], stride_h=stride, act=1)
    cin = f0
    prev_q = "c0"
    for b in range(arch.n_blocks):
        ks = arch.mixconv_kernel_sizes[b]
        kmax = max(ks)
        dq, pq = "d%d" % (b + 1), "p%d" % (b + 1)
        if kmax == 1:
            # mixednet.py:346-348: no MixConv layer at all -- the 1x1 conv reads the previous activation tensor
            cout = arch.pointwise_filters[b]
            d = cur
            if quant:
                wq = np.transpose(tensors["q/b%d/pw/w" % b])[:, None, None, :]
                sw = wscales("b%d/pw" % b, aq[prev_q][0], aq[pq][0], cout)
                wt = gw.tensor("pw_%d/kernel" % b, wq.shape, np.int8, data=wq, scale=sw, zero_point=np.zeros(cout, np.int64), qdim=0)
                bt = gw.tensor("pw_%d/bias" % b, (cout,), np.int32, data=tensors["q/b%d/pw/bias" % b], scale=sw * np.float32(aq[prev_q][0]),
                               zero_point=np.zeros(cout, np.int64))
            else:
                wf = np.transpose(tensors["b%d/pw/w" % b])[:, None, None, :]
                wt = gw.tensor("pw_%d/kernel" % b, wf.shape, np.float32, data=wf)
                bt = gw.tensor("pw_%d/bias" % b, (cout,), np.float32, data=tensors["b%d/pw/b" % b])
            cur = act("pw_%d/Relu" % b, (1, 1, 1, cout), pq)
            gw.op("CONV_2D", [d, wt, bt], [cur], act=1)
            cin, prev_q = cout, pq
            continue
        mem = ring("stream_%d" % (b + 1), cur, kmax - 1, cin, prev_q)
        split = [cin // len(ks)] * len(ks)
        split[0] += cin - sum(split)
        if len(ks) > 1:
            parts = [act("mixconv_%d/split:%d" % (b, j), (1, kmax, 1, c), prev_q) for j, c in enumerate(split)]
            gw.op("SPLIT_V", [mem, const_i32("mixconv_%d/size_splits" % b, split), gw.tensor("mixconv_%d/axis" % b, (), np.int32, data=np.asarray(3, np.int32))],
                  parts, num_splits=len(ks))
        else:
            parts = [mem]
        outs, c_off = [], 0
        for j, (k, c) in enumerate(zip(ks, split)):
            src = parts[j]
            if k != kmax:
                kept = act("mixconv_%d/keep_%d" % (b, j), (1, k, 1, c), prev_q)
                gw.op("STRIDED_SLICE", [src, const_i32("mixconv_%d/keep_%d/begin" % (b, j), [0, kmax - k, 0, 0]),
                                        const_i32("mixconv_%d/keep_%d/end" % (b, j), [1, kmax, 1, c]),
                                        const_i32("mixconv_%d/keep_%d/strides" % (b, j), [1, 1, 1, 1])], [kept])
                src = kept
            if quant:
                wq = tensors["q/b%d/dw/w" % b][kmax - k:, c_off:c_off + c][None, :, None, :]
                sw = wscales("b%d/dw" % b, aq[prev_q][0], aq[dq][0], cin)[c_off:c_off + c]
                wt = gw.tensor("dw_%d_%d/kernel" % (b, j), wq.shape, np.int8, data=wq, scale=sw, zero_point=np.zeros(c, np.int64), qdim=3)
                bt = gw.tensor("dw_%d_%d/bias" % (b, j), (c,), np.int32, data=tensors["q/b%d/dw/bias" % b][c_off:c_off + c],
                               scale=sw * np.float32(aq[prev_q][0]), zero_point=np.zeros(c, np.int64))
            else:
                wf = tensors["b%d/dw/w" % b][kmax - k:, c_off:c_off + c][None, :, None, :]
                wt = gw.tensor("dw_%d_%d/kernel" % (b, j), wf.shape, np.float32, data=wf)
                bt = gw.tensor("dw_%d_%d/bias" % (b, j), (c,), np.float32, data=tensors["b%d/dw/b" % b][c_off:c_off + c])
            o = act("dw_%d_%d/out" % (b, j), (1, 1, 1, c), dq)
            gw.op("DEPTHWISE_CONV_2D", [src, wt, bt], [o])
            outs.append(o)
            c_off += c
        if len(outs) > 1:
            d = act("mixconv_%d/concat" % b, (1, 1, 1, cin), dq)
            gw.op("CONCATENATION", outs, [d], axis=3)
        else:
            d = outs[0]
        cout = arch.pointwise_filters[b]
        if quant:
            wq = tensors["q/b%d/pw/w" % b].T[:, None, None, :]
            sw = wscales("b%d/pw" % b, aq[dq][0], aq[pq][0], cout)
            wt = gw.tensor("pw_%d/kernel" % b, wq.shape, np.int8, data=wq, scale=sw, zero_point=np.zeros(cout, np.int64), qdim=0)
            bt = gw.tensor("pw_%d/bias" % b, (cout,), np.int32, data=tensors["q/b%d/pw/bias" % b], scale=sw * np.float32(aq[dq][0]),
                           zero_point=np.zeros(cout, np.int64))
        else:
            wf = tensors["b%d/pw/w" % b].T[:, None, None, :]
            wt = gw.tensor("pw_%d/kernel" % b, wf.shape, np.float32, data=wf)
            bt = gw.tensor("pw_%d/bias" % b, (cout,), np.float32, data=tensors["b%d/pw/b" % b])
        cur = act("activation_%d/Relu" % (b + 1), (1, 1, 1, cout), pq)
        gw.op("CONV_2D", [d, wt, bt], [cur], act=1)
        cin, prev_q = cout, pq
    t_head = arch.head_rows
    mem = ring("stream_head", cur, t_head - 1, cin, prev_q)
    flat = act("flatten/Reshape", (1, t_head * cin), prev_q)
    gw.op("RESHAPE", [mem, const_i32("flatten/shape", [1, t_head * cin])], [flat])
    if quant:
        hw = tensors["q/head/w"].reshape(1, -1)
        sw = wscales("head", aq[prev_q][0], aq["fc"][0], 1)
        wt = gw.tensor("dense/kernel", hw.shape, np.int8, data=hw, scale=sw, zero_point=[0])
        bt = gw.tensor("dense/bias", (1,), np.int32, data=tensors["q/head/bias"], scale=sw * np.float32(aq[prev_q][0]), zero_point=[0])
    else:
        wt = gw.tensor("dense/kernel", (1, t_head * cin), np.float32, data=tensors["head/w"].reshape(1, -1))
        bt = gw.tensor("dense/bias", (1,), np.float32, data=tensors["head/b"])
    logit = act("dense/BiasAdd", (1, 1), "fc")
    gw.op("FULLY_CONNECTED", [flat, wt, bt], [logit])
    prob = act("dense/Sigmoid", (1, 1), "prob")
    gw.op("LOGISTIC", [logit], [prob])
    out = prob
    if quant:
        out = gw.tensor("StatefulPartitionedCall:0", (1, 1), np.uint8, scale=[1.0 / 256.0], zero_point=[0])
        gw.op("QUANTIZE", [prob], [out])
    if shuffle_seed is not None:
        rng = np.random.default_rng(shuffle_seed)
        n = len(gw.tensors)
        perm = rng.permutation(n)                         # old index i -> new index perm[i]
        new_tensors = [None] * n
        for i, t in enumerate(gw.tensors):
            new_tensors[perm[i]] = t
        gw.tensors = new_tensors
        gw.ops = [(name, [int(perm[i]) for i in ins], [int(perm[i]) for i in outs], o) for name, ins, outs, o in gw.ops]
        t_in, out = int(perm[t_in]), int(perm[out])
        nb = len(gw.buffers)
        bperm = np.concatenate([[0], 1 + rng.permutation(nb - 1)]) if nb > 1 else np.arange(nb)      # buffer 0 stays the empty one
        new_buffers = [None] * nb
        for i, raw in enumerate(gw.buffers):
            new_buffers[bperm[i]] = raw
        gw.buffers = new_buffers
        for t in gw.tensors:
            t["buffer"] = int(bperm[t["buffer"]])
        for t in init_tensors:
            t["buffer"] = int(bperm[t["buffer"]])
    return gw.serialise([t_in], [out], init=(init_tensors, init_ops))
