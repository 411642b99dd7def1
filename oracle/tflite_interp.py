"""TEST INFRASTRUCTURE (never imported by the product): op-by-op executor for the TFL3 streaming graphs of
SURVEY.md section 8 f-2.  It runs the graph exactly as written in the file -- variables, CONCATENATION,
STRIDED_SLICE, SPLIT_V, CONV_2D, DEPTHWISE_CONV_2D, FULLY_CONNECTED, LOGISTIC, QUANTIZE -- with the TFLite
reference-kernel integer semantics of SURVEY.md Appendix C, so that the *recogniser* in
``microwakeword_b200/tflite_file.py`` (which never executes the graph) can be checked against an independent
reading of the same bytes: executing the file must equal running ``oracle/mixednet_ref`` on the recognised tensors.

It borrows only the flatbuffer wire decoding (``tflite_file.Graph``); operator semantics are restated here.
PARITY UNPINNED: no TensorFlow-written model or tf.lite.Interpreter exists in this environment.
"""

from __future__ import annotations

import numpy as np

from microwakeword_b200 import tflite_file as TF

from . import mixednet_ref as R


class Interpreter:
    def __init__(self, blob: bytes):
        self.g = TF.Graph(blob)
        self.vars = {}
        self.reset()

    # the CALL_ONCE initialiser (subgraph 1) assigns every ring its start value
    def reset(self):
        self.vars = {}
        if len(self.g.subgraphs) > 1:
            tensors, ops, _, _ = self.g.subgraphs[1]
            handle = {}
            for op in ops:
                if op.code == TF.OP_VAR_HANDLE:
                    handle[op.outputs[0]] = op.options.string(1)
                elif op.code == TF.OP_ASSIGN_VARIABLE:
                    self.vars[handle[op.inputs[0]]] = tensors[op.inputs[1]].data().copy()

    @staticmethod
    def _requant(acc, s_in, w_t, out_t, n_out, relu):
        sw = w_t.scale.astype(np.float32)
        sw = np.repeat(sw, n_out) if sw.size == 1 else sw
        s_out, zp_out = out_t.per_tensor_q()
        ms = [TF.quantize_multiplier(np.float64(s_in) * np.float64(s) / np.float64(s_out)) for s in sw]
        y = R.mbqm(acc, np.asarray([m for m, _ in ms], np.int64), np.asarray([s for _, s in ms], np.int64)) + zp_out
        return np.clip(y, zp_out if relu else -128, 127).astype(np.int8)

    def invoke(self, x: np.ndarray) -> np.ndarray:
        g = self.g
        val = {}
        handle = {}
        tin = g.tensors[g.inputs[0]]
        val[g.inputs[0]] = np.asarray(x, tin.dtype).reshape(tin.shape)

        def get(i):
            if i in val:
                return val[i]
            return g.tensors[i].data()

        for op in g.ops:
            c = op.code
            if c == TF.OP_CALL_ONCE:
                continue
            if c == TF.OP_VAR_HANDLE:
                handle[op.outputs[0]] = op.options.string(1)
            elif c == TF.OP_READ_VARIABLE:
                val[op.outputs[0]] = self.vars[handle[op.inputs[0]]].copy()
            elif c == TF.OP_ASSIGN_VARIABLE:
                self.vars[handle[op.inputs[0]]] = get(op.inputs[1]).copy()
            elif c in (TF.OP_RESHAPE, TF.OP_SQUEEZE, TF.OP_EXPAND_DIMS):
                val[op.outputs[0]] = get(op.inputs[0]).reshape(g.tensors[op.outputs[0]].shape)
            elif c == TF.OP_CONCATENATION:
                val[op.outputs[0]] = np.concatenate([get(i) for i in op.inputs], axis=op.opt(0, "i"))
            elif c == TF.OP_STRIDED_SLICE:
                a = get(op.inputs[0])
                begin, end, strides = (get(i).reshape(-1) for i in op.inputs[1:4])
                bm, em = op.opt(0, "i"), op.opt(1, "i")
                if op.opt(2, "i") or op.opt(3, "i") or op.opt(4, "i"):
                    raise NotImplementedError("STRIDED_SLICE ellipsis / new-axis / shrink masks")
                sl = tuple(slice(None if bm >> d & 1 else int(begin[d]), None if em >> d & 1 else int(end[d]), int(strides[d])) for d in range(a.ndim))
                val[op.outputs[0]] = a[sl]
            elif c == TF.OP_SPLIT_V:
                a = get(op.inputs[0])
                sizes = [int(v) for v in get(op.inputs[1]).reshape(-1)]
                axis = int(get(op.inputs[2]).reshape(-1)[0])
                for o, part in zip(op.outputs, np.split(a, np.cumsum(sizes)[:-1], axis=axis)):
                    val[o] = part
            elif c in (TF.OP_CONV_2D, TF.OP_DEPTHWISE_CONV_2D):
                a, w_t = get(op.inputs[0]), g.tensors[op.inputs[1]]
                w = w_t.data()
                bias = get(op.inputs[2]).reshape(-1) if len(op.inputs) > 2 and op.inputs[2] >= 0 else 0
                out_t = g.tensors[op.outputs[0]]
                sh = op.opt(2, "i", 1)
                relu = op.opt(3 if c == TF.OP_CONV_2D else 4, "b") == TF.ACT_RELU
                kh = w.shape[1]
                n_out_rows = (a.shape[1] - kh) // sh + 1
                quant = a.dtype == np.int8
                rows = []
                for r in range(n_out_rows):
                    win = a[0, r * sh:r * sh + kh, 0, :]                         # [kh, Cin]
                    if quant:
                        s_in, zp_in = g.tensors[op.inputs[0]].per_tensor_q()
                        wi = win.astype(np.int64) - zp_in
                        if c == TF.OP_CONV_2D:
                            acc = np.einsum("kc,okc->o", wi, w[:, :, 0, :].astype(np.int64)) + bias
                        else:
                            acc = (wi * w[0, :, 0, :].astype(np.int64)).sum(0) + bias
                        rows.append(self._requant(acc, s_in, w_t, out_t, acc.shape[0], relu))
                    else:
                        if c == TF.OP_CONV_2D:
                            y = np.einsum("kc,okc->o", win.astype(np.float32), w[:, :, 0, :]).astype(np.float32) + bias
                        else:
                            y = (win.astype(np.float32) * w[0, :, 0, :]).sum(0, dtype=np.float32) + bias
                        rows.append(np.maximum(y, 0).astype(np.float32) if relu else y.astype(np.float32))
                val[op.outputs[0]] = np.stack(rows)[None, :, None, :]
            elif c == TF.OP_FULLY_CONNECTED:
                a, w_t = get(op.inputs[0]).reshape(-1), g.tensors[op.inputs[1]]
                w = w_t.data()
                bias = get(op.inputs[2]).reshape(-1) if len(op.inputs) > 2 and op.inputs[2] >= 0 else 0
                if a.dtype == np.int8:
                    s_in, zp_in = g.tensors[op.inputs[0]].per_tensor_q()
                    acc = w.astype(np.int64) @ (a.astype(np.int64) - zp_in) + bias
                    val[op.outputs[0]] = self._requant(acc, s_in, w_t, g.tensors[op.outputs[0]], acc.shape[0], False).reshape(1, -1)
                else:
                    val[op.outputs[0]] = (w @ a.astype(np.float32) + bias).astype(np.float32).reshape(1, -1)
            elif c == TF.OP_LOGISTIC:
                a = get(op.inputs[0])
                if a.dtype == np.int8:
                    s_in, zp_in = g.tensors[op.inputs[0]].per_tensor_q()
                    _, zp_out = g.tensors[op.outputs[0]].per_tensor_q()
                    lut = TF.logistic_lut(s_in, zp_in, zp_out)
                    val[op.outputs[0]] = lut[a.view(np.uint8)]
                else:
                    val[op.outputs[0]] = R._sigmoid(a.astype(np.float32)).astype(np.float32)
            elif c == TF.OP_QUANTIZE:
                a = get(op.inputs[0])
                s_in, zp_in = g.tensors[op.inputs[0]].per_tensor_q()
                s_out, zp_out = g.tensors[op.outputs[0]].per_tensor_q()
                if s_in != s_out:
                    raise NotImplementedError("QUANTIZE with a scale change")
                val[op.outputs[0]] = np.clip(a.astype(np.int32) - zp_in + zp_out, 0, 255).astype(np.uint8)
            else:
                raise NotImplementedError(op.name)
        return val[g.outputs[0]]
