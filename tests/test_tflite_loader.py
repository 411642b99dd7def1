"""the `.tflite` loader.

Three independent readings of the same TFL3 bytes must agree:
  (1) the recogniser (product, `microwakeword_b200/tflite_file.py`) -> MWW tensors,
  (2) the op-by-op interpreter (`oracle/tflite_interp.py`) executing the graph as written,
  (3) the oracle's streaming MixedNet run on the recognised tensors.
The files come from `tests/tflite_writer.py` (no TensorFlow exists here: PARITY UNPINNED for this row).
"""

import os
import struct

import numpy as np
import pytest

from microwakeword_b200 import model_file as MF
from microwakeword_b200 import tflite_file as TF
from oracle import mixednet_ref as R
from oracle.tflite_interp import Interpreter
import tflite_writer as W

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _tensors(kind):
    return MF.load(os.path.join(GOLDEN, "okay_nabu_synth_%s.mww" % kind))


@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_round_trip_is_exact(kind):
    t = _tensors(kind)
    blob = W.write_streaming_mixednet(t)
    assert TF.is_tflite(blob) and not TF.is_tflite(MF.write_container(t))
    got = TF.tensors_from_tflite(blob)
    assert sorted(got) == sorted(t)
    for k in t:
        assert got[k].dtype == t[k].dtype and got[k].shape == t[k].shape and np.array_equal(got[k], t[k]), k
    assert MF.Arch.decode(got["arch"]) == MF.OKAY_NABU
    # and the container bytes the C-ABI would receive parse back
    assert sorted(MF.read_container(TF.container_from_tflite(blob))) == sorted(t)


@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_interpreter_equals_oracle_on_recognised_tensors(kind):
    t = _tensors(kind)
    blob = W.write_streaming_mixednet(t)
    got = TF.tensors_from_tflite(blob)
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))[:150]           # uint16 [150, 40]
    it = Interpreter(blob)
    if kind == "int8":
        model = R.StreamingInt8(got)
        x = R.quantize_input(feats.astype(np.float32) * R.FEATURE_SCALE, model.input_scale, model.input_zero_point)
        for s in range(50):
            a = int(it.invoke(x[3 * s:3 * s + 3]).reshape(-1)[0])
            assert a == model.step(x[3 * s:3 * s + 3]), s
    else:
        model = R.FoldedStreamingF32(got)
        x = feats.astype(np.float32) * R.FEATURE_SCALE
        for s in range(50):
            a = float(it.invoke(x[3 * s:3 * s + 3]).reshape(-1)[0])
            assert abs(a - float(model.step(x[3 * s:3 * s + 3]))) <= 2e-6, s
    # reset re-runs the CALL_ONCE initialiser
    it.reset()
    first = it.invoke(x[0:3]).reshape(-1)[0]
    it2 = Interpreter(blob)
    assert first == it2.invoke(x[0:3]).reshape(-1)[0]


@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_block_without_mixconv_round_trips(kind):
    """mixednet.py:346-348: a block whose largest kernel is 1 has no MixConv layer -- no ring, no depthwise op, the 1x1
    conv reads the previous activation.  Writer (Keras-shaped graph) -> recogniser (identity depthwise in the container) ->
    interpreter of the written bytes == oracle on the recognised tensors; and the Keras-form model (which really has no
    layer there) == the folded container form."""
    spec = R.Spec(24, 5, 2, (32, 16, 24), ((3, 5), (1,), (7,)), head_rows=4)
    p = R.init_synthetic(spec, 11)
    assert p["b1/dw/kernels"] == [] and len(p["b0/dw/kernels"]) == 2
    t = R.fold_bn(spec, p)
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 26, (60, 2, 40)).astype(np.float32)
    keras, folded = R.KerasStreamingF32(spec, p), R.FoldedStreamingF32(t)
    assert max(abs(float(keras.step(c)) - float(folded.step(c))) for c in x) <= 2e-6
    if kind == "int8":
        t = R.quantize_model(t, x[:40])
        assert np.array_equal(t["q/b1/dw/w"], np.ones((1, 32), np.int8)) and t["q/scales"][4] == t["q/scales"][3] and t["q/zps"][4] == t["q/zps"][3]
    blob = W.write_streaming_mixednet(t)
    got = TF.tensors_from_tflite(blob)
    assert sorted(got) == sorted(t)
    for k in t:
        assert got[k].dtype == t[k].dtype and got[k].shape == t[k].shape and np.array_equal(got[k], t[k]), k
    it = Interpreter(blob)
    assert not any("dw_1" in name or "stream_2" in name for name in it.tensor_names()) if hasattr(it, "tensor_names") else True
    if kind == "int8":
        model = R.StreamingInt8(got)
        for c in x:
            qc = R.quantize_input(c, model.input_scale, model.input_zero_point)
            assert int(it.invoke(qc).reshape(-1)[0]) == model.step(qc)
    else:
        model = R.FoldedStreamingF32(got)
        for c in x:
            assert abs(float(it.invoke(c).reshape(-1)[0]) - float(model.step(c))) <= 2e-6


def test_unsupported_graphs_fail_loudly():
    t = _tensors("f32")
    # 1. not a flatbuffer / wrong identifier
    with pytest.raises(TF.TfliteError):
        TF.tensors_from_tflite(b"\0" * 64)
    blob = bytearray(W.write_streaming_mixednet(t))
    bad = bytes(blob[:4]) + b"TFL2" + bytes(blob[8:])
    with pytest.raises(TF.TfliteError):
        TF.tensors_from_tflite(bad)
    # 2. truncated file
    with pytest.raises((TF.TfliteError, struct.error, ValueError)):
        TF.tensors_from_tflite(bytes(blob[:len(blob) // 2]))
    # 3. structurally different graphs, produced by patching the writer
    class NoWriteBack(W.GraphWriter):
        def op(self, name, inputs, outputs, **options):
            if name == "ASSIGN_VARIABLE" and len(self.ops) > 40:          # drop a late state update
                return
            super().op(name, inputs, outputs, **options)

    saved = W.GraphWriter
    try:
        W.GraphWriter = NoWriteBack
        with pytest.raises(TF.TfliteError, match="never written back"):
            TF.tensors_from_tflite(W.write_streaming_mixednet(t))

        class ReluLess(saved):
            def op(self, name, inputs, outputs, **options):
                if name == "CONV_2D" and options.get("stride_h", 1) == 1:
                    options["act"] = 0
                super().op(name, inputs, outputs, **options)
        W.GraphWriter = ReluLess
        with pytest.raises(TF.TfliteError, match="without ReLU"):
            TF.tensors_from_tflite(W.write_streaming_mixednet(t))

        class WrongKeep(saved):
            def tensor(self, name, shape, dtype, data=None, **kw):
                if name.endswith("keep_0/begin"):
                    data = np.asarray([0, 0, 0, 0], np.int32)              # keeps the FIRST rows instead of the last
                return super().tensor(name, shape, dtype, data=data, **kw)
        W.GraphWriter = WrongKeep
        with pytest.raises(TF.TfliteError, match="StridedKeep"):
            TF.tensors_from_tflite(W.write_streaming_mixednet(t))
    finally:
        W.GraphWriter = saved
    # 4. an initial ring state that is not zero cannot be represented by the engine's reset
    tq = _tensors("int8")

    blob_q = bytearray(W.write_streaming_mixednet(tq))
    g = TF.Graph(bytes(blob_q))
    tensors1, ops1, _, _ = g.subgraphs[1]
    victim = [op for op in ops1 if op.code == TF.OP_ASSIGN_VARIABLE][0]
    raw = g.buffers[tensors1[victim.inputs[1]].buffer]
    off = bytes(blob_q).find(raw.tobytes())
    assert off > 0
    blob_q[off] = (blob_q[off] + 1) & 0xFF
    with pytest.raises(TF.TfliteError, match="initialiser"):
        TF.tensors_from_tflite(bytes(blob_q))


def test_flatbuffer_reader_handles_defaults_and_vtables():
    """fields omitted because they equal the schema default read back as the default"""
    t = _tensors("f32")
    g = TF.Graph(W.write_streaming_mixednet(t))
    convs = [op for op in g.ops if op.code == TF.OP_CONV_2D]
    assert convs[0].opt(2, "i", 1) == 3 and convs[1].opt(2, "i", 1) == 1           # stride_h; absent -> default 1
    assert g.tensors[g.inputs[0]].shape == (1, 3, 40) and g.tensors[g.outputs[0]].shape == (1, 1)
    assert sum(op.code == TF.OP_READ_VARIABLE for op in g.ops) == 6
    assert len(g.subgraphs) == 2 and g.version == 3


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_model_loads_tflite_and_matches_container(kind, tmp_path, torch_cuda):
    """`Model("x.tflite")` (inference.py:34-45) == `Model("x.mww")` on the GPU, probability for probability."""
    from microwakeword.inference import Model
    t = _tensors(kind)
    p = tmp_path / ("okay_nabu_synth_%s.tflite" % kind)
    p.write_bytes(W.write_streaming_mixednet(t))
    feats = np.load(os.path.join(GOLDEN, "config0_features.npy"))
    a = Model(str(p)).predict_spectrogram(feats)
    b = Model(os.path.join(GOLDEN, "okay_nabu_synth_%s.mww" % kind)).predict_spectrogram(feats)
    assert len(a) == len(b) == 332 and np.array_equal(np.asarray(a), np.asarray(b))
    m = Model(str(p))
    assert m.is_quantized_model == (kind == "int8") and m.input_feature_slices == 3
    # and against the interpreter executing the file itself (first 40 steps)
    it = Interpreter(p.read_bytes())
    if kind == "int8":
        x = R.quantize_input(feats.astype(np.float32) * R.FEATURE_SCALE, m.input_details[0]["quantization"][0], m.input_details[0]["quantization"][1])
        want = [int(it.invoke(x[3 * s:3 * s + 3]).reshape(-1)[0]) for s in range(40)]                 # uint8 output of the graph
        assert np.array_equal(np.round(np.asarray(a[:40], np.float64) * 255.0).astype(np.int64), np.asarray(want, np.int64))
    else:
        x = feats.astype(np.float32) * R.FEATURE_SCALE
        want = [float(it.invoke(x[3 * s:3 * s + 3]).reshape(-1)[0]) for s in range(40)]
        assert np.abs(np.asarray(a[:40]) - np.asarray(want)).max() <= 1e-5


@pytest.mark.parametrize("kind", ["f32", "int8"])
def test_reader_survives_truncation_and_corruption(kind):
    """a real converter-written file is the first thing that will break the reader, so it must fail
    LOUDLY and only with TfliteError -- never an IndexError / struct.error from deep inside, never a hang, never a silently
    different model.  Every prefix of the file (sampled), and single-byte corruptions all over the metadata region (vtables,
    offsets, operator codes, tensor shapes) and the data region."""
    import time
    t = _tensors(kind)
    blob = W.write_streaming_mixednet(t)
    good = TF.tensors_from_tflite(blob)
    rng = np.random.default_rng(123)
    t0 = time.time()
    cuts = sorted(set([0, 1, 4, 7, 8, 12, len(blob) - 1] + [int(v) for v in rng.integers(0, len(blob), 250)]))
    for n in cuts:
        with pytest.raises(TF.TfliteError):
            TF.tensors_from_tflite(blob[:n])
    outcomes = {"raised": 0, "same": 0, "different_weights": 0}
    positions = [int(v) for v in rng.integers(0, len(blob), 1200)]
    for pos in positions:
        b = bytearray(blob)
        b[pos] ^= int(rng.integers(1, 256))
        try:
            got = TF.tensors_from_tflite(bytes(b))
        except TF.TfliteError:
            outcomes["raised"] += 1
            continue
        # a flip that survives must have landed in tensor DATA (weights / biases / quantisation numbers): the structure --
        # architecture, tensor names, shapes, dtypes -- is what it was
        assert sorted(got) == sorted(good)
        assert np.array_equal(got["arch"], good["arch"])
        assert all(got[k].shape == good[k].shape and got[k].dtype == good[k].dtype for k in good)
        outcomes["same" if all(np.array_equal(got[k], good[k]) for k in good) else "different_weights"] += 1
    assert outcomes["raised"] > 50 and time.time() - t0 < 120, outcomes


def test_reader_accepts_operator_and_tensor_permutations():
    """The converter is free to order tensors, buffers and (topologically) operators differently from this repo's writer: the
    recogniser must depend on the dataflow only.  Tensors are renumbered at random, and independent operators are swapped."""
    t = _tensors("int8")
    base = TF.tensors_from_tflite(W.write_streaming_mixednet(t))
    for seed in range(4):
        blob = W.write_streaming_mixednet(t, shuffle_seed=seed)
        got = TF.tensors_from_tflite(blob)
        assert sorted(got) == sorted(base) and all(np.array_equal(got[k], base[k]) for k in base), seed
