"""Run-time-geometry MixedNet (microwakeword_b200/csrc/mww_nn_generic.cuh): the phase functions of the generic kernels,
executed thread by thread on the CPU (tests/host_emul), against the oracle for several architectures the reference's
builder can emit (mixednet.py flags --pointwise_filters / --mixconv_kernel_sizes / --first_conv_filters /
--first_conv_kernel_size / --stride): fp32 within 1e-5, int8 bit-exact, ring state / pending rows identical to the
oracle's after ragged call sequences.  The GPU side of the same path is tests/test_zz_generic_arch_gpu.py."""

import numpy as np
import pytest

import oracle
from conftest import synth_audio
from host_emul import emul
from microwakeword_b200 import model_file as MF
from oracle import mixednet_ref as R

ARCHS = {
    "okay_nabu": R.Spec(),
    "stride1_two_blocks": R.Spec(24, 3, 1, (32, 48), ((3,), (5, 9)), head_rows=5),
    "stride2_three_groups": R.Spec(16, 6, 2, (40,), ((7, 11, 13),), head_rows=9),
    "wide_first_conv": R.Spec(48, 7, 3, (64, 32, 96), ((9,), (3, 5, 7, 9), (1,)), head_rows=3),
    "k0_equals_stride": R.Spec(8, 4, 4, (16, 16), ((2,), (4,)), head_rows=2),
}


@pytest.fixture(autouse=True, params=["ascending", "descending"])
def thread_order(request):
    """Every test below runs twice: threads of a phase executed in ascending and in descending order.  Identical results in
    both orders rule out a read-after-write dependency between threads inside one phase (a missing barrier)."""
    emul.lib().emul_gen_thread_order(int(request.param == "descending"))
    yield
    emul.lib().emul_gen_thread_order(0)


def _features(n_streams, n_samples, seed):
    audio = np.stack([synth_audio(n_samples, seed + i) for i in range(n_streams)])
    feats, _ = oracle.run_pipeline(None, audio, want_probs=False)
    return feats                                                     # uint16 [S, T, 40]


def _models(spec, seed=0):
    f32 = R.fold_bn(spec, R.init_synthetic(spec, seed))
    calib = _features(2, 16000, 7000).reshape(-1, 40)
    calib = calib[: calib.shape[0] // spec.stride * spec.stride].reshape(-1, spec.stride, 40)
    q8 = R.quantize_model(f32, calib.astype(np.float32) * R.FEATURE_SCALE)
    return f32, q8


@pytest.mark.parametrize("name", list(ARCHS))
def test_geometry_matches_the_container_arch(name):
    spec = ARCHS[name]
    arch = MF.Arch.decode(spec.encode()) if hasattr(MF.Arch, "decode") else None
    info = emul.gen_arch_info(spec.encode())
    assert info["rc"] == 0
    ring0 = max(spec.first_conv_kernel_size - spec.stride, 0)
    state = ring0 * 40 + sum((max(ks) - 1) * spec.cin(i) for i, ks in enumerate(spec.mixconv_kernel_sizes)) + \
        (spec.head_rows - 1) * spec.pointwise_filters[-1]
    assert info["state_elems"] == state and info["ring0"] == ring0 and info["stride"] == spec.stride
    assert info["pend_cap"] == max(spec.stride - 1, 1) and info["c_last"] == spec.pointwise_filters[-1]
    if arch is not None:
        assert info["state_elems"] == arch.state_elements and info["macs_per_step"] == arch.macs_per_step
    if name == "okay_nabu":
        assert info["state_elems"] == 4176 and info["macs_per_step"] == 24800 and info["sm_elems"] * 4 < 48 * 1024


def test_unsupported_geometries_are_refused():
    bad = R.Spec(8, 2, 3, (16,), ((3,),)).encode()            # first_conv_kernel_size < stride
    assert emul.gen_arch_info(bad)["rc"] == -2
    assert emul.gen_arch_info(R.Spec(8, 5, 3, (1024,), ((3,),)).encode())["rc"] == -2
    assert emul.gen_arch_info(np.asarray([32, 5, 3, 39, 1, 17, 64, 1, 5, 0, 0, 0], np.int32))["rc"] == -1     # not 40 features
    assert emul.gen_arch_info(np.asarray([32, 5, 3, 40, 2, 17, 64, 1, 5, 0, 0, 0], np.int32))["rc"] == -1     # truncated


@pytest.mark.parametrize("name", list(ARCHS))
def test_fp32_phases_match_the_oracle(name):
    spec = ARCHS[name]
    f32, _ = _models(spec)
    feats = _features(3, 8000, 100)                                   # 47 rows per stream
    c = [oracle.MixedNet(MF.write_container(f32)) for _ in range(3)]
    g = emul.GenF32(f32, 3)
    # ragged call sequence: pending rows of every phase, a call with no full step, uint16 and float32 rows
    cuts = [0, 1, 2, 9, 9 + spec.stride - 1, 30, 47]
    got, want = [], [[] for _ in range(3)]
    fed = 0
    for a, b in zip(cuts[:-1], cuts[1:]):
        rows = feats[:, a:b]
        if (a // 3) % 2:
            rows = (rows.astype(np.float32) * R.FEATURE_SCALE).astype(np.float32)
        got.append(g.infer(rows))
    got = np.concatenate(got, 1)
    for s in range(3):
        want[s] = c[s].predict_u16(feats[s, : 47 // spec.stride * spec.stride])
    want = np.stack(want)
    assert got.shape == want.shape and got.shape[1] == 47 // spec.stride
    assert np.abs(got - want).max() <= 1e-5
    assert 0.01 < want.mean() < 0.99 and want.std() > 1e-3            # the comparison is not vacuous


@pytest.mark.parametrize("name", list(ARCHS))
def test_int8_phases_are_bit_exact(name):
    spec = ARCHS[name]
    _, q8 = _models(spec)
    feats = _features(3, 8000, 300)
    g = emul.GenI8(q8, 3)
    ref = [R.StreamingInt8(q8) for _ in range(3)]
    cuts = [0, 2, 3, 11, 11 + spec.stride - 1, 29, 47]
    got = []
    for i, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        rows = feats[:, a:b]
        if i % 3 == 1:
            rows = (rows.astype(np.float32) * R.FEATURE_SCALE).astype(np.float32)
        elif i % 3 == 2:                                              # already-quantised rows (inference.py:110)
            rows = R.quantize_input(rows.astype(np.float32) * R.FEATURE_SCALE, ref[0].input_scale, ref[0].input_zero_point)
        got.append(g.infer(rows))
    got = np.concatenate(got, 1)
    usable = 47 // spec.stride * spec.stride
    want = np.stack([oracle.MixedNet(MF.write_container(q8)).predict_u16(feats[s, :usable]) for s in range(3)])
    assert got.shape == want.shape and np.array_equal(got, want)
    assert len(np.unique(want)) > 3


@pytest.mark.parametrize("name", ["stride2_three_groups", "okay_nabu"])
def test_state_layout_is_the_oracles(name):
    """After the same rows the generic path's per-stream state (first-conv ring, block rings, head ring; oldest row first)
    and its pending rows equal the NumPy oracle's buffers element for element: mww_get_state / mww_set_state stay
    interchangeable between the compiled-in and the generic path."""
    spec = ARCHS[name]
    f32, q8 = _models(spec)
    feats = _features(1, 6000, 500)[0]                                # 35 rows
    n = 35 // spec.stride * spec.stride
    extra = min(spec.stride - 1, 35 - n)
    g = emul.GenI8(q8, 1)
    g.infer(feats[None, :n + extra])
    m = R.StreamingInt8(q8)
    R.predict_spectrogram(m, feats[:n], quantized=True, in_scale=m.input_scale, in_zp=m.input_zero_point)
    flat = np.concatenate([np.asarray(m.st_first).ravel()] + [np.asarray(r).ravel() for r in m.st_block] + [np.asarray(m.st_head).ravel()])
    assert g.state.shape[1] == flat.size and np.array_equal(g.state[0], flat.astype(np.int8))
    assert g.n_pend == extra
    if extra:
        pend_want = R.quantize_input(feats[n:n + extra].astype(np.float32) * R.FEATURE_SCALE, m.input_scale, m.input_zero_point)
        assert np.array_equal(g.pend[0, :extra * 40], pend_want.ravel())
    # fp32 twin
    gf = emul.GenF32(f32, 1)
    gf.infer(feats[None, :n])
    mf = R.FoldedStreamingF32(f32)
    R.predict_spectrogram(mf, feats[:n])
    flat_f = np.concatenate([np.asarray(mf.st_first).ravel()] + [np.asarray(r).ravel() for r in mf.st_block] + [np.asarray(mf.st_head).ravel()])
    assert np.abs(gf.state[0] - flat_f).max() <= 1e-5


def test_okay_nabu_generic_equals_the_compiled_in_kernels():
    """Same model, same rows through the specialised clip kernels' phases (tensor-core formulation) and through the generic
    phases: int8 identical, fp32 within the summation-order tolerance."""
    f32, q8 = _models(ARCHS["okay_nabu"])
    feats = _features(2, 12000, 800)
    n = feats.shape[1] // 3 * 3
    assert np.array_equal(emul.GenI8(q8, 2).infer(feats[:, :n]), emul.NnI8(q8, 2).infer(feats[:, :n]))
    assert np.abs(emul.GenF32(f32, 2).infer(feats[:, :n]) - emul.NnF32(f32, 2).infer(feats[:, :n])).max() <= 1e-5


@pytest.mark.parametrize("name", ["stride2_three_groups", "stride1_two_blocks"])      # (the test writer has no 1-tap / stateless Stream form)
def test_tflite_file_of_another_architecture_runs_on_the_generic_phases(name):
    """The whole drop-in chain for a model that is not okay_nabu, on the CPU: a streaming .tflite (tests/tflite_writer.py) ->
    the product's flatbuffer recogniser -> container tensors -> the generic kernel's phase functions, against the op-by-op
    interpreter executing the same flatbuffer bytes (oracle/tflite_interp.py): int8 identical, fp32 within 1e-5."""
    import tflite_writer as W
    from microwakeword_b200 import tflite_file as TF
    from oracle.tflite_interp import Interpreter
    spec = ARCHS[name]
    s = spec.stride
    feats = _features(1, 6000, 900)[0]
    n = feats.shape[0] // s * s
    for kind, tensors in zip(("f32", "int8"), _models(spec)):
        blob = W.write_streaming_mixednet(tensors)
        got = TF.tensors_from_tflite(blob)
        assert np.array_equal(got["arch"], spec.encode())
        it = Interpreter(blob)
        if kind == "int8":
            g = emul.GenI8(got, 1)
            x = R.quantize_input(feats[:n].astype(np.float32) * R.FEATURE_SCALE, g.in_scale, int(g.zp[0]))
            want = np.asarray([int(it.invoke(x[i:i + s]).reshape(-1)[0]) for i in range(0, n, s)], np.float32) * np.float32(1.0 / 255.0)
            assert np.array_equal(g.infer(x[None])[0], want)
        else:
            g = emul.GenF32(got, 1)
            x = (feats[:n].astype(np.float32) * R.FEATURE_SCALE).astype(np.float32)
            want = np.asarray([float(it.invoke(x[i:i + s]).reshape(-1)[0]) for i in range(0, n, s)], np.float32)
            assert np.abs(g.infer(x[None])[0] - want).max() <= 1e-5
